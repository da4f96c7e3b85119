"""Build recipe for libnb200.so: plain nvcc, sm_100a only, in-tree output (the .so travels to the GPU box).

    python -m libnano_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libnb200.so")
STAMP_PATH = LIB_PATH + ".sha256"  # travels with the .so (git-ignored, not gpurun-ignored)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]

# translation units: the host layer + small kernels, four groups of single-target kernel variants, the tensor-pipe
# kernels. They compile in parallel, and no kernel family's code depends on what else shares its module.
UNITS = ["nb200_api.cu", "host_qp.cu", "kernels_t1_a.cu", "kernels_t1_b.cu", "kernels_t1_c.cu", "kernels_t1_d.cu", "kernels_mt.cu", "kernels_ft.cu", "kernels_fd.cu", "kernels_fl.cu"]
OBJ_DIR = os.path.join(CSRC, "build")


def _nvcc() -> str:
    for candidate in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if candidate and os.path.exists(candidate):
            return candidate
    raise RuntimeError("nvcc not found: libnb200.so cannot be built (there is no CPU fallback)")


def sources():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(ROOT, "include", "nb200.h"))
    return files


def _fingerprint() -> str:
    """Content hash of every source plus the flags: file times do not survive the snapshot to the GPU box."""
    import hashlib

    digest = hashlib.sha256(" ".join(NVCC_FLAGS + LINK_FLAGS + UNITS).encode())
    for path in sources():
        digest.update(os.path.basename(path).encode())
        with open(path, "rb") as handle:
            digest.update(handle.read())
    return digest.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as handle:
        return handle.read().strip() != _fingerprint()


def _compile(unit: str, verbose: bool):
    obj = os.path.join(OBJ_DIR, unit[:-3] + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-I", os.path.join(ROOT, "include"), "-c", "-o", obj, os.path.join(CSRC, unit)]
    result = subprocess.run(cmd, capture_output=True, text=True)
    return obj, cmd, result


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
        outcomes = list(pool.map(lambda unit: _compile(unit, verbose), UNITS))
    objects = []
    for obj, cmd, result in outcomes:
        if verbose:
            sys.stderr.write(result.stderr)
        if result.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + result.stdout + result.stderr)
        objects.append(obj)
    cmd = [_nvcc(), *LINK_FLAGS, "-o", LIB_PATH, *objects]
    result = subprocess.run(cmd, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError("nvcc (link) failed:\n" + " ".join(cmd) + "\n" + result.stdout + result.stderr)
    with open(STAMP_PATH, "w") as handle:
        handle.write(_fingerprint() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
