"""The softmax loss steps' numerical shortcuts on the CPU: csrc/loss.cuh's `exp_nonpositive` (host + device code) against
std::exp, and the one-logarithm-per-32-rows identity of `log_product` (tests/cpp/softmax_math.cu, built with nvcc; only
host code runs)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exp_nonpositive_and_log_product_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    binary = str(tmp_path / "softmax_math")
    build = subprocess.run([nvcc, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", binary,
                            os.path.join(ROOT, "tests", "cpp", "softmax_math.cu")], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    result = subprocess.run([binary], capture_output=True, text=True, timeout=300)
    assert result.returncode == 0 and result.stdout.strip().endswith("OK"), result.stdout + result.stderr
