"""The C++ side of the boundary: include/nano_b200/function.h compiled with g++ against libnb200.so.
CPU: the headers compile (syntax + link). GPU: the compiled host program evaluates the objective like libnano would."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp")
BINARY = os.path.join(ROOT, "tests", "cpp", "host_mirror.bin")


def build():
    from libnano_b200 import build as nb_build

    lib = nb_build.build()
    cmd = ["g++", "-std=c++20", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SOURCE, lib,
           f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", BINARY]
    result = subprocess.run(cmd, capture_output=True, text=True)
    assert result.returncode == 0, result.stderr
    return BINARY


def test_cpp_host_layer_compiles_and_links():
    build()
    shim = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c++",
                           os.path.join(ROOT, "include", "nano_b200", "libnano_shim.h")], capture_output=True, text=True)
    assert shim.returncode == 0, shim.stderr  # compiles to nothing without libnano's headers, by design


@pytest.mark.gpu
def test_cpp_host_layer_evaluates_on_the_gpu():
    result = subprocess.run([build()], capture_output=True, text=True, timeout=120)
    assert result.returncode == 0 and "\nOK f=" in "\n" + result.stdout, result.stdout + result.stderr
