"""The C++ host mirror (include/kryst_b200.hpp): compiles against the C ABI (CPU), runs the reference-style tests (GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_api")


def _compile():
    lib = os.path.join(ROOT, "kryst_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"),
           "-o", EXE, "-L", lib, "-lkryst_b200", "-Wl,-rpath," + lib]
    subprocess.check_call(cmd)


def test_cpp_host_layer_compiles_and_links(built):
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_layer_reference_style_tests(built):
    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "CPP_HOST_API_OK" in r.stdout, r.stdout + r.stderr
