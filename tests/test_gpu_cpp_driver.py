"""Builds tests/cpp/kat_driver.cpp against include/msbwt_gpu.hpp + the in-tree .so and runs
the reference's known-answer tests from C++ through the C ABI (GPU box only)."""
import os
import subprocess

import pytest

import rust_msbwt_b200 as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    exe = str(tmp_path / "kat_driver")
    libdir = os.path.dirname(M.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "kat_driver.cpp"), "-o", exe,
                           "-L", libdir, "-lmsbwt_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_cpp_host_layer_compiles(tmp_path):
    _compile(tmp_path)  # CPU: header + ABI link check only


@pytest.mark.gpu
def test_cpp_kat_driver(tmp_path, two_string_npy):
    exe = _compile(tmp_path)
    res = subprocess.run([exe, two_string_npy], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all passed" in res.stdout
