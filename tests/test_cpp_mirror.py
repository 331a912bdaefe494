"""include/b200zk.hpp -- the C++ host-side mirror of the reference interface -- compiles against the C ABI (CPU check) and
passes its property tests on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
LIBDIR = os.path.join(ROOT, "zcash-gpu-thesis_b200")


def _build(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", LIBDIR, "-lb200zk",
           "-L/usr/local/cuda/lib64", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_runs(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror OK" in out.stdout
