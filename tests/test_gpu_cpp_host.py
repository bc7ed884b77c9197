"""Builds and runs the C++ host-mirror test (tests/cpp/test_host_mirror.cpp): the reference's own test shape
(plain main(), throws on failure) against include/zpcb200/zensim_b200.hpp + libzpcb200.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(out):
    from oracle import pyoracle
    from zpc_b200 import build
    build.build()
    pyoracle.build_oracle()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", out,
           "-L", os.path.join(ROOT, "zpc_b200"), "-lzpcb200", "-L", os.path.join(ROOT, "oracle"), "-lzpcoracle",
           "-L", os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "zpc_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.check_call(cmd)


def test_cpp_host_mirror_compiles(tmp_path):
    """no GPU needed: the header and the C++ test compile and link against the C ABI"""
    _compile(str(tmp_path / "test_host_mirror"))


@pytest.mark.gpu
def test_cpp_host_mirror_runs(tmp_path):
    exe = str(tmp_path / "test_host_mirror")
    _compile(exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all host-mirror tests passed" in r.stdout
