"""TEST INFRASTRUCTURE: builds tests/hostmath/_hostmath.so — the product's __host__ __device__ per-particle math
(zpc_b200/csrc/mpm_math.cuh) compiled for the CPU by nvcc's host pass, so it can be checked without a GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostmath.so")


def build_hostmath(force=False):
    src = os.path.join(HERE, "hostmath.cu")
    deps = [src] + [os.path.join(HERE, "..", "..", "zpc_b200", "csrc", f) for f in ("mpm_math.cuh", "lbvh_core.cuh", "mpm_particle.cuh", "mpm_kernels.cuh", "common.cuh", "p2g_sweep.cuh")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # -ffp-contract=off: the host pass must not fuse, like the reference's host build (oracle/Makefile)
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler",
                           "-ffp-contract=off", "-shared", "-o", SO, src])
    return SO
