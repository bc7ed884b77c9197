#!/usr/bin/env python
"""Builds experimental copies of libzpcb200.so with extra -D flags (kernel tuning experiments).

  python benchmarks/build_exp.py name1:-DZPC_G2P_NT=128,-DZPC_G2P_MINB=8 name2:-DZPC_P2G_MINB=3

Each lands in zpc_b200/build/exp/<name>.so (git-ignored, shipped by gpurun); select one with ZPCB200_LIB=<path>."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zpc_b200 import build as b  # noqa: E402


def main():
    out = os.path.join(b.HERE, "build", "exp")
    os.makedirs(out, exist_ok=True)
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        objs, procs = [], []
        for src in b.SOURCES:
            obj = os.path.join(out, name + "_" + src.replace(".cu", ".o"))
            objs.append(obj)
            procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + flags + ["-c", os.path.join(b.CSRC, src), "-o", obj]))
        for p in procs:
            if p.wait():
                raise SystemExit("nvcc failed for " + name)
        lib = os.path.join(out, name + ".so")
        subprocess.check_call([b._nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", lib] + objs)
        for o in objs:
            os.unlink(o)
        print(lib)


if __name__ == "__main__":
    main()
