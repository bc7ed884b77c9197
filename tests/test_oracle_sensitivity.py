"""Measures how reproducible the reference's stress computation is ACROSS BUILDS OF THE SAME SOURCE: the oracle
compiled with and without FMA contraction.  This is the measured basis of tests/parity.py::RTOL_STRESS."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from zpc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rhs_moves_by_1e5_when_the_host_contracts_fma(oracle):
    with tempfile.TemporaryDirectory() as d:
        so = os.path.join(d, "liboracle_fma.so")
        try:
            subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-ffp-contract=fast", "-mfma", "-o", so,
                                   os.path.join(ROOT, "oracle", "mpm_oracle.c"), os.path.join(ROOT, "oracle", "prims_oracle.c"),
                                   "-lm"])
        except (OSError, subprocess.CalledProcessError):
            pytest.skip("gcc -mfma unavailable")
        from oracle.pyoracle import Oracle
        try:
            o2 = Oracle.__new__(Oracle)
            o2.lib = C.CDLL(so)
            for name in ("zo_table_size_for", "zo_partition_build", "zo_table_query", "zo_hash_slot0"):
                getattr(o2.lib, name).restype = C.c_int
            P = synth.elastic_cube(8, 32, jitter_F=0.05, jitter_C=0.5, shuffle_seed=11)
            n, dx = P["x"].shape[0], P["dx"]
            tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(n // 8))
            g1 = oracle.p2g(P, tab, dx, synth.DT, 5e4, 0.4, P["volume"])
            g2 = o2.p2g(P, tab, dx, synth.DT, 5e4, 0.4, P["volume"])
        except OSError:
            pytest.skip("host CPU has no FMA")
    dev = [float(np.abs(g1[:, c] - g2[:, c]).max() / np.abs(g1[:, c]).max()) for c in range(7)]
    assert max(dev[:4]) < 1e-6            # mass / momentum: plain rounding noise
    assert 2e-6 < max(dev[4:]) < 1e-4     # rhs (stress through the approximate SVD): ~1.5e-5
