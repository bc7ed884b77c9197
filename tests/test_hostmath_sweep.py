"""The sweep phase of the binned P2G (zpc_b200/csrc/p2g_sweep.cuh) compiled for the host, CPU only.  The three variants
(3: lane = node; 4: lane = node column, three z-nodes; 5: variant 4 on packed fp32 pairs with its own record layout) read
records written by the same write_record<VAR> the kernel calls.  Variant 5 must reproduce variant 4 BIT FOR BIT — that is the
claim that lets it replace variant 4 without a new parity argument — and all three must equal the float64 evaluation of
  sum_p  W_p(o) * (A0_p + BX_p o_x + BY_p o_y + BZ_p o_z),   W = product of the quadratic B-spline weights at offset o."""
import ctypes as C

import numpy as np
import pytest

from tests.hostmath import build_hostmath


def _weights(d):   # InterpolationKernel.hpp:105-113 as polynomials in d0, per offset 0, 1, 2
    return np.stack([0.5 * d * d - 1.5 * d + 1.125, -d * d + 2.0 * d - 0.25, 0.5 * d * d - 0.5 * d + 0.125], -1)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 64, 257])
def test_sweep_variants_agree(n):
    hm = C.CDLL(build_hostmath())
    rs = np.random.RandomState(100 + n)
    d0 = rs.uniform(0.5, 1.5, (n, 3)).astype(np.float32)
    mass = rs.uniform(0.5, 2.0, n).astype(np.float32)
    A, a = rs.normal(0, 1, (n, 3)).astype(np.float32), rs.normal(0, 30, (n, 3)).astype(np.float32)
    B, Kd = rs.normal(0, 0.3, (n, 9)).astype(np.float32), rs.normal(0, 10, (n, 9)).astype(np.float32)
    out3, out4, out5 = np.zeros((27, 7), np.float32), np.zeros((9, 7, 3), np.float32), np.zeros((9, 7, 3), np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    hm.hm_p2g_sweeps(C.c_int(n), p(d0), p(mass), p(A), p(a), p(B), p(Kd), p(out3), p(out4), p(out5))
    # packed == scalar column sweep, bit for bit (including the record layout of the channel pairs)
    assert np.array_equal(out5.view(np.uint32), out4.view(np.uint32))
    # float64 evaluation of the definition
    w = _weights(d0.astype(np.float64))                                   # [n, axis, offset]
    want = np.zeros((3, 3, 3, 7))
    vec = np.concatenate([A, a], 1).astype(np.float64)                    # channels 1..6: A0
    lin = np.concatenate([B.reshape(n, 3, 3), Kd.reshape(n, 3, 3)], 1).astype(np.float64)   # [n, channel, (x, y, z)]
    for ox in range(3):
        for oy in range(3):
            for oz in range(3):
                W = w[:, 0, ox] * w[:, 1, oy] * w[:, 2, oz]
                want[ox, oy, oz, 0] = (W * mass).sum()
                val = vec + lin[:, :, 0] * ox + lin[:, :, 1] * oy + lin[:, :, 2] * oz
                want[ox, oy, oz, 1:] = (W[:, None] * val).sum(0)
    got4 = out4.reshape(3, 3, 7, 3).transpose(0, 1, 3, 2)                 # [ox, oy, oz, ch]
    got3 = out3.reshape(3, 3, 3, 7)
    scale = np.abs(want).max(axis=(0, 1, 2))
    assert (np.abs(got4 - want) <= 2e-6 * scale * max(1, np.sqrt(n) / 4)).all()
    assert (np.abs(got3 - want) <= 2e-6 * scale * max(1, np.sqrt(n) / 4)).all()
    assert np.abs(want).min() > 0


def test_partition_of_unity_through_the_sweep():
    """mass channel: the 27 weights of a particle sum to one, so the node sums of the mass channel add up to the total mass"""
    hm = C.CDLL(build_hostmath())
    rs = np.random.RandomState(5)
    n = 100
    d0 = rs.uniform(0.5, 1.5, (n, 3)).astype(np.float32)
    mass = rs.uniform(0.5, 2.0, n).astype(np.float32)
    z3, z9 = np.zeros((n, 3), np.float32), np.zeros((n, 9), np.float32)
    out3, out4, out5 = np.zeros((27, 7), np.float32), np.zeros((9, 7, 3), np.float32), np.zeros((9, 7, 3), np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    hm.hm_p2g_sweeps(C.c_int(n), p(d0), p(mass), p(z3), p(z3), p(z9), p(z9), p(out3), p(out4), p(out5))
    for out in (out3[:, 0], out4[:, 0, :], out5[:, 0, :]):
        assert abs(out.astype(np.float64).sum() - mass.astype(np.float64).sum()) <= 1e-5 * mass.sum()
    assert not out4[:, 1:, :].any() and not out5[:, 1:, :].any() and not out3[:, 1:].any()
