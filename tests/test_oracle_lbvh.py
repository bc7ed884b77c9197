"""LBvh<3,int,f32> (container/Bvh.hpp:835-1000, 1229-1259), CPU only: the oracle's build / refit is
bit-exact against the reference's own LBvh on seq_exec and omp_exec (duplicate Morton codes, tiny trees, coordinates
where the 10-eps padding vanishes), its traversal agrees with brute force, and the per-node functions the CUDA kernels
call (zpc_b200/csrc/lbvh_core.cuh), compiled for the host, reproduce the oracle's arrays."""
import ctypes as C

import numpy as np
import pytest

KEYS = ("auxIndices", "parents", "levels", "leafInds")


def boxes(rs, n, dup=False, scale=1.0):
    c = rs.uniform(0, 1, (n, 3)).astype(np.float32) * np.float32(scale)
    h = rs.uniform(0.001, 0.03, (n, 3)).astype(np.float32)
    if dup:
        c[n // 3: 2 * n // 3] = c[n // 3]          # a third of the primitives share one Morton code
    return np.concatenate([c - h, c + h], 1).astype(np.float32)


CASES = [(1, 0, 1.0), (2, 0, 1.0), (3, 0, 1.0), (4, 0, 1.0), (5, 1, 1.0), (17, 0, 1.0), (1000, 0, 1.0), (1000, 1, 1.0),
         (20000, 0, 1.0), (3000, 0, 100.0), (20000, 1, 1.0)]


@pytest.mark.parametrize("n,dup,scale", CASES)
def test_lbvh_build_bit_exact_vs_reference(oracle, ref, n, dup, scale):
    rs = np.random.RandomState(n + dup)
    b = boxes(rs, n, dup, scale)
    A = oracle.lbvh_build(b)
    for nthreads in (0, 8):
        B = ref.lbvh_build(b, True, nthreads)
        for k in KEYS if n > 2 else ("auxIndices", "leafInds"):
            assert np.array_equal(A[k], B[k]), (k, nthreads)
        assert np.array_equal(A["orderedBvs"].view(np.uint32), B["orderedBvs"].view(np.uint32))


def test_lbvh_refit_bit_exact_vs_reference(oracle, ref):
    rs = np.random.RandomState(3)
    b0 = boxes(rs, 5000)
    b1 = (b0 + rs.uniform(-0.01, 0.01, (5000, 1)).astype(np.float32)).astype(np.float32)
    A = oracle.lbvh_build(b0)
    oracle.lbvh_refit(A, b1)
    assert np.array_equal(A["orderedBvs"], ref.lbvh_build_then_refit(b0, b1))
    assert np.array_equal(A["orderedBvs"], ref.lbvh_build_then_refit(b0, b1, 8))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 5000])
def test_lbvh_structure_and_traversal(oracle, n):
    rs = np.random.RandomState(10 + n)
    bvs = boxes(rs, n, dup=n > 50)
    t = oracle.lbvh_build(bvs)
    if n > 2:
        nn = 2 * n - 1
        assert np.array_equal(np.sort(t["auxIndices"][t["leafInds"]]), np.arange(n))       # every primitive is one leaf
        assert t["parents"][0] == -1 and (t["levels"][t["leafInds"]] == 0).all() and (t["levels"] >= 0).all()
        assert (np.delete(t["parents"], 0) < np.arange(1, nn)).all()                       # DFS pre-order: parents come first
        lo, hi = t["orderedBvs"][0, :3], t["orderedBvs"][0, 3:]
        assert np.array_equal(lo, bvs[:, :3].min(0)) and np.array_equal(hi, bvs[:, 3:].max(0))   # root box = union
    for _ in range(40):
        qc, qh = rs.uniform(0, 1, 3).astype(np.float32), rs.uniform(0.01, 0.1, 3).astype(np.float32)
        qb = np.concatenate([qc - qh, qc + qh])
        brute = np.nonzero(~((qb[None, :3] > bvs[:, 3:]).any(1) | (qb[None, 3:] < bvs[:, :3]).any(1)))[0]
        assert np.array_equal(np.sort(oracle.lbvh_iter_neighbors(t, qb)), brute)


@pytest.mark.parametrize("n,dup", [(3, 0), (4, 0), (5, 1), (64, 0), (1000, 1), (30000, 0), (30000, 1)])
def test_device_lbvh_functions_on_the_host_reproduce_the_oracle(oracle, n, dup):
    from tests.hostmath import build_hostmath
    hm = C.CDLL(build_hostmath())
    rs = np.random.RandomState(20 + n + dup)
    b = boxes(rs, n, dup)
    A = oracle.lbvh_build(b, False)
    box, _ = oracle.lbvh_whole_box_and_codes(b)
    nn = 2 * n - 1
    out = {k: np.full(nn if k != "leafInds" else n, -7, np.int32) for k in KEYS}
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    hm.hm_lbvh_build(C.c_int(n), p(b), p(box), p(out["auxIndices"]), p(out["parents"]), p(out["levels"]), p(out["leafInds"]))
    for k in KEYS:
        assert np.array_equal(out[k], A[k]), k


def test_device_traversal_on_the_host_visits_what_the_oracle_visits(oracle):
    """zpcb::iter_neighbors (the body of the batched query kernel) on the oracle's tree: same primitive ids, same visiting order"""
    from tests.hostmath import build_hostmath
    hm = C.CDLL(build_hostmath())
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for n in (1, 2, 3, 50, 4000):
        rs = np.random.RandomState(30 + n)
        bvs = boxes(rs, n, dup=n > 10)
        t = oracle.lbvh_build(bvs)
        lev = t["levels"] if n > 2 else np.zeros(max(n, 1), np.int32)
        for _ in range(60):
            qc, qh = rs.uniform(0, 1, 3).astype(np.float32), rs.uniform(0.01, 0.15, 3).astype(np.float32)
            qb = np.concatenate([qc - qh, qc + qh]).astype(np.float32)
            out = np.empty(n, np.int32)
            c = hm.hm_lbvh_iter_neighbors(C.c_int(n), p(t["orderedBvs"]), p(t["auxIndices"]), p(lev), p(qb), p(out), C.c_int(n))
            assert np.array_equal(out[:c], oracle.lbvh_iter_neighbors(t, qb))
