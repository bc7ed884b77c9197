"""Oracle against the reference itself (oracle/_ref/libzpcref.so) on fresh seeded inputs, larger than the
committed golden vectors, and against the reference's OpenMP policy.  Skipped where the reference library
was not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

from zpc_b200 import synth
from tests.parity import check_channels, grid_by_key


def _copy(P):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in P.items()}


@pytest.mark.parametrize("seed,s,G,origin", [(1, 10, 32, 7), (2, 7, 16, -11)])
def test_substep_bit_exact_vs_reference_seq(oracle, ref, seed, s, G, origin):
    P = synth.elastic_cube(s, G, seed=seed, jitter_F=0.08, jitter_C=1.0, shuffle_seed=seed + 5, origin_cells=origin)
    n, dx = P["x"].shape[0], P["dx"]
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    nb = h.partition()
    tab_r = h.table()
    tab_o = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    assert nb == tab_o["nblocks"] and np.array_equal(tab_r["active_keys"], tab_o["active_keys"])
    assert np.array_equal(tab_r["keys"], tab_o["keys"]) and np.array_equal(tab_r["indices"], tab_o["indices"])
    h.clean_grid()
    h.p2g(synth.DT, 5e4, 0.4, P["volume"])
    g_o = oracle.p2g(P, tab_o, dx, synth.DT, 5e4, 0.4, P["volume"])
    assert np.array_equal(h.grid(), g_o)
    for mode in (1,):
        mx_r = h.grid_update(synth.DT, synth.GRAVITY, mode)
        mx_o = oracle.grid_update(g_o, synth.DT, (0, synth.GRAVITY, 0), mode)
        assert np.array_equal(h.grid(), g_o) and mx_r == mx_o
    h.g2p(synth.DT)
    Po = _copy(P)
    oracle.g2p(Po, tab_o, g_o, dx, synth.DT)
    Pr = h.get_particles()
    h.close()
    for k in "xvCF":
        assert np.array_equal(Pr[k], Po[k]), k


def test_reference_openmp_agrees_within_tolerance(oracle, ref):
    """The reference's own parallel path re-orders the float atomics and renumbers blocks: compare by block
    key with the fp32 tolerance — this is the rule the GPU tests apply."""
    P = synth.elastic_cube(9, 32, seed=3, jitter_F=0.05, jitter_C=0.5, shuffle_seed=8)
    n, dx = P["x"].shape[0], P["dx"]
    h = ref.mpm(n, dx, 4)
    h.set_particles(P)
    h.partition()
    keys_r = h.keys()
    h.clean_grid()
    h.p2g(synth.DT, 5e4, 0.4, P["volume"])
    kr, gr = grid_by_key(keys_r, h.grid())
    tab_o = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    ko, go = grid_by_key(tab_o["active_keys"], oracle.p2g(P, tab_o, dx, synth.DT, 5e4, 0.4, P["volume"]))
    h.close()
    assert np.array_equal(kr, ko)
    check_channels(gr, go, 1, "omp p2g")


@pytest.mark.parametrize("nthreads", [0, 4])
def test_primitives_vs_reference(oracle, ref, nthreads):
    rs = np.random.RandomState(9)
    for n in (0, 1, 100, 70001):
        k = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
        v = np.arange(n, dtype=np.int32)
        for kind, kk in (("u32", k), ("i32", k.view(np.int32))):
            a, b = oracle.radix_sort_pair(kind, kk, v), ref.radix_sort_pair(kind, kk, v, nthreads=nthreads)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        a, b = oracle.radix_sort_pair("u32", k, v, 3, 17), ref.radix_sort_pair("u32", k, v, 3, 17, nthreads=nthreads)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        x = rs.randint(-99, 99, size=n).astype(np.int32)
        if n:
            assert np.array_equal(oracle.scan("exclusive", "i32", x), ref.scan("exclusive", "i32", x, nthreads))
            assert np.array_equal(oracle.scan("inclusive", "i32", x), ref.scan("inclusive", "i32", x, nthreads))
        for op in ("sum", "prod", "min", "max"):   # prod wraps mod 2^32 on both sides (py_interop/cuda/ExecutionPolicy.cpp:48-54)
            assert oracle.reduce(op, "i32", x) == ref.reduce(op, "i32", x, nthreads)


@pytest.mark.parametrize("kind,dt", [("i32", np.int32), ("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("nthreads", [0, 4])
def test_merge_sort_pair_matches_reference(oracle, ref, kind, dt, nthreads):
    """reference merge_sort_pair (stable) on seq_exec / omp_exec vs the oracle, incl. duplicates and signed zeros"""
    rs = np.random.RandomState(11)
    for n in (1, 2, 17, 1000, 50001):
        k = rs.randint(-50, 50, size=n).astype(dt)
        if kind != "i32" and n > 4:
            k[0] = -0.0; k[2] = 0.0
        v = np.arange(n, dtype=np.int32)
        kr, vr = ref.merge_sort_pair(kind, k, v, nthreads)
        ko, vo = oracle.merge_sort_pair(kind, k, v)
        assert np.array_equal(vr, vo) and np.array_equal(kr.view(np.uint8), ko.view(np.uint8)), (kind, n)


def test_f64_scan_reduce_match_reference(oracle, ref):
    a = np.random.RandomState(3).uniform(-1, 1, 5000).astype(np.float64)
    for nthreads in (0,):
        assert np.array_equal(ref.scan("exclusive", "f64", a, nthreads), oracle.scan("exclusive", "f64", a))
        assert np.array_equal(ref.scan("inclusive", "f64", a, nthreads), oracle.scan("inclusive", "f64", a))
        for op in ("sum", "prod", "min", "max"):
            assert ref.reduce(op, "f64", a, nthreads) == oracle.reduce(op, "f64", a)


def test_vonmises_stress_and_p2g_bit_exact_vs_reference(oracle, ref):
    """compute_stress_vonmisesfixedcorotated (ConstitutiveModel_Vol_dP.hpp:49-110) in the elastic regime (huge yield
    stress) and deep in the plastic regime (radial return, projected F), then the whole P2G functor"""
    from zpc_b200 import synth
    rs = np.random.RandomState(8)
    E, nu, vol = synth.MODEL["E"], synth.MODEL["nu"], 1e-6
    yielded = 0
    for ys in (240e6, 2000.0, 50.0):
        for _ in range(200):
            F = (np.eye(3) + rs.uniform(-0.2, 0.2, (3, 3))).astype(np.float32).reshape(9)
            a, b = ref.stress_vonmises(vol, E, nu, ys, F), oracle.stress_vonmises(vol, E, nu, ys, F)
            assert np.array_equal(a, b), (ys, F)
            yielded += int(not np.array_equal(b, oracle.stress_fixedcorotated(vol, E, nu, F)))
    assert 200 < yielded <= 400                       # the two small yield stresses really project
    P = synth.elastic_cube(6, 16, jitter_F=0.1, jitter_C=0.5, shuffle_seed=2)
    n, dx = P["x"].shape[0], P["dx"]
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g_vonmises(synth.DT, E, nu, 300.0, P["volume"])
    g_ref = h.grid()
    h.close()
    g = oracle.p2g_vonmises(P, tab, dx, synth.DT, E, nu, 300.0, P["volume"])
    assert np.array_equal(g, g_ref)
    assert not np.array_equal(g, oracle.p2g(P, tab, dx, synth.DT, E, nu, P["volume"]))


from tests.parity import MOVING_COLLIDERS as MOVING, motion_vec  # noqa: E402


def test_moving_colliders_bit_exact_vs_reference(oracle, ref):
    """Collider::resolveCollision with translation, rotation, angular velocity and scaling (geometry/Collider.h:16-24,98-127)"""
    from zpc_b200 import synth
    P = synth.elastic_cube(7, 32, jitter_C=0.6, jitter_F=0.03, shuffle_seed=4)
    n, dx = P["x"].shape[0], P["dx"]
    for geom, ctype, p0, p1, motion in MOVING:
        h = ref.mpm(n, dx, 0)
        h.set_particles(P)
        h.partition()
        tab = h.table()
        h.clean_grid()
        h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
        h.grid_update(synth.DT, synth.GRAVITY, 1)
        before = h.grid()
        h.apply_boundary(geom, ctype, p0, p1, motion_vec(motion))
        after = h.grid()
        h.close()
        g = before.copy()
        oracle.apply_boundary(g, tab["active_keys"], dx, geom, ctype, p0, p1, motion_vec(motion))
        assert np.array_equal(g, after), (geom, ctype)
        assert (after != before).any(), "collider %d/%d touches nothing" % (geom, ctype)
