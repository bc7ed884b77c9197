"""AnalyticLevelSet<Cuboid> colliders (geometry/AnalyticLevelSet.h:55-126 under Collider.h:98-127),
CPU only: the oracle is bit-exact against the reference (static and moving boxes, sticky / slip / separate), the golden
vectors pin it where the reference is absent, and the product's device functions compiled for the host
(zpcm::cuboid_sdf / cuboid_normal) agree with it bit for bit — they must: the normal is a central difference with
eps = 1e-6 in float, which turns a one-ulp difference of the distance into a 1e-2 difference of the normal."""
import ast
import ctypes as C
import os

import numpy as np

from tests.parity import CUBOID_COLLIDERS, motion_vec
from zpc_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def _state(ref_or_oracle_tab=None):
    return synth.elastic_cube(7, 32, jitter_C=0.6, jitter_F=0.03, shuffle_seed=4)


def test_cuboid_colliders_bit_exact_vs_reference(oracle, ref):
    P = _state()
    n, dx = P["x"].shape[0], P["dx"]
    for geom, ctype, p0, p1, motion in CUBOID_COLLIDERS:
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
        assert np.array_equal(g.view(np.uint32), after.view(np.uint32)), (ctype, motion is not None)
        assert (after != before).any(axis=1).sum() > 20, "box %d touches too few cells" % ctype
        assert not np.isnan(after).any()


def test_cuboid_golden_pins_the_oracle(oracle):
    z = np.load(os.path.join(G, "mpm_cube7_boundary_cuboid.npz"))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    assert np.array_equal(tab["active_keys"], z["active_keys"])
    g0 = oracle.p2g(P, tab, dx, synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
    oracle.grid_update(g0, synth.DT, (0.0, synth.GRAVITY, 0.0), 1)
    for i, (geom, ctype, p0, p1, motion) in enumerate(CUBOID_COLLIDERS):
        g = g0.copy()
        oracle.apply_boundary(g, tab["active_keys"], dx, geom, ctype, p0, p1, motion_vec(motion))
        assert np.array_equal(g, z["grid_%d" % i]), (geom, ctype)
        assert (g != g0).any()


def test_device_cuboid_functions_on_the_host_are_bit_exact(oracle):
    from tests.hostmath import build_hostmath
    hm = C.CDLL(build_hostmath())
    rs = np.random.RandomState(4)
    mn, mx = np.float32([0.2, 0.1, 0.25]), np.float32([0.45, 0.3, 0.33])
    # inside, outside, and within a few eps of the faces / edges (where the central difference straddles a kink)
    x = np.concatenate([rs.uniform(0.0, 0.6, (4000, 3)),
                        mn[None] + rs.uniform(-3e-6, 3e-6, (500, 3)) + (mx - mn)[None] * rs.randint(0, 2, (500, 3)),
                        ((mn + mx) / 2)[None] + rs.uniform(-1e-3, 1e-3, (500, 3))]).astype(np.float32)
    n = x.shape[0]
    sdf, nm = np.empty(n, np.float32), np.empty((n, 3), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    hm.hm_cuboid(C.c_int(n), p(x), p(mn), p(mx), p(sdf), p(nm))
    so, no = oracle.cuboid(x, mn, mx)
    assert np.array_equal(sdf.view(np.uint32), so.view(np.uint32))
    assert np.array_equal(nm.view(np.uint32), no.view(np.uint32))
    assert (so < 0).sum() > 100 and (so > 0).sum() > 1000


def test_device_collide_on_the_host_equals_the_oracle_for_every_collider():
    """zpcm::collide (the function both boundary kernels call) compiled for the host vs the oracle's
    ApplyBoundaryConditionOnGridBlocks on a synthetic grid whose every cell is occupied: plane / sphere / cuboid, sticky /
    slip / separate, static and moving — value for value (the host build does not contract, like the oracle)."""
    from oracle.pyoracle import Oracle
    from tests.golden.make_golden import BOUNDARY_CASES
    from tests.hostmath import build_hostmath
    from tests.parity import MOVING_COLLIDERS
    from zpc_b200 import api
    oracle = Oracle()
    hm = C.CDLL(build_hostmath())
    rs = np.random.RandomState(9)
    dx = np.float32(1.0 / 32)
    keys = np.array([[bx, by, bz] for bx in range(1, 4) for by in range(1, 4) for bz in range(1, 4)], np.int32)
    nb = keys.shape[0]
    grid = np.zeros((nb, 7, 64), np.float32)
    grid[:, 0] = 1.0                                                    # mass > 0 everywhere
    grid[:, 1:4] = rs.uniform(-1, 1, (nb, 3, 64)).astype(np.float32)
    cc = np.array([[(c >> 4) & 3, (c >> 2) & 3, c & 3] for c in range(64)], np.float32)
    pos = ((keys[:, None, :].astype(np.float32) * np.float32(4.0) + cc[None]) * dx).astype(np.float32).reshape(-1, 3)
    vel0 = np.ascontiguousarray(grid[:, 1:4].transpose(0, 2, 1).reshape(-1, 3))
    cases = [(g, t, p0, p1, None) for g, t, p0, p1 in BOUNDARY_CASES] + list(MOVING_COLLIDERS) + list(CUBOID_COLLIDERS)
    # colliders placed inside this grid's extent (0.125 .. 0.5)
    for geom, ctype, p0, p1, motion in cases:
        kw = {}
        if motion is not None:
            b, dbdt, R, om, s, dsdt = motion
            kw = dict(translation=b, velocity=dbdt, rotation=np.asarray(R).tolist(), omega=om, scale=s, dscale_dt=dsdt)
        col = api._collider(geom, ctype, p0, p1, **kw)
        want = grid.copy()
        oracle.apply_boundary(want, keys, float(dx), geom, ctype, p0, p1, motion_vec(motion))
        v = vel0.copy()
        hm.hm_collide(C.c_int(v.shape[0]), col, pos.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        got = v.reshape(nb, 64, 3).transpose(0, 2, 1)
        assert np.array_equal(got, want[:, 1:4]), (geom, ctype, motion is not None)
        assert (want[:, 1:4] != grid[:, 1:4]).any(), "collider (%d, %d) touches nothing on this grid" % (geom, ctype)
