"""Oracle (oracle/*.c) against the committed golden vectors, which were produced by executing the
reference itself (tests/golden/make_golden.py).  Bit-exact: the oracle restates the reference's host
arithmetic expression by expression."""
import ast
import os

import numpy as np
import pytest

from zpc_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")
MPM_CASES = ["mpm_cube6_mode0", "mpm_cube6_mode1", "mpm_cube8_rest", "mpm_cube5_neg"]


def load_case(name):
    z = np.load(os.path.join(G, name + ".npz"))
    kw = dict(ast.literal_eval(str(z["kw"])))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **kw)
    return z, P


@pytest.mark.parametrize("name", MPM_CASES)
def test_oracle_substep_matches_reference_golden(oracle, name):
    z, P = load_case(name)
    n, dx = P["x"].shape[0], P["dx"]
    ts = oracle.table_size_for(max(n // 8, 1))
    tab = oracle.partition_build(P["x"], dx, ts)
    assert tab["nblocks"] == int(z["nblocks"])
    # serial insertion order == the reference's seq_exec order: identical tables, not just identical key sets
    assert np.array_equal(tab["active_keys"], z["active_keys"])
    assert np.array_equal(tab["keys"], z["table_keys"])
    assert np.array_equal(tab["indices"], z["table_indices"])
    grid = oracle.p2g(P, tab, dx, synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
    assert np.array_equal(grid, z["grid_p2g"])
    mx = oracle.grid_update(grid, synth.DT, (0.0, synth.GRAVITY, 0.0), int(z["mode"]))
    assert np.array_equal(grid, z["grid_upd"])
    assert mx == float(z["max_vel_sqr"])
    oracle.g2p(P, tab, grid, dx, synth.DT)
    for k in "xvCF":
        assert np.array_equal(P[k], z[k]), k


def test_mass_is_conserved_on_grid(oracle):
    z, P = load_case("mpm_cube8_rest")
    assert abs(float(z["grid_p2g"][:, 0].sum(dtype=np.float64)) / float(P["m"].sum(dtype=np.float64)) - 1) < 1e-5


def test_oracle_svd_and_stress_match_golden(oracle):
    z = np.load(os.path.join(G, "svd.npz"))
    for i, F in enumerate(z["F"]):
        U, S, V = oracle.svd3(F)
        assert np.array_equal(U, z["U"][i]) and np.array_equal(S, z["S"][i]) and np.array_equal(V, z["V"][i])
        PF = oracle.stress_fixedcorotated(1.0e-6, 5.0e4, 0.4, F)
        assert np.array_equal(PF, z["PF"][i], equal_nan=True)


def test_svd_reconstructs(oracle):
    z = np.load(os.path.join(G, "svd.npz"))
    # 4 Jacobi sweeps (SVD.hpp:96): ~1e-6 near the identity (the MPM regime), percent-level on arbitrary matrices
    for i, (F, U, S, V) in enumerate(zip(z["F"][:250], z["U"], z["S"], z["V"])):
        A = F.reshape(3, 3).T
        R = U.reshape(3, 3).T @ np.diag(S) @ V.reshape(3, 3)
        tol = 1e-4 if i < 200 else 2e-2
        assert np.abs(A - R).max() <= tol * max(1.0, np.abs(A).max()), i


def test_oracle_prims_match_golden(oracle):
    z = np.load(os.path.join(G, "prims.npz"))
    for n in (0, 1, 7, 1000, 4096, 10007):
        v = np.arange(n, dtype=np.int32)
        for kind in ("u32", "i32", "u64"):
            k = z["in_u32_%d" % n] if kind != "u64" else z["in_u64_%d" % n]
            if kind == "i32":
                k = k.view(np.int32)
            ko, vo = oracle.radix_sort_pair(kind, k, v)
            assert np.array_equal(ko, z["sortpair_%s_%d_k" % (kind, n)])
            assert np.array_equal(vo, z["sortpair_%s_%d_v" % (kind, n)])
        ko, vo = oracle.radix_sort_pair("u32", z["in_u32_%d" % n], v, 4, 20)
        assert np.array_equal(ko, z["sortpair_u32_%d_bits4_20_k" % n])
        assert np.array_equal(vo, z["sortpair_u32_%d_bits4_20_v" % n])
        a = z["in_i32_%d" % n]
        assert np.array_equal(oracle.scan("exclusive", "i32", a), z["exscan_i32_%d" % n])
        assert np.array_equal(oracle.scan("inclusive", "i32", a), z["inscan_i32_%d" % n])
        for op in ("sum", "min", "max"):
            assert oracle.reduce(op, "i32", a) == z["reduce_%s_i32_%d" % (op, n)][0]


def test_oracle_eos_substep_matches_reference_golden(oracle):
    """EquationOfStateConfig branch (weakly compressible fluid, per-particle J)."""
    z, P = load_case("mpm_cube6_eos")
    n, dx = P["x"].shape[0], P["dx"]
    P["J"] = z["J_in"].copy()
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    grid = oracle.p2g_eos(P, tab, dx, synth.DT, 4.0e4, 0.01, P["volume"])
    assert np.array_equal(grid, z["grid_p2g"])
    mx = oracle.grid_update(grid, synth.DT, (0.0, synth.GRAVITY, 0.0), 1)
    assert mx == float(z["max_vel_sqr"])
    F0 = P["F"].copy()
    oracle.g2p_eos(P, tab, grid, dx, synth.DT)
    for k in ("x", "v", "C", "J"):
        assert np.array_equal(P[k], z[k]), k
    assert np.array_equal(P["F"], F0)


def test_oracle_boundary_conditions_match_reference_golden(oracle):
    """ApplyBoundaryConditionOnGridBlocks with static plane / sphere colliders, sticky / slip / separate."""
    from tests.golden.make_golden import BOUNDARY_CASES
    z, P = load_case("mpm_cube7_boundary")
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    assert np.array_equal(tab["active_keys"], z["active_keys"])
    g0 = oracle.p2g(P, tab, dx, synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
    oracle.grid_update(g0, synth.DT, (0.0, synth.GRAVITY, 0.0), 1)
    for i, (geom, ctype, p0, p1) in enumerate(BOUNDARY_CASES):
        g = g0.copy()
        oracle.apply_boundary(g, tab["active_keys"], dx, geom, ctype, p0, p1)
        assert np.array_equal(g, z["grid_%d" % i]), (geom, ctype)
        assert (g != g0).any()          # the collider actually bites in every case


def test_vonmises_golden_pins_the_oracle(oracle):
    """reference-generated von Mises vectors (39 % of the particles yield) vs the oracle, by block key"""
    import ast
    from zpc_b200 import synth
    z = np.load(os.path.join(G, "mpm_cube6_vonmises.npz"))
    kw = dict(ast.literal_eval(str(z["kw"])))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **kw)
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    g = oracle.p2g_vonmises(P, tab, dx, synth.DT, synth.MODEL["E"], synth.MODEL["nu"], float(z["ys"]), P["volume"])
    ko = tab["active_keys"]
    o = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0]))
    kr = z["active_keys"]
    r = np.lexsort((kr[:, 2], kr[:, 1], kr[:, 0]))
    assert np.array_equal(ko[o], kr[r])
    assert np.array_equal(g[o], z["grid_p2g"][r])


def test_oracle_moving_colliders_match_reference_golden(oracle):
    """moving colliders (translation, rotation, angular velocity, scaling): reference-generated grids, bit for bit"""
    from tests.parity import MOVING_COLLIDERS, motion_vec
    z, P = load_case("mpm_cube7_boundary_moving")
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    assert np.array_equal(tab["active_keys"], z["active_keys"])
    g0 = oracle.p2g(P, tab, dx, synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
    oracle.grid_update(g0, synth.DT, (0.0, synth.GRAVITY, 0.0), 1)
    for i, (geom, ctype, p0, p1, motion) in enumerate(MOVING_COLLIDERS):
        g = g0.copy()
        oracle.apply_boundary(g, tab["active_keys"], dx, geom, ctype, p0, p1, motion_vec(motion))
        assert np.array_equal(g, z["grid_%d" % i]), (geom, ctype)
        assert (g != g0).any()
