"""The product's AoS transfer code on the CPU: zpcp::p2g_scatter_* / p2g_scatter_core / g2p_aos_particle (what the any-order
kernels call, zpc_b200/csrc/mpm_particle.cuh + mpm_kernels.cuh) compiled for the host by tests/hostmath and run particle by
particle against the reference-generated golden vectors and the oracle, for all five constitutive models and both G2P variants —
same parity rule as the GPU tests (tests/parity.py).  The kernels add launch plumbing only (one thread per particle)."""
import ast
import ctypes as C
import os

import numpy as np
import pytest

from tests.golden.make_golden import NACC, SAND
from tests.parity import GRID_RTOL, RTOL, RTOL_STRESS, check_channels, check_particles
from zpc_b200 import api, synth

G = os.path.join(os.path.dirname(__file__), "golden")
E, NU = synth.MODEL["E"], synth.MODEL["nu"]


@pytest.fixture(scope="module")
def hm():
    from tests.hostmath import build_hostmath
    return C.CDLL(build_hostmath())


def _views(P, tab):
    """host arrays behind the ABI's view structs"""
    keep = {k: np.ascontiguousarray(P[k], np.float32) for k in ("x", "v", "m", "C", "F", "J", "logJp") if k in P}
    ptr = lambda k: keep[k].ctypes.data if k in keep else None  # noqa: E731
    pv = api.zpc_particles_view(ptr("m"), ptr("x"), ptr("v"), None, ptr("J"), ptr("F"), ptr("C"), ptr("logJp"), keep["x"].shape[0])
    tk = {k: np.ascontiguousarray(tab[k]) for k in ("keys", "indices")}
    tv = api.zpc_hashtable_view(tk["keys"].ctypes.data, tk["indices"].ctypes.data, None, None, int(tab["table_size"]), None)
    return pv, tv, keep, tk


def _p2g(hm, oracle, model, P, tab, dx, prm, En=E, nun=NU):
    pv, tv, keep, tk = _views(P, tab)
    mu, lam = oracle.lame(En, nun)
    grid = np.zeros((tab["nblocks"], 7, 64), np.float32)
    prm = np.array(prm, np.float32)
    hm.hm_p2g_aos(C.c_int(model), pv, tv, grid.ctypes.data_as(C.c_void_p), C.c_float(dx), C.c_float(synth.DT), C.c_float(P["volume"]),
                  C.c_float(mu), C.c_float(lam), prm.ctypes.data_as(C.c_void_p))
    return grid, keep


def _load(name):
    z = np.load(os.path.join(G, name + ".npz"))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    return z, P


def test_fixed_corotated_substep_on_the_host(hm, oracle):
    z, P = _load("mpm_cube6_mode1")
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    grid, keep = _p2g(hm, oracle, 0, P, tab, dx, [0])
    check_channels(grid, z["grid_p2g"], 1, "host-compiled P2G vs golden", GRID_RTOL, strict_frac=0.99)
    g = z["grid_upd"].copy()                                  # G2P from the reference's own updated grid
    pv, tv, keep, tk = _views(P, tab)
    hm.hm_g2p_aos(C.c_int(0), pv, tv, g.ctypes.data_as(C.c_void_p), C.c_float(dx), C.c_float(synth.DT))
    check_particles({k: keep[k] for k in "xvCF"}, {k: z[k] for k in "xvCF"}, dx, "host-compiled G2P vs golden")


def test_vonmises_and_eos_on_the_host(hm, oracle):
    z, P = _load("mpm_cube6_vonmises")
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    grid, _ = _p2g(hm, oracle, 1, P, tab, dx, [float(z["ys"])])
    ko, kr = tab["active_keys"], z["active_keys"]
    o, r = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0])), np.lexsort((kr[:, 2], kr[:, 1], kr[:, 0]))
    check_channels(grid[o], z["grid_p2g"][r], 1, "host-compiled von Mises P2G vs golden", GRID_RTOL)
    z, P = _load("mpm_cube6_eos")
    P["J"] = z["J_in"].copy()
    del P["F"]
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    grid, keep = _p2g(hm, oracle, 4, P, tab, dx, [4.0e4, 0.01])
    ko, kr = tab["active_keys"], z["active_keys"]
    o, r = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0])), np.lexsort((kr[:, 2], kr[:, 1], kr[:, 0]))
    check_channels(grid[o], z["grid_p2g"][r], 1, "host-compiled EOS P2G vs golden", RTOL)
    g = grid.copy()
    oracle.grid_update(g, synth.DT, (0.0, synth.GRAVITY, 0.0), 1)
    pv, tv, keep, tk = _views(P, tab)
    hm.hm_g2p_aos(C.c_int(1), pv, tv, g.ctypes.data_as(C.c_void_p), C.c_float(dx), C.c_float(synth.DT))
    for k in "xv":
        check_channels(keep[k], z[k], 1, "host-compiled EOS G2P " + k, 3e-5, floor=float(np.abs(z[k]).max()))
    check_channels(keep["J"][:, None], z["J"][:, None], 1, "host-compiled EOS G2P J", 3e-5)


@pytest.mark.parametrize("model", ["sand", "nacc"])
def test_plastic_models_on_the_host(hm, oracle, model):
    """the body of p2g_aos_plastic_kernel: grid and the logJp written back vs the reference-generated golden vectors"""
    z, P = _load("mpm_cube6_" + model)
    P["logJp"] = z["logJp_in"].copy()
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    if model == "sand":
        grid, keep = _p2g(hm, oracle, 2, P, tab, dx, [SAND["cohesion"], SAND["beta"], SAND["yieldSurface"], 1.0])
    else:
        bm, msqr = oracle.nacc_consts(NACC["E"], NACC["nu"], NACC["fa"])
        grid, keep = _p2g(hm, oracle, 3, P, tab, dx, [bm, NACC["xi"], NACC["beta"], msqr, 1.0], NACC["E"], NACC["nu"])
    ko, kr = tab["active_keys"], z["active_keys"]
    o, r = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0])), np.lexsort((kr[:, 2], kr[:, 1], kr[:, 0]))
    rhs = RTOL_STRESS if model == "sand" else 1e-3          # the reference's own build-to-build distance on this case: 1.5e-5 / 2.4e-4
    check_channels(grid[o], z["grid_p2g"][r], 1, "host-compiled %s P2G vs golden" % model, [RTOL] * 4 + [rhs] * 3)
    assert np.abs(keep["logJp"] - z["logJp"]).max() <= 2e-5
    assert np.abs(keep["logJp"] - z["logJp_in"]).max() > 1e-3


@pytest.mark.parametrize("model", [0, 1, 2, 3, 4])
def test_g2p2g_on_the_host_matches_the_restated_functor(hm, oracle, model):
    """g2p2g_particle (the body of zpcb200_g2p2g_apic's kernel) vs oracle.zo_g2p2g for all five models.  The reference's G2P2G
    cannot be compiled by gcc here (types/View.h); the oracle is a restatement of simulation/transfer/G2P2G.hpp:49-141 that the GPU
    suite pins against the reference's own functor (test_g2p2g_matches_the_references_own_functor).  This CPU test checks that the
    product code and the restatement agree, and that the functor does what its algebra says (a rigid translation field gives C = 0,
    hence the stress of F itself)."""
    P = synth.elastic_cube(6, 32, jitter_F=0.04, jitter_C=0.3, shuffle_seed=21)
    n, dx = P["x"].shape[0], P["dx"]
    rs = np.random.RandomState(5)
    En, nun = (NACC["E"], NACC["nu"]) if model == 3 else (E, NU)
    if model in (2, 3):
        P["logJp"] = rs.uniform(-1.8, 0.2, n).astype(np.float32) if model == 3 else rs.uniform(-0.06, 0.03, n).astype(np.float32)
    if model == 4:
        P["J"] = (1.0 + rs.uniform(-0.05, 0.05, n)).astype(np.float32)
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    nb = tab["nblocks"]
    gridv = rs.uniform(-1, 1, (nb * 64, 3)).astype(np.float32)
    bm, msqr = oracle.nacc_consts(NACC["E"], NACC["nu"], NACC["fa"])
    prm_o = {0: [0], 1: [2946.0], 2: [SAND["cohesion"], SAND["beta"], SAND["yieldSurface"], 1.0],
             3: [NACC["xi"], NACC["beta"], 1.0, NACC["fa"], 3.0], 4: [4.0e4, 0.01]}[model]
    prm_d = {0: [0, 0, 0, 0, 0], 1: [2946.0, 0, 0, 0, 0], 2: [SAND["cohesion"], SAND["beta"], SAND["yieldSurface"], 1.0, 0],
             3: [bm, NACC["xi"], NACC["beta"], msqr, 1.0], 4: [4.0e4, 0.01, 0, 0, 0]}[model]
    want = oracle.g2p2g(model, prm_o, P, tab, dx, synth.DT, En, nun, P["volume"], gridv)
    pv, tv, keep, tk = _views(P, tab)
    mu, lam = oracle.lame(En, nun)
    got = np.zeros_like(gridv)
    prm = np.array(prm_d, np.float32)
    hm.hm_g2p2g(C.c_int(model), pv, tv, gridv.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_float(dx), C.c_float(synth.DT),
                C.c_float(P["volume"]), C.c_float(mu), C.c_float(lam), prm.ctypes.data_as(C.c_void_p))
    scale = float(np.abs(want).max())
    assert scale > 0
    tol = 1e-3 if model == 3 else RTOL_STRESS
    assert np.abs(got - want).max() <= tol * scale, (model, np.abs(got - want).max() / scale)
    for k in ("F", "logJp", "J"):                       # particles are read only
        if k in P:
            assert np.array_equal(keep[k], P[k])
    if model == 0:
        # rigid translation: v_i = const  =>  C = 0 (partition of unity)  =>  r = sum W (P(F) F^T vol D_inv) xixp, the rhs of P2G / (-dt)
        gv = np.tile(np.float32([0.3, -0.2, 0.1]), (nb * 64, 1))
        r = oracle.g2p2g(0, [0], P, tab, dx, synth.DT, E, NU, P["volume"], gv)
        g = oracle.p2g(P, tab, dx, synth.DT, E, NU, P["volume"])
        rhs = g[:, 4:7].transpose(0, 2, 1).reshape(-1, 3) / np.float32(-synth.DT)
        assert np.abs(r - rhs).max() <= 2e-5 * np.abs(rhs).max()


def test_table_lookups_on_the_host_resolve_the_oracles_tables(hm, oracle):
    """zpcm::table_query (legacy HashTable: 64-bit hash_combine, stride-127 probing) and zpcm::bht_query (three universal hashes,
    buckets of 16) — the lookups every transfer kernel performs — on tables the ORACLE built (i.e. the reference's insert order):
    every active key resolves to its index, absent keys to -1"""
    rs = np.random.RandomState(3)
    P = synth.elastic_cube(9, 32, origin_cells=-11, shuffle_seed=2)
    n, dx = P["x"].shape[0], P["dx"]
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    nb = tab["nblocks"]
    absent = (tab["active_keys"][:50] + np.int32([1000, 0, 0])).astype(np.int32)
    keys = np.ascontiguousarray(np.concatenate([tab["active_keys"], absent]), np.int32)
    out = np.empty(keys.shape[0], np.int32)
    tk, ti = np.ascontiguousarray(tab["keys"]), np.ascontiguousarray(tab["indices"])
    hm.hm_table_query(C.c_int(keys.shape[0]), keys.ctypes.data_as(C.c_void_p), C.c_int(tab["table_size"]), tk.ctypes.data_as(C.c_void_p),
                      ti.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out[:nb], np.arange(nb)) and (out[nb:] == -1).all()
    # bht: block ORIGINS in cell coordinates as keys
    t = oracle.bht_new(4 * nb)
    bkeys = np.ascontiguousarray(tab["active_keys"] * 8, np.int32)
    idx = oracle.bht_insert(t, bkeys)
    assert np.array_equal(np.sort(idx), np.arange(nb))
    hf = oracle.bht_params()
    view = api.zpc_bht_view(t["keys16"].ctypes.data, t["indices"].ctypes.data, None, None, int(t["table_size"]), int(t["table_size"]) // 16,
                            None, None, (C.c_uint32 * 6)(*[int(v) for v in hf]))
    q = np.ascontiguousarray(np.concatenate([bkeys, bkeys[:50] + np.int32([8000, 0, 0])]), np.int32)
    out2 = np.empty(q.shape[0], np.int32)
    hm.hm_bht_query(C.c_int(q.shape[0]), q.ctypes.data_as(C.c_void_p), view, out2.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out2[:nb], idx) and (out2[nb:] == -1).all()
