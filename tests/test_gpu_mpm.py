"""CUDA MPM path (through the C ABI) against the oracle and the reference-generated golden vectors."""
import ast
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from zpc_b200 import synth  # noqa: E402
from tests.parity import GRID_RTOL, RTOL, RTOL_STRESS, check_channels, check_particles, grid_by_key  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
E, NU = synth.MODEL["E"], synth.MODEL["nu"]


def _copy(P):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in P.items()}


def host_table(table):
    nb = table.size()
    return dict(keys=table.keys.cpu().numpy(), indices=table.indices.cpu().numpy(), table_size=table.table_size,
                nblocks=nb, active_keys=table.active_keys[:nb].cpu().numpy())


def build_partition(P, expected=None):
    from zpc_b200 import api
    n = P["x"].shape[0]
    pars = api.Particles(P)
    table = api.HashTable(expected or max(n // 8, 64))
    api.partition_for_particles(api.vec3_port(pars.x), n, P["dx"], table)
    torch.cuda.synchronize()
    assert table.overflow.item() == 0
    return pars, table


@pytest.fixture(params=[(6, 1), (4, 1), (3, 0)], ids=["plane-staged", "sweep4-staged", "sweep3-plain"])
def binned_variant(request):
    """every binned-path test runs on both kernel variants of the binned P2G / G2P (zpcb200_set_tuning)"""
    from zpc_b200 import api
    api.set_tuning(*request.param)
    yield request.param
    api.set_tuning(4, 1)


CASES = {
    "cube8": dict(s=8, G=32, jitter_F=0.05, jitter_C=0.5, shuffle_seed=11),
    "cube12_sorted": dict(s=12, G=32, jitter_F=0.03, jitter_C=0.3),
    "cube7_negative_coords": dict(s=7, G=16, jitter_F=0.05, jitter_C=0.5, shuffle_seed=3, origin_cells=-13),
    "cube16_rest": dict(s=16, G=64),
}


def make(case):
    kw = dict(CASES[case])
    return synth.elastic_cube(kw.pop("s"), kw.pop("G"), **kw)


@pytest.mark.parametrize("case", list(CASES))
def test_partition_matches_oracle(oracle, case):
    P = make(case)
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    tab_o = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    ko = tab_o["active_keys"]
    order = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0]))
    assert ht["nblocks"] == tab_o["nblocks"]
    # same active block set; our numbering is deterministic: index == lexicographic rank
    assert np.array_equal(ht["active_keys"], ko[order])
    # the table is readable by the reference's query (oracle restatement of HashTable.hpp:447-456): bijection onto [0,cnt)
    got = np.array([oracle.table_query(k, ht) for k in ht["active_keys"]])
    assert np.array_equal(got, np.arange(ht["nblocks"]))
    assert oracle.table_query(np.array([999, 999, 999], np.int32), ht) == -1
    # untouched slots keep the CleanSparsity sentinels
    free = ht["indices"] == -1
    assert free.sum() == ht["table_size"] - ht["nblocks"]
    assert (ht["keys"][free] == np.iinfo(np.int32).max).all() and (table.status.cpu().numpy() == -1).all()


def test_wide_block_codes_far_from_the_origin(oracle):
    """a cloud 6 000 cells away from the origin: block coordinates ~1 500, outside the 10 bits per axis of the default block codes —
    zpcb200_partition_build raises *overflow, zpcb200_partition_build_wide (3 x 21-bit codes in 64 bits) builds the same table the
    oracle builds, and the whole substep on it matches the oracle"""
    from zpc_b200 import api
    P = synth.elastic_cube(8, 32, jitter_F=0.04, jitter_C=0.3, shuffle_seed=2, origin_cells=6000)
    P["x"][:, 1] -= np.float32(9000 * P["dx"])          # and negative on one axis
    n, dx = P["x"].shape[0], P["dx"]
    pars = api.Particles(P)
    table = api.HashTable(max(n // 8, 64))
    api.partition_for_particles(api.vec3_port(pars.x), n, dx, table)
    torch.cuda.synchronize()
    assert table.overflow.item() == 1
    table.overflow.zero_()
    api.partition_for_particles(api.vec3_port(pars.x), n, dx, table, wide=True)
    torch.cuda.synchronize()
    assert table.overflow.item() == 0
    ht = host_table(table)
    tab_o = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    ko = tab_o["active_keys"]
    order = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0]))
    assert ht["nblocks"] == tab_o["nblocks"] and np.array_equal(ht["active_keys"], ko[order])
    assert np.abs(ht["active_keys"]).max() > 600
    got = np.array([oracle.table_query(k, ht) for k in ht["active_keys"]])
    assert np.array_equal(got, np.arange(ht["nblocks"]))
    # the transfers do not care how the table was built: binned substep on it vs the oracle.  Positions are ~25 (6 000 cells of
    # 1/256): an fp32 ulp there is 2e-6, 5e-4 of a cell — the parity rule's scales (max |x|) absorb that on both sides
    bins = api.ParticleBins(n, max(ht["nblocks"] * 2, 64))
    order_t = torch.empty(n, dtype=torch.int32, device="cuda")
    api.bin_particles(pars, table, dx, bins, order_t)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(bins, table, grids, synth.DT, api.model_fcr(P["volume"], E, NU))
    o1, o2, omx, Po = run_oracle_on_table(oracle, P, ht, 1)
    check_channels(grids.tiles.cpu().numpy(), o1, 1, "binned p2g far from the origin", GRID_RTOL, strict_frac=0.99)
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    api.g2p_transfer(bins, table, grids, synth.DT)
    perm = order_t.cpu().numpy()
    check_particles({k: bins.attr(k).cpu().numpy() for k in "xvCF"}, {k: Po[k][perm] for k in "xvCF"}, dx, "binned g2p far from the origin")
    assert int(bins.status.item()) == 0


def test_partition_empty_and_aosoa_port(oracle):
    from zpc_b200 import api
    table = api.HashTable(64)
    x = torch.zeros(0, 3, device="cuda")
    api.partition_for_particles(api.zpc_port(x.data_ptr(), 0, 0, 0, 3), 0, 0.1, table)
    assert table.size() == 0
    P = make("cube8")
    n = P["x"].shape[0]
    tv = api.TileVector(n, api.PB_NCH)
    tv.set_channel(api.PB_X, torch.from_numpy(P["x"]).cuda())
    t2 = api.HashTable(max(n // 8, 64))
    api.partition_for_particles(tv.port(api.PB_X), n, P["dx"], t2)
    _, t1 = build_partition(P)
    assert t1.size() == t2.size() and torch.equal(t1.active_keys[: t1.size()], t2.active_keys[: t2.size()])


def run_oracle_on_table(oracle, P, ht, mode):
    dx = P["dx"]
    g1 = oracle.p2g(P, ht, dx, synth.DT, E, NU, P["volume"])
    g2 = g1.copy()
    mx = oracle.grid_update(g2, synth.DT, (0.0, synth.GRAVITY, 0.0), mode)
    Po = _copy(P)
    oracle.g2p(Po, ht, g2, dx, synth.DT)
    return g1, g2, mx, Po


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", list(CASES))
def test_aos_path_matches_oracle(oracle, case, mode):
    """drop-in path: reference Particles layout, any order (P2G.hpp / GridOp.hpp / G2P.hpp)."""
    from zpc_b200 import api
    P = make(case)
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    grids = api.Grids(dx, ht["nblocks"] + 3)
    grids.tiles.fill_(123.0)                                   # CleanGridBlocks must clear exactly cnt blocks
    api.clean_grid_blocks(grids, table)
    assert float(grids.tiles[: ht["nblocks"]].abs().max()) == 0.0 and float(grids.tiles[ht["nblocks"]:].min()) == 123.0
    model = api.model_fcr(P["volume"], E, NU)
    api.p2g_transfer(pars, table, grids, synth.DT, model)
    g1 = grids.tiles[: ht["nblocks"]].cpu().numpy()
    o1, o2, omx, Po = run_oracle_on_table(oracle, P, ht, mode)
    check_channels(g1, o1, 1, "p2g", GRID_RTOL, strict_frac=0.99)
    # mass conservation (size independent)
    assert abs(g1[:, 0].sum(dtype=np.float64) / P["m"].sum(dtype=np.float64) - 1) < 1e-5
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), mode, mx)
    g2 = grids.tiles[: ht["nblocks"]].cpu().numpy()
    check_channels(g2[:, 1:4], o2[:, 1:4], 1, "grid velocity")
    assert abs(mx.item() - omx) <= 1e-5 * omx
    api.g2p_transfer(pars, table, grids, synth.DT)
    check_particles(pars.to_host(), Po, dx, "g2p")


@pytest.mark.parametrize("case", list(CASES))
def test_binned_path_matches_oracle(oracle, case, binned_variant):
    """fast path: bin -> P2G (smem arena, bulk reduce-add) -> update -> G2P (TMA staged) -> unbin."""
    from zpc_b200 import api
    P = make(case)
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    bins = api.ParticleBins(n, max(ht["nblocks"] * 2, 64))
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    api.bin_particles(pars, table, dx, bins, order)
    torch.cuda.synchronize()
    perm = order.cpu().numpy()
    assert np.array_equal(np.sort(perm), np.arange(n))                        # a permutation
    nbins = bins.num_bins.item()
    bs = bins.bin_start[: nbins + 1].cpu().numpy()
    assert bs[0] == 0 and bs[-1] == n and (np.diff(bs) > 0).all() and (np.diff(bs) <= api.BIN_MAX).all()
    # binned attributes are the permuted inputs, bit for bit, and every particle sits in its home block's bin
    for k in "xvCF":
        assert np.array_equal(bins.attr(k).cpu().numpy(), P[k][perm]), k
    bk = bins.bin_key[:nbins].cpu().numpy()
    home = np.empty((n, 3), np.int32)
    import ctypes as C
    for i in range(0, n, max(n // 200, 1)):
        b = np.zeros(3, np.int32)
        oracle.lib.zo_block_of_particle(P["x"][perm[i]].ctypes.data_as(C.c_void_p), C.c_float(dx), b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(bk[np.searchsorted(bs, i, side="right") - 1], b)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    model = api.model_fcr(P["volume"], E, NU)
    api.p2g_transfer(bins, table, grids, synth.DT, model)
    g1 = grids.tiles.cpu().numpy()
    o1, o2, omx, Po = run_oracle_on_table(oracle, P, ht, 1)
    check_channels(g1, o1, 1, "binned p2g", GRID_RTOL, strict_frac=0.99)
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    api.g2p_transfer(bins, table, grids, synth.DT)
    check_particles({k: bins.attr(k).cpu().numpy() for k in "xvCF"}, {k: Po[k][perm] for k in "xvCF"}, dx, "binned g2p")
    # unbin restores slot order
    back = api.Particles(P)
    api.unbin_particles(bins, back)
    assert np.array_equal(back.x.cpu().numpy(), bins.attr("x").cpu().numpy())


def test_binned_multistep_with_strays_matches_oracle(oracle, binned_variant):
    """5 substeps without re-binning (particles drift across cells and blocks -> stray path + arena margin),
    then a re-bin, then 2 more; the oracle runs the reference's composed substep on the same particles."""
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(10, 32, jitter_F=0.03, jitter_C=0.3, seed=5)
    P["v"][:] = P["v"] * 8.0                      # ~0.26 cell per step: plenty of cell and block crossings
    n0 = P["m"].shape[0]
    P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n0) / n0)).astype(np.float32)   # unique masses = particle identity
    assert np.unique(P["m"]).size == n0
    dx = P["dx"]
    sol = MpmSolver(P, dx, P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout="binned", rebin_every=5)
    perm = sol.order.cpu().numpy()
    Po = {k: (np.ascontiguousarray(v[perm]) if isinstance(v, np.ndarray) else v) for k, v in P.items()}
    for step in range(7):
        sol.substep()
        oracle.substep(Po, dx, synth.DT * 10, E, NU, P["volume"], synth.GRAVITY, 1)
    torch.cuda.synchronize()
    got = sol.particles_host()
    # the re-bin permuted the slots again: match particles through their (unique, never modified) mass
    def canon(Q):
        o = np.argsort(Q["m"], kind="stable")
        return {k: Q[k][o] for k in "xvCF"}
    a, b = canon(got), canon(Po)
    check_particles(a, b, dx, "multistep (7 substeps)", rtol=5e-5)   # rounding differences compound over substeps


def test_partition_with_rebin_equals_partition_every_step():
    """partition="with_rebin" (one extra ring, rebuilt only at re-bin time) must give the same particles as rebuilding
    the reference's EnlargeSparsity{0,2} partition every substep: the extra blocks stay empty."""
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(10, 32, jitter_F=0.03, jitter_C=0.3, seed=9)
    P["v"][:] = P["v"] * 6.0
    n0 = P["m"].shape[0]
    P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n0) / n0)).astype(np.float32)
    res = []
    for mode in ("every_step", "with_rebin"):
        sol = MpmSolver(P, P["dx"], P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout="binned", rebin_every=3, partition=mode)
        for _ in range(6):
            sol.substep()
        torch.cuda.synchronize()
        Q = sol.particles_host()
        o = np.argsort(Q["m"], kind="stable")
        res.append({k: Q[k][o] for k in "xvCF"})
        assert sol.table.overflow.item() == 0
    check_particles(res[0], res[1], P["dx"], "with_rebin vs every_step", rtol=1e-5)


def test_aos_multistep_matches_oracle(oracle):
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(8, 32, jitter_F=0.03, jitter_C=0.3, seed=6, shuffle_seed=2)
    sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, layout="aos")
    Po = _copy(P)
    for _ in range(3):
        sol.substep()
        oracle.substep(Po, P["dx"], synth.DT, E, NU, P["volume"], synth.GRAVITY, 1)
    check_particles(sol.particles_host(), Po, P["dx"], "aos multistep (3 substeps)", rtol=3e-5)


@pytest.mark.parametrize("name", ["mpm_cube6_mode0", "mpm_cube6_mode1", "mpm_cube8_rest", "mpm_cube5_neg"])
@pytest.mark.parametrize("layout", ["aos", "binned"])
def test_against_reference_golden(name, layout, binned_variant):
    """golden vectors produced by executing the reference (tests/golden/make_golden.py); compared by block key."""
    from zpc_b200 import api
    z = np.load(os.path.join(G, name + ".npz"))
    kw = dict(ast.literal_eval(str(z["kw"])))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **kw)
    n, dx, mode = P["x"].shape[0], P["dx"], int(z["mode"])
    pars, table = build_partition(P)
    ht = host_table(table)
    kr, g1r = grid_by_key(z["active_keys"], z["grid_p2g"])
    _, g2r = grid_by_key(z["active_keys"], z["grid_upd"])
    assert np.array_equal(ht["active_keys"], kr)              # ours are already in key order
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    model = api.model_fcr(P["volume"], E, NU)
    perm = np.arange(n)
    src = pars
    if layout == "binned":
        src = api.ParticleBins(n, max(ht["nblocks"] * 2, 64))
        order = torch.empty(n, dtype=torch.int32, device="cuda")
        api.bin_particles(pars, table, dx, src, order)
        perm = order.cpu().numpy()
    api.p2g_transfer(src, table, grids, synth.DT, model)
    check_channels(grids.tiles.cpu().numpy(), g1r, 1, "golden p2g", GRID_RTOL, strict_frac=0.99)
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), mode, mx)
    check_channels(grids.tiles.cpu().numpy()[:, 1:4], g2r[:, 1:4], 1, "golden grid v")
    assert abs(mx.item() - float(z["max_vel_sqr"])) <= 1e-5 * float(z["max_vel_sqr"])
    api.g2p_transfer(src, table, grids, synth.DT)
    out = {k: src.attr(k).cpu().numpy() for k in "xvCF"} if layout == "binned" else pars.to_host()
    check_particles(out, {k: z[k][perm] for k in "xvCF"}, dx, "golden g2p")


def test_full_size_properties_c2():
    """8 M particles (BASELINE configs[1]): size-independent checks — mass and momentum conservation through
    P2G, binned == AoS grid, G2P of a uniform velocity field returns it, run-to-run determinism of the binned P2G."""
    from zpc_b200 import api
    P = synth.config("C2")
    n, dx = P["x"].shape[0], P["dx"]
    pars = api.Particles(P)
    table = api.HashTable(max(n // 256, 1024))
    api.partition_for_particles(api.vec3_port(pars.x), n, dx, table)
    nb = table.size()
    assert table.overflow.item() == 0 and 15625 <= nb <= 20000
    grids = api.Grids(dx, nb)
    model = api.model_fcr(P["volume"], E, NU)
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(pars, table, grids, synth.DT, model)
    g_aos = grids.tiles.clone()
    M = float(pars.m.double().sum())
    assert abs(float(g_aos[:, 0].double().sum()) / M - 1) < 1e-5
    mom = g_aos[:, 1:4].double().sum(dim=(0, 2))
    assert abs(float(mom[1]) / (-M) - 1) < 1e-5 and abs(float(mom[0])) < 1e-5 * M and abs(float(mom[2])) < 1e-5 * M
    bins = api.ParticleBins(n, 2 * nb)
    api.bin_particles(pars, table, dx, bins)
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(bins, table, grids, synth.DT, model)
    g_bin = grids.tiles.clone()
    scale = g_aos.abs().amax(dim=(0, 2), keepdim=True).clamp_min(1e-30)
    assert float(((g_bin - g_aos).abs() / scale).max()) < 1e-5
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(bins, table, grids, synth.DT, model)
    # per-CTA sums are deterministic; only the order of the <= 8 bulk reductions per tile varies
    assert float(((grids.tiles - g_bin).abs() / scale).max()) < 1e-6
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, 0.0, 0.0), 1, mx)
    assert abs(mx.item() - 1.0) < 1e-4                               # |v|^2 of the uniform (0,-1,0) field
    api.g2p_transfer(bins, table, grids, synth.DT)
    v = bins.attr("v")
    assert float((v - torch.tensor([0.0, -1.0, 0.0], device="cuda")).abs().max()) < 1e-4
    assert float(bins.attr("C").abs().max()) < 1e-1 / dx * 1e-3      # affine part of a uniform field vanishes


def test_multi_gpu_substeps_match_single_gpu():
    """2 ranks over NCCL (needs >= 2 GPUs; the 1-GPU driver tier skips it — run with gpurun --gpus 2)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tests", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_eos_fluid_model_matches_oracle_and_golden(oracle):
    """EquationOfStateConfig (P2G.hpp:66-87, G2P.hpp:69-73) on the AoS drop-in path: vs oracle on the GPU-built table
    and vs the reference-generated golden vectors by block key."""
    from zpc_b200 import api
    z = np.load(os.path.join(G, "mpm_cube6_eos.npz"))
    kw = dict(ast.literal_eval(str(z["kw"])))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **kw)
    P["J"] = z["J_in"].copy()
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    model = api.model_eos(P["volume"], 4.0e4, 7.15, 0.01)
    api.p2g_transfer(pars, table, grids, synth.DT, model)
    g1 = grids.tiles.cpu().numpy()
    o1 = oracle.p2g_eos(P, ht, dx, synth.DT, 4.0e4, 0.01, P["volume"])
    check_channels(g1, o1, 1, "eos p2g", RTOL, strict_frac=0.99)          # no SVD on this path: 1e-5 on all 7 channels
    kr, g1r = grid_by_key(z["active_keys"], z["grid_p2g"])
    assert np.array_equal(ht["active_keys"], kr)
    check_channels(g1, g1r, 1, "eos golden p2g", RTOL)
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    F0 = pars.F.clone()
    api.g2p_transfer(pars, table, grids, synth.DT, model=model)
    assert abs(mx.item() - float(z["max_vel_sqr"])) <= 1e-5 * float(z["max_vel_sqr"])
    out = pars.to_host()
    vmax = float(np.abs(z["v"]).max())
    check_channels(out["x"], z["x"], 1, "eos x", RTOL, floor=float(np.abs(z["x"]).max()))
    check_channels(out["v"], z["v"], 1, "eos v", RTOL, floor=vmax)
    check_channels(out["C"], z["C"], 1, "eos C", RTOL, floor=4.0 / dx * vmax)
    assert np.abs(pars.J.cpu().numpy() - z["J"]).max() <= 1e-5
    assert torch.equal(pars.F, F0)                                        # F is not touched by the fluid model


def test_boundary_conditions_match_oracle_and_golden(oracle):
    """ApplyBoundaryConditionOnGridBlocks (GridOp.hpp:112-164): static plane / sphere x sticky / slip / separate.
    The inside/outside decision is exact arithmetic on node positions, so the SAME cells must change; values 1e-5."""
    from tests.golden.make_golden import BOUNDARY_CASES
    from zpc_b200 import api
    z = np.load(os.path.join(G, "mpm_cube7_boundary.npz"))
    kw = dict(ast.literal_eval(str(z["kw"])))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **kw)
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    kr, _ = grid_by_key(z["active_keys"], z["grid_0"])
    assert np.array_equal(ht["active_keys"], kr)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(pars, table, grids, synth.DT, api.model_fcr(P["volume"], E, NU))
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    base = grids.tiles.clone()
    for i, (geom, ctype, p0, p1) in enumerate(BOUNDARY_CASES):
        grids.tiles.copy_(base)
        col = api.plane_collider(p0, p1, ctype) if geom == 0 else api.sphere_collider(p0, p1[0], ctype)
        api.apply_boundary_condition(col, table, grids)
        got = grids.tiles.cpu().numpy()
        want = base.cpu().numpy()
        oracle.apply_boundary(want, ht["active_keys"], dx, geom, ctype, p0, p1)      # oracle on the GPU's own pre-state
        assert np.array_equal(got[:, [0, 4, 5, 6]], want[:, [0, 4, 5, 6]])          # m and rhs untouched
        check_channels(got[:, 1:4], want[:, 1:4], 1, "boundary %d/%d" % (geom, ctype))
        assert np.array_equal((got != base.cpu().numpy()).any(axis=1), (want != base.cpu().numpy()).any(axis=1))
        _, gold = grid_by_key(z["active_keys"], z["grid_%d" % i])
        check_channels(got[:, 1:4], gold[:, 1:4], 1, "boundary golden %d/%d" % (geom, ctype))


# ---- edge cases (ragged / empty / degenerate inputs) ---------------------------------------------------------------
def _single_particle(pos=(0.3, 0.3, 0.3), G=32):
    dx = 1.0 / G
    vol = dx ** 3 / 8
    return dict(x=np.array([pos], np.float32), v=np.array([[0.5, -1.0, 0.25]], np.float32), m=np.array([1000 * vol], np.float32),
                C=np.zeros((1, 9), np.float32), F=np.eye(3, dtype=np.float32).reshape(1, 9).copy(), dx=dx, volume=vol)


def test_empty_particle_set_is_a_noop():
    from zpc_b200 import api
    P = _single_particle()
    for k in ("x", "v", "m", "C", "F"):
        P[k] = P[k][:0].copy()
    pars = api.Particles(P)
    table = api.HashTable(64)
    api.partition_for_particles(api.zpc_port(pars.x.data_ptr(), 0, 0, 0, 3), 0, P["dx"], table)
    assert table.size() == 0
    grids = api.Grids(P["dx"], 8)
    grids.tiles.fill_(7.0)
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(pars, table, grids, synth.DT, api.model_fcr(P["volume"]))
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    api.g2p_transfer(pars, table, grids, synth.DT)
    bins = api.ParticleBins(0, 16)
    api.bin_particles(pars, table, P["dx"], bins)
    api.p2g_transfer(bins, table, grids, synth.DT, api.model_fcr(P["volume"]))
    api.g2p_transfer(bins, table, grids, synth.DT)
    torch.cuda.synchronize()
    assert bins.num_bins.item() == 0 and mx.item() == 0.0 and float(grids.tiles.min()) == 7.0   # nothing was touched


@pytest.mark.parametrize("pos", [(0.3, 0.3, 0.3), (-0.41, 0.02, -0.77), (0.5 - 1e-7, 0.25, 0.125)])
def test_single_particle_both_layouts(oracle, pos, binned_variant):
    """one particle: 8 blocks, 27 nodes; negative coordinates and a position sitting on a cell boundary"""
    from zpc_b200 import api
    P = _single_particle(pos)
    dx = P["dx"]
    pars, table = build_partition(P, expected=64)
    ht = host_table(table)
    tab_o = oracle.partition_build(P["x"], dx, oracle.table_size_for(64))
    assert ht["nblocks"] == tab_o["nblocks"] == 8
    model = api.model_fcr(P["volume"], E, NU)
    o1, o2, omx, Po = run_oracle_on_table(oracle, P, ht, 1)
    for layout in ("aos", "binned"):
        src = pars if layout == "aos" else api.ParticleBins(1, 64)
        if layout == "binned":
            api.bin_particles(api.Particles(P), table, dx, src)
        else:
            src = api.Particles(P)
        grids = api.Grids(dx, 8)
        api.clean_grid_blocks(grids, table)
        api.p2g_transfer(src, table, grids, synth.DT, model)
        g = grids.tiles.cpu().numpy()
        check_channels(g, o1, 1, layout + " single p2g", GRID_RTOL)
        assert (g[:, 0] > 0).sum() <= 27 and abs(g[:, 0].sum() / P["m"][0] - 1) < 1e-6
        mx = torch.zeros(1, device="cuda")
        api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
        api.g2p_transfer(src, table, grids, synth.DT)
        out = {k: src.attr(k).cpu().numpy() for k in "xvCF"} if layout == "binned" else src.to_host()
        check_particles(out, Po, dx, layout + " single g2p")


def test_dense_cluster_splits_bins_and_groups(oracle, binned_variant):
    """3 000 particles inside ONE cell: the home block exceeds ZPCB200_BIN_MAX (3 bins), every bin has a single
    1 000-particle cell group spanning several record chunks; compared with the oracle."""
    from zpc_b200 import api
    G_, n = 32, 3000
    dx = 1.0 / G_
    rs = np.random.RandomState(5)
    vol = dx ** 3 / 8
    P = dict(x=((np.array([9.0, 9.0, 9.0]) + rs.uniform(0.02, 0.98, (n, 3))) * dx).astype(np.float32),
             v=rs.uniform(-1, 1, (n, 3)).astype(np.float32), m=np.full(n, 1000 * vol / 300, np.float32),
             C=rs.uniform(-0.5, 0.5, (n, 9)).astype(np.float32),
             F=(np.eye(3).reshape(1, 9) + rs.uniform(-0.03, 0.03, (n, 9))).astype(np.float32), dx=dx, volume=vol / 300)
    pars, table = build_partition(P, expected=64)
    ht = host_table(table)
    bins = api.ParticleBins(n, 64)
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    api.bin_particles(pars, table, dx, bins, order)
    perm = order.cpu().numpy()
    nb = bins.num_bins.item()
    bs = bins.bin_start[: nb + 1].cpu().numpy()
    assert nb >= 3 and (np.diff(bs) <= api.BIN_MAX).all() and bs[-1] == n
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    model = api.model_fcr(P["volume"], E, NU)
    api.p2g_transfer(bins, table, grids, synth.DT, model)
    o1, o2, omx, Po = run_oracle_on_table(oracle, P, ht, 1)
    check_channels(grids.tiles.cpu().numpy(), o1, 1, "cluster p2g", GRID_RTOL)
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    api.g2p_transfer(bins, table, grids, synth.DT)
    check_particles({k: bins.attr(k).cpu().numpy() for k in "xvCF"}, {k: Po[k][perm] for k in "xvCF"}, dx, "cluster g2p")
    # second substep on the cached cell order (all particles still in one bin group each)
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(bins, table, grids, synth.DT, model)
    P2 = {k: (np.ascontiguousarray(Po[k]) if isinstance(Po[k], np.ndarray) else Po[k]) for k in Po}
    o1b = oracle.p2g(P2, ht, dx, synth.DT, E, NU, P["volume"])
    check_channels(grids.tiles.cpu().numpy(), o1b, 1, "cluster p2g (cached order)", [3e-5] * 4 + [1e-4] * 3)


def test_overflow_flags_are_raised_not_thrown():
    """too small a table / too few bins: device flags, no exception, no out-of-bounds write (reference convention)"""
    from zpc_b200 import api
    P = make("cube8")
    n, dx = P["x"].shape[0], P["dx"]
    pars = api.Particles(P)
    table = api.HashTable(2)                      # 32 slots, list capacity 64 < 125 blocks... force the flag
    api.partition_for_particles(api.vec3_port(pars.x), n, dx, table)
    torch.cuda.synchronize()
    assert table.overflow.item() == 1
    # domain beyond the packed block-code range (|block| >= 512)
    far = dict(P)
    far["x"] = (P["x"] + np.float32(100.0)).astype(np.float32)
    t2 = api.HashTable(max(n // 8, 64))
    api.partition_for_particles(api.vec3_port(api.Particles(far).x), n, dx, t2)
    torch.cuda.synchronize()
    assert t2.overflow.item() == 1


def test_vonmises_model_matches_oracle_and_golden(oracle):
    """VonMisesFixedCorotatedConfig (P2G.hpp:89-90, ConstitutiveModel_Vol_dP.hpp:49-110) on the AoS drop-in path: 39 % of
    the particles yield (radial return + projected F), none within 1e-3 of the yield surface.  vs the oracle on the
    GPU-built table and vs reference-generated golden vectors by block key."""
    from zpc_b200 import api
    z = np.load(os.path.join(G, "mpm_cube6_vonmises.npz"))
    kw = dict(ast.literal_eval(str(z["kw"])))
    ys = float(z["ys"])
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **kw)
    n, dx = P["x"].shape[0], P["dx"]
    assert 0.3 < float(z["yielded_fraction"]) < 0.5
    pars, table = build_partition(P)
    ht = host_table(table)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    model = api.model_vonmises(P["volume"], E, NU, ys)
    api.p2g_transfer(pars, table, grids, synth.DT, model)
    g1 = grids.tiles.cpu().numpy()
    o1 = oracle.p2g_vonmises(P, ht, dx, synth.DT, E, NU, ys, P["volume"])
    check_channels(g1, o1, 1, "vonmises p2g", GRID_RTOL, strict_frac=0.99)
    # the plastic projection changes the stress channels by far more than the tolerance
    fcr = oracle.p2g(P, ht, dx, synth.DT, E, NU, P["volume"])
    assert np.abs(fcr[:, 4:7] - o1[:, 4:7]).max() > 1e-2 * np.abs(o1[:, 4:7]).max()
    kr, g1r = grid_by_key(z["active_keys"], z["grid_p2g"])
    assert np.array_equal(ht["active_keys"], kr)
    check_channels(g1, g1r, 1, "vonmises golden p2g", GRID_RTOL)
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    # mode 1 folds rhs (held to RTOL_STRESS, tests/parity.py) into v = (mv + rhs)/m: the maximum sits on a light
    # corner node where rhs/m dominates, so max |v|^2 inherits the stress tolerance
    assert abs(mx.item() - float(z["max_vel_sqr"])) <= RTOL_STRESS * float(z["max_vel_sqr"])
    api.g2p_transfer(pars, table, grids, synth.DT)
    check_particles(pars.to_host(), {k: z[k] for k in "xvCF"}, dx, "vonmises golden g2p", rtol=1e-5)


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_grid_update_with_boundaries_equals_the_sequence(mode):
    """zpcb200_grid_update_bc == ComputeGridBlockVelocity, then ApplyBoundaryConditionOnGridBlocks per collider, bit for bit
    (both functor sequences are parity-tested against the oracle / golden vectors above)"""
    from zpc_b200 import api
    P = synth.elastic_cube(10, 32, jitter_F=0.03, jitter_C=0.4, seed=9)
    dx = P["dx"]
    pars, table = build_partition(P)
    nb = table.size()
    grids = api.Grids(dx, nb)
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(pars, table, grids, synth.DT, api.model_fcr(P["volume"], E, NU))
    after_p2g = grids.tiles.clone()
    cols = [api.plane_collider((0.0, 0.30, 0.0), (0.0, 1.0, 0.0), api.COLLIDER_SEPARATE),
            api.sphere_collider((0.33, 0.36, 0.33), 0.07, api.COLLIDER_SLIP),
            api.plane_collider((0.27, 0.0, 0.0), (1.0, 0.0, 0.0), api.COLLIDER_STICKY)]
    ext = (0.0, synth.GRAVITY, 0.0)
    mx_a, mx_b = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, ext, mode, mx_a)
    for c in cols:
        api.apply_boundary_condition(c, table, grids)
    want = grids.tiles.clone()
    grids.tiles.copy_(after_p2g)
    api.compute_grid_block_velocity_with_boundaries(grids, table, synth.DT, ext, mode, cols, mx_b)
    assert torch.equal(grids.tiles, want)
    assert mx_a.item() == mx_b.item() and mx_a.item() > 0
    moved = (want[:, 1:4] != 0).any()
    assert bool(moved)
    # no colliders: identical to the plain update
    grids.tiles.copy_(after_p2g)
    mx_c = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity_with_boundaries(grids, table, synth.DT, ext, mode, [], mx_c)
    grids2 = api.Grids(dx, nb)
    grids2.tiles.copy_(after_p2g)
    api.compute_grid_block_velocity(grids2, table, synth.DT, ext, mode, mx_a.zero_())
    assert torch.equal(grids.tiles, grids2.tiles)


def test_moving_colliders_match_oracle_and_golden(oracle):
    """Collider with translation / rotation / angular velocity / scaling (geometry/Collider.h:16-24,98-127): the same
    cells change as in the oracle (the inside test is evaluated without contraction), values within 1e-5; also through
    the fused update+boundary entry"""
    from tests.parity import MOVING_COLLIDERS, motion_vec
    from zpc_b200 import api
    z = np.load(os.path.join(G, "mpm_cube7_boundary_moving.npz"))
    kw = dict(ast.literal_eval(str(z["kw"])))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **kw)
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    kr, _ = grid_by_key(z["active_keys"], z["grid_0"])
    assert np.array_equal(ht["active_keys"], kr)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(pars, table, grids, synth.DT, api.model_fcr(P["volume"], E, NU))
    after_p2g = grids.tiles.clone()
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    base = grids.tiles.clone()
    for i, (geom, ctype, p0, p1, motion) in enumerate(MOVING_COLLIDERS):
        b, dbdt, R, om, s, dsdt = motion
        kwm = dict(translation=b, velocity=dbdt, rotation=np.asarray(R).tolist(), omega=om, scale=s, dscale_dt=dsdt)
        col = api.plane_collider(p0, p1, ctype, **kwm) if geom == 0 else api.sphere_collider(p0, p1[0], ctype, **kwm)
        grids.tiles.copy_(base)
        api.apply_boundary_condition(col, table, grids)
        got = grids.tiles.cpu().numpy()
        want = base.cpu().numpy()
        oracle.apply_boundary(want, ht["active_keys"], dx, geom, ctype, p0, p1, motion_vec(motion))
        assert np.array_equal(got[:, [0, 4, 5, 6]], want[:, [0, 4, 5, 6]])
        vscale = float(np.abs(want[:, 1:4]).max())
        check_channels(got[:, 1:4], want[:, 1:4], 1, "moving boundary %d/%d" % (geom, ctype), floor=vscale)
        changed_got = (got != base.cpu().numpy()).any(axis=1)
        changed_want = (want != base.cpu().numpy()).any(axis=1)
        assert (changed_got != changed_want).mean() < 1e-4         # a node within rounding of the surface may flip
        _, gold = grid_by_key(z["active_keys"], z["grid_%d" % i])
        check_channels(got[:, 1:4], gold[:, 1:4], 1, "moving boundary golden %d/%d" % (geom, ctype), floor=vscale)
        # fused with the grid update
        grids.tiles.copy_(after_p2g)
        mx2 = torch.zeros(1, device="cuda")
        api.compute_grid_block_velocity_with_boundaries(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, [col], mx2)
        assert torch.equal(grids.tiles, torch.from_numpy(got).cuda())
