"""GPU parity tests of the callers and data formats either side of the hot path: the plastic / fluid constitutive models on
the AoS, binned and SparseGrid paths, cuboid and moving colliders, the fused boundary update, index buckets, LBVH, G2P2G, the grid
momentum functors, the overlay on the reference's own containers, CUDA-graph replay.  First green run on a B200: round 2, call 1
(profiles/r02_pending_tests_first_run.log); every test is strict since.
"""
import ast
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu,
              pytest.mark.timeout(600, method="thread")]     # a kernel that never returns ends the run instead of holding the box

from zpc_b200 import synth  # noqa: E402
from tests.golden.make_golden import NACC, SAND  # noqa: E402
from tests.parity import RTOL, RTOL_STRESS, check_channels, check_particles, grid_by_key  # noqa: E402
from tests.test_gpu_mpm import build_partition, host_table  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
E, NU = synth.MODEL["E"], synth.MODEL["nu"]
# rhs tolerance per model = the reference's own build-to-build distance on this very case (oracle with / without FMA
# contraction: sand 1.5e-5, NACC 2.4e-4 of channel scale) with the margin tests/parity.py uses for fixed-corotated
RHS_RTOL = {"sand": RTOL_STRESS, "nacc": 1e-3}


@pytest.mark.parametrize("model", ["sand", "nacc"])
def test_plastic_model_matches_oracle_and_golden(oracle, model):
    """DruckerPragerConfig / NACCConfig (P2G.hpp:92-102) on the AoS drop-in path: grid after P2G and the logJp written back,
    vs the oracle on the GPU-built table and vs reference-generated golden vectors by block key; then update + G2P."""
    from zpc_b200 import api
    z = np.load(os.path.join(G, "mpm_cube6_%s.npz" % model))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    P["logJp"] = z["logJp_in"].copy()
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    assert pars.logJp is not None
    ht = host_table(table)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    if model == "sand":
        m = api.model_drucker_prager(P["volume"], E, NU, SAND["cohesion"], SAND["beta"], SAND["volumeCorrection"], SAND["yieldSurface"])
    else:
        m = api.model_nacc(P["volume"], NACC["E"], NACC["nu"], NACC["fa"], NACC["xi"], NACC["beta"], NACC["hardeningOn"])
    api.p2g_transfer(pars, table, grids, synth.DT, m)
    torch.cuda.synchronize()
    g1 = grids.tiles.cpu().numpy()
    lj = pars.logJp.cpu().numpy()
    Po = dict(P, logJp=P["logJp"].copy())
    if model == "sand":
        o1 = oracle.p2g_sand(Po, ht, dx, synth.DT, E, NU, SAND, P["volume"])
    else:
        o1 = oracle.p2g_nacc(Po, ht, dx, synth.DT, NACC["E"], NACC["nu"], NACC, P["volume"])
    rtol = [RTOL] * 4 + [RHS_RTOL[model]] * 3
    check_channels(g1, o1, 1, model + " p2g", rtol)
    assert np.abs(lj - Po["logJp"]).max() <= 2e-5            # logf / expf / powf of the device vs glibc: a few ulp
    assert np.abs(lj - P["logJp"]).max() > 1e-3               # and P2G really wrote it back
    kr, g1r = grid_by_key(z["active_keys"], z["grid_p2g"])
    assert np.array_equal(ht["active_keys"], kr)
    check_channels(g1, g1r, 1, model + " golden p2g", rtol)
    assert np.abs(lj - z["logJp"]).max() <= 2e-5
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    assert abs(mx.item() - float(z["max_vel_sqr"])) <= RHS_RTOL[model] * float(z["max_vel_sqr"])
    api.g2p_transfer(pars, table, grids, synth.DT)
    check_particles(pars.to_host(), {k: z[k] for k in "xvCF"}, dx, model + " golden g2p", rtol=1e-5)


def test_plastic_models_need_logjp():
    from zpc_b200 import api
    P = synth.elastic_cube(4, 16)
    pars, table = build_partition(P)
    grids = api.Grids(P["dx"], table.size())
    with pytest.raises(ValueError):
        api.p2g_transfer(pars, table, grids, synth.DT, api.model_nacc(P["volume"]))


def test_cuboid_colliders_match_oracle_and_golden(oracle):
    """Collider over AnalyticLevelSet<Cuboid> (AnalyticLevelSet.h:55-126), static and moving, sticky / slip / separate: the
    same cells change as in the oracle (distance and finite-difference normal are evaluated without contraction, i.e.
    bit-identically), velocities within 1e-5; also through the fused update + boundary entry"""
    from tests.parity import CUBOID_COLLIDERS, motion_vec
    from zpc_b200 import api
    z = np.load(os.path.join(G, "mpm_cube7_boundary_cuboid.npz"))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    dx = P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(pars, table, grids, synth.DT, api.model_fcr(P["volume"], E, NU))
    after_p2g = grids.tiles.clone()
    mx = torch.zeros(1, device="cuda")
    ext = (0.0, synth.GRAVITY, 0.0)
    api.compute_grid_block_velocity(grids, table, synth.DT, ext, 1, mx)
    base = grids.tiles.clone()
    for i, (geom, ctype, p0, p1, motion) in enumerate(CUBOID_COLLIDERS):
        kwm = {}
        if motion is not None:
            b, dbdt, R, om, s, dsdt = motion
            kwm = dict(translation=b, velocity=dbdt, rotation=np.asarray(R).tolist(), omega=om, scale=s, dscale_dt=dsdt)
        col = api.cuboid_collider(p0, p1, ctype, **kwm)
        grids.tiles.copy_(base)
        api.apply_boundary_condition(col, table, grids)
        got = grids.tiles.cpu().numpy()
        want = base.cpu().numpy()
        oracle.apply_boundary(want, ht["active_keys"], dx, geom, ctype, p0, p1, motion_vec(motion))
        assert np.array_equal(got[:, [0, 4, 5, 6]], want[:, [0, 4, 5, 6]])
        vscale = float(np.abs(want[:, 1:4]).max())
        check_channels(got[:, 1:4], want[:, 1:4], 1, "cuboid %d/%d" % (i, ctype), floor=vscale)
        changed_got = (got != base.cpu().numpy()).any(axis=1)
        changed_want = (want != base.cpu().numpy()).any(axis=1)
        assert changed_want.sum() > 20 and (changed_got != changed_want).mean() < 1e-4
        _, gold = grid_by_key(z["active_keys"], z["grid_%d" % i])
        check_channels(got[:, 1:4], gold[:, 1:4], 1, "cuboid golden %d/%d" % (i, ctype), floor=vscale)
        grids.tiles.copy_(after_p2g)      # fused with the grid update
        api.compute_grid_block_velocity_with_boundaries(grids, table, synth.DT, ext, 1, [col], mx.zero_())
        assert torch.equal(grids.tiles, torch.as_tensor(got, device="cuda"))


@pytest.mark.parametrize("sweep", [4, 3])
def test_vonmises_on_the_binned_path_matches_oracle_and_golden(oracle, sweep):
    """VonMisesFixedCorotatedConfig through the block-binned P2G (zpcb200_p2g_apic_vonmises_binned: the model is a template
    parameter of the record phase): vs the oracle, the reference-generated golden grid and the AoS kernel"""
    from zpc_b200 import api
    from tests.parity import GRID_RTOL
    z = np.load(os.path.join(G, "mpm_cube6_vonmises.npz"))
    ys = float(z["ys"])
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    bins = api.ParticleBins(n, max(ht["nblocks"] * 2, 64))
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    api.bin_particles(pars, table, dx, bins, order)
    model = api.model_vonmises(P["volume"], E, NU, ys)
    api.set_tuning(sweep, -1)
    try:
        grids = api.Grids(dx, ht["nblocks"])
        api.clean_grid_blocks(grids, table)
        api.p2g_transfer(bins, table, grids, synth.DT, model)
        torch.cuda.synchronize()
    finally:
        api.set_tuning(4, -1)
    g1 = grids.tiles.cpu().numpy()
    o1 = oracle.p2g_vonmises(P, ht, dx, synth.DT, E, NU, ys, P["volume"])
    check_channels(g1, o1, 1, "binned vonmises p2g", GRID_RTOL, strict_frac=0.99)
    fcr = oracle.p2g(P, ht, dx, synth.DT, E, NU, P["volume"])
    assert np.abs(fcr[:, 4:7] - g1[:, 4:7]).max() > 1e-2 * np.abs(o1[:, 4:7]).max()      # it is not the elastic stress
    _, gold = grid_by_key(z["active_keys"], z["grid_p2g"])
    check_channels(g1, gold, 1, "binned vonmises golden p2g", GRID_RTOL)
    grids2 = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids2, table)
    api.p2g_transfer(pars, table, grids2, synth.DT, model)
    check_channels(g1, grids2.tiles.cpu().numpy(), 1, "binned vs AoS vonmises", GRID_RTOL)


@pytest.mark.parametrize("layout", ["aos", "binned"])
def test_solver_with_model_and_colliders_matches_oracle(oracle, layout):
    """MpmSolver(model=von Mises, colliders=[separating floor, slipping box]) for three substeps vs the oracle composing
    the reference's functor sequence: partition, P2G (von Mises), grid update, boundary per collider, G2P"""
    from zpc_b200 import api
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(8, 32, jitter_F=0.04, jitter_C=0.4, seed=12)
    P["v"][:] = P["v"] * 4.0
    n, dx, dt, ys = P["x"].shape[0], P["dx"], synth.DT * 10, 2946.0
    P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n) / n)).astype(np.float32)     # unique masses = particle identity
    floor = (0, 2, (0.0, 0.25, 0.0), (0.0, 1.0, 0.0))
    box = (2, 1, (0.2, 0.2, 0.2), (0.3, 0.3, 0.45))
    cols = [api.plane_collider(floor[2], floor[3], floor[1]), api.cuboid_collider(box[2], box[3], box[1])]
    sol = MpmSolver(P, dx, P["volume"], dt, synth.GRAVITY, mode=1, layout=layout, rebin_every=2,
                    model=api.model_vonmises(P["volume"], E, NU, ys), colliders=cols)
    Po = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in P.items()}
    for _ in range(3):
        sol.substep()
        tab = oracle.partition_build(Po["x"], dx, oracle.table_size_for(max(n // 8, 1)))
        g = oracle.p2g_vonmises(Po, tab, dx, dt, E, NU, ys, P["volume"])
        oracle.grid_update(g, dt, (0.0, synth.GRAVITY, 0.0), 1)
        for geom, ctype, p0, p1 in (floor, box):
            oracle.apply_boundary(g, tab["active_keys"], dx, geom, ctype, p0, p1)
        oracle.g2p(Po, tab, g, dx, dt)
    torch.cuda.synchronize()
    got = sol.particles_host()

    def canon(Q):
        o = np.argsort(Q["m"], kind="stable")
        return {k: Q[k][o] for k in "xvCF"}
    check_particles(canon(got), canon(Po), dx, "solver %s, von Mises + colliders" % layout, rtol=5e-5)
    assert (Po["x"][:, 1].min() < 0.25 + dx)          # the cloud reaches the floor region


@pytest.mark.parametrize("chunks", [1, 5])
def test_pipelined_host_call_equals_the_plain_one(chunks):
    """MpmSolver.substep_host_pipelined (PCIe copies overlapped with chunked P2G / G2P) returns what substep_host returns
    (the AoS scatter uses unordered float atomics: equal up to fp32 re-association), for two consecutive substeps"""
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(9, 32, jitter_F=0.04, jitter_C=0.4, shuffle_seed=6)
    res = []
    for pipelined in (False, True):
        sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, layout="aos")
        hin = {k: torch.from_numpy(P[k].copy()).pin_memory() for k in ("x", "v", "m", "C", "F")}
        hout = {k: torch.empty_like(hin[k]).pin_memory() for k in ("x", "v", "C", "F")}
        mx = []
        for _ in range(2):
            mx.append(sol.substep_host_pipelined(hin, hout, chunks) if pipelined else sol.substep_host(hin, hout))
            torch.cuda.synchronize()
            for k in ("x", "v", "C", "F"):
                hin[k], hout[k] = hout[k], hin[k]
        res.append(({k: hin[k].numpy().copy() for k in "xvCF"}, mx))
    check_particles(res[1][0], res[0][0], P["dx"], "pipelined host call", rtol=1e-5)
    assert all(abs(a - b) <= 1e-4 * b for a, b in zip(res[1][1], res[0][1]))
    assert not np.array_equal(res[0][0]["x"], P["x"])


@pytest.mark.parametrize("n,dup,scale", [(0, 0, 1.0), (1, 0, 1.0), (2, 0, 1.0), (3, 0, 1.0), (5, 1, 1.0), (1000, 1, 1.0), (4097, 0, 1.0),
                                         (3000, 0, 100.0), (300000, 0, 1.0), (300000, 1, 1.0)])
def test_lbvh_build_and_refit_match_oracle(oracle, n, dup, scale):
    """zpcb200_lbvh_build / _refit (reduce -> Morton -> radix_sort_pair -> Karras topology -> exclusive_scan -> DFS layout ->
    bottom-up refit) vs the oracle, which is pinned bit-exact against the reference's LBvh: every array identical"""
    from tests.test_oracle_lbvh import boxes
    from zpc_b200 import api
    rs = np.random.RandomState(n + dup)
    b = boxes(rs, n, dup, scale) if n else np.zeros((0, 6), np.float32)
    bvh = api.LBvh().build(torch.from_numpy(b).cuda())
    torch.cuda.synchronize()
    if n == 0:
        return
    A = oracle.lbvh_build(b)
    for k in ("auxIndices", "leafInds") + (("parents", "levels") if n > 2 else ()):
        assert np.array_equal(getattr(bvh, k).cpu().numpy()[: len(A[k])], A[k]), k
    assert np.array_equal(bvh.orderedBvs.cpu().numpy().view(np.uint32), A["orderedBvs"].view(np.uint32))
    b1 = (b + rs.uniform(-0.01, 0.01, (n, 1)).astype(np.float32)).astype(np.float32)
    bvh.refit(torch.from_numpy(b1).cuda())
    oracle.lbvh_refit(A, b1)
    assert np.array_equal(bvh.orderedBvs.cpu().numpy().view(np.uint32), A["orderedBvs"].view(np.uint32))
    with pytest.raises(RuntimeError):
        bvh.refit(torch.zeros(n + 1, 6, device="cuda"))


@pytest.mark.parametrize("disp", [0.5, 0.0])
@pytest.mark.parametrize("case", [dict(s=12, G=32, shuffle_seed=5), dict(s=6, G=16, origin_cells=-9, shuffle_seed=2), dict(s=1, G=8),
                                  dict(s=40, G=64, shuffle_seed=1)])
def test_index_buckets_match_oracle(oracle, case, disp):
    """zpcb200_index_buckets_build vs the oracle (pinned against the reference's functor sequence): same cell set, and — bucket by
    bucket through the cell key — the same particle ids in the same (ascending) order; the table resolves through the reference's
    query; offsets are the exclusive scan of counts"""
    from zpc_b200 import api
    kw = dict(case)
    P = synth.elastic_cube(kw.pop("s"), kw.pop("G"), **kw)
    x, n, dx = P["x"], P["x"].shape[0], P["dx"]
    xs = torch.from_numpy(x).cuda()
    ib = api.index_buckets_for_particles(api.vec3_port(xs), n, dx, disp)
    torch.cuda.synchronize()
    assert ib.table.overflow.item() == 0
    nb = ib.num_buckets()
    A = oracle.index_buckets(x, dx, disp, oracle.table_size_for(n))
    assert nb == A["nblocks"]
    keys = ib.table.active_keys[:nb].cpu().numpy()
    ko = A["active_keys"]
    order = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0]))
    assert np.array_equal(keys, ko[order])                                    # bucket number = rank of the cell key
    counts, offsets, ids = ib.counts.cpu().numpy(), ib.offsets.cpu().numpy(), ib.indices.cpu().numpy()
    assert np.array_equal(counts[:nb], A["counts"][order]) and (counts[nb:] == 0).all()
    assert np.array_equal(offsets[1:], np.cumsum(counts)[:-1]) and offsets[0] == 0 and offsets[nb] == n
    for j in range(0, nb, max(nb // 200, 1)):
        b = order[j]
        assert np.array_equal(ids[offsets[j]: offsets[j] + counts[j]], A["ids"][A["offsets"][b]: A["offsets"][b] + A["counts"][b]])
    ht = host_table(ib.table)
    got = np.array([oracle.table_query(k, ht) for k in keys[:: max(nb // 100, 1)]])
    assert np.array_equal(got, np.arange(nb)[:: max(nb // 100, 1)])


@pytest.mark.parametrize("n", [1, 2, 3, 5000, 200000])
def test_lbvh_batched_query_matches_oracle(oracle, n):
    """zpcb200_lbvh_query (count -> exclusive_scan -> fill): per query the oracle's primitive ids in the oracle's visiting
    order, i.e. LBvhView::iter_neighbors; checked against brute force as well"""
    from tests.test_oracle_lbvh import boxes
    from zpc_b200 import api
    rs = np.random.RandomState(40 + n)
    b = boxes(rs, n, dup=n > 10)
    bvh = api.LBvh().build(torch.from_numpy(b).cuda())
    nq = 300
    qc, qh = rs.uniform(0, 1, (nq, 3)).astype(np.float32), rs.uniform(0.005, 0.05, (nq, 3)).astype(np.float32)
    qb = np.concatenate([qc - qh, qc + qh], 1).astype(np.float32)
    offsets, ids = bvh.query(torch.from_numpy(qb).cuda())
    torch.cuda.synchronize()
    offsets, ids = offsets.cpu().numpy(), ids.cpu().numpy()
    t = oracle.lbvh_build(b)
    for q in range(nq):
        want = oracle.lbvh_iter_neighbors(t, qb[q], cap=max(n, 1))
        assert np.array_equal(ids[offsets[q]: offsets[q + 1]], want), q
        brute = np.nonzero(~((qb[q][None, :3] > b[:, 3:]).any(1) | (qb[q][None, 3:] < b[:, :3]).any(1)))[0]
        assert np.array_equal(np.sort(want), brute)


def test_against_the_references_own_cuda_path(oracle, tmp_path):
    """"pick the CUDA reference as primary for GPU parity".  oracle/_ref/libzpcref_cuda.so is the unmodified reference
    (headers + CUDA backend) compiled for sm_100; it runs one substep in a process of its own (oracle/refcuda_runner.py).  Its
    grid after P2G / update and its particles after G2P are compared with this library's AoS path on the same input, by block
    key, and with the host oracle (the reference's device build differs from its host build by ::rsqrtf and FMA contraction:
    the measured basis of RTOL_STRESS)."""
    import subprocess
    import sys
    from oracle.refcuda_runner import RefCuda
    from tests.parity import GRID_RTOL
    from zpc_b200 import api
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libzpcref_cuda.so not built (make -C oracle refcuda, where /root/reference is mounted)")
    P = synth.elastic_cube(8, 32, jitter_F=0.05, jitter_C=0.5, shuffle_seed=11)
    n, dx = P["x"].shape[0], P["dx"]
    fin, fout = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(fin, dt=synth.DT, E=E, nu=NU, gravity=synth.GRAVITY, mode=1, **P)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "substep", fin, fout], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(fout)
    # ours, same input
    pars, table = build_partition(P)
    ht = host_table(table)
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(pars, table, grids, synth.DT, api.model_fcr(P["volume"], E, NU))
    g1 = grids.tiles.cpu().numpy()
    kr, g1r = grid_by_key(z["active_keys"], z["grid_p2g"])
    assert int(z["nblocks"]) == ht["nblocks"] and np.array_equal(ht["active_keys"], kr)       # same block set (ours is key-ordered)
    check_channels(g1, g1r, 1, "P2G vs reference CUDA", GRID_RTOL)
    mx = torch.zeros(1, device="cuda")
    api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
    _, g2r = grid_by_key(z["active_keys"], z["grid_upd"])
    check_channels(grids.tiles.cpu().numpy()[:, 1:4], g2r[:, 1:4], 1, "grid update vs reference CUDA", RTOL_STRESS)
    api.g2p_transfer(pars, table, grids, synth.DT)
    check_particles(pars.to_host(), {k: z[k] for k in "xvCF"}, dx, "G2P vs reference CUDA", rtol=1e-5)
    # and the host oracle against the reference's device build: the distance the parity rule allows for
    o1 = oracle.p2g(P, ht, dx, synth.DT, E, NU, P["volume"])
    check_channels(o1, g1r, 1, "host oracle vs reference CUDA", GRID_RTOL)


@pytest.mark.parametrize("model", ["fcr", "vonmises", "nacc", "eos"])
def test_g2p2g_matches_the_restated_functor(oracle, model):
    """zpcb200_g2p2g_apic vs oracle.zo_g2p2g (a restatement of G2P2G.hpp:49-141, itself pinned against the reference's own functor by
    the next test) for the models the reference driver does not instantiate as well: force terms added to gridr, particles untouched"""
    from zpc_b200 import api
    P = synth.elastic_cube(8, 32, jitter_F=0.04, jitter_C=0.3, shuffle_seed=21)
    n, dx = P["x"].shape[0], P["dx"]
    rs = np.random.RandomState(5)
    En, nun = (NACC["E"], NACC["nu"]) if model == "nacc" else (E, NU)
    if model == "nacc":
        P["logJp"] = rs.uniform(-1.8, 0.2, n).astype(np.float32)
    if model == "eos":
        P["J"] = (1.0 + rs.uniform(-0.05, 0.05, n)).astype(np.float32)
    pars, table = build_partition(P)
    ht = host_table(table)
    nb = ht["nblocks"]
    gridv = rs.uniform(-1, 1, (nb * 64, 3)).astype(np.float32)
    m, kind, prm = {"fcr": (api.model_fcr(P["volume"], E, NU), 0, [0]),
                    "vonmises": (api.model_vonmises(P["volume"], E, NU, 2946.0), 1, [2946.0]),
                    "nacc": (api.model_nacc(P["volume"], NACC["E"], NACC["nu"], NACC["fa"], NACC["xi"], NACC["beta"], True), 3,
                             [NACC["xi"], NACC["beta"], 1.0, NACC["fa"], 3.0]),
                    "eos": (api.model_eos(P["volume"], 4.0e4, 7.15, 0.01), 4, [4.0e4, 0.01])}[model]
    gv, gr = torch.from_numpy(gridv).cuda(), torch.zeros(nb * 64, 3, device="cuda")
    api.g2p2g_transfer(pars, table, dx, synth.DT, m, gv, gr)
    torch.cuda.synchronize()
    want = oracle.g2p2g(kind, prm, P, ht, dx, synth.DT, En, nun, P["volume"], gridv)
    scale = float(np.abs(want).max())
    assert np.abs(gr.cpu().numpy() - want).max() <= (1e-3 if model == "nacc" else RTOL_STRESS) * scale
    assert np.array_equal(pars.x.cpu().numpy(), P["x"])


@pytest.mark.parametrize("model", ["fcr", "vonmises"])
def test_g2p2g_matches_the_references_own_functor(oracle, model, tmp_path):
    """PINS G2P2G: the reference's own G2P2GTransfer (simulation/transfer/G2P2G.hpp:49-141), compiled by nvcc into
    oracle/_ref/libzpcref_cuda.so with a plain three-floats-per-node DOF view (oracle/ref_driver_cuda.cu: the reference's DofView
    does not compile, the functor only needs get / ref) and run on cuda_exec() in a process of its own, against
    zpcb200_g2p2g_apic AND the restated oracle zo_g2p2g on the same particles and the same grid velocity field, by block key."""
    import subprocess
    import sys
    from oracle.refcuda_runner import RefCuda, node_field
    from zpc_b200 import api
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libzpcref_cuda.so not built (make -C oracle refcuda, where /root/reference is mounted)")
    P = synth.elastic_cube(8, 32, jitter_F=0.04, jitter_C=0.3, shuffle_seed=23)
    dx = P["dx"]
    kind, prm, m = {"fcr": (0, [E, NU], api.model_fcr(P["volume"], E, NU)),
                    "vonmises": (1, [E, NU, 2946.0], api.model_vonmises(P["volume"], E, NU, 2946.0))}[model]
    fin, fout = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(fin, dt=synth.DT, model=kind, prm=np.array(prm, np.float32), **P)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "g2p2g", fin, fout], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(fout)
    kr, want = grid_by_key(z["active_keys"], z["gridr"])
    pars, table = build_partition(P)
    ht = host_table(table)
    nb = ht["nblocks"]
    assert int(z["nblocks"]) == nb and np.array_equal(ht["active_keys"], kr)
    gridv = node_field(ht["active_keys"])
    gv, gr = torch.from_numpy(gridv).cuda(), torch.zeros(nb * 64, 3, device="cuda")
    api.g2p2g_transfer(pars, table, dx, synth.DT, m, gv, gr)
    torch.cuda.synchronize()
    got = gr.cpu().numpy().reshape(nb, 64, 3)
    scale = float(np.abs(want).max())
    assert scale > 0
    err = float(np.abs(got - want).max()) / scale
    assert err <= RTOL_STRESS, "zpcb200_g2p2g_apic vs the reference's G2P2GTransfer: %g" % err
    # the restatement the CPU tests use, pinned the same way
    rest = oracle.g2p2g(kind, prm[2:] or [0], P, ht, dx, synth.DT, E, NU, P["volume"], gridv).reshape(nb, 64, 3)
    err_o = float(np.abs(rest - want).max()) / scale
    assert err_o <= RTOL_STRESS, "zo_g2p2g vs the reference's G2P2GTransfer: %g" % err_o
    print("g2p2g %s: max err / max |r|  ours %.2e  oracle %.2e" % (model, err, err_o))


def test_overlay_on_the_references_containers(oracle, tmp_path):
    """include/zpcb200/zs_overlay.cuh compiled against the unmodified reference headers (oracle/_ref/libzpcref_cuda.so): the
    reference's own Particles / HashTable / Grids on the device, the composed substep once through the reference's functors on
    cuda_exec() and once through b200_exec() + zs::b200::* — same block set, grids and particles within the parity rule; and
    zs::radix_sort_pair / exclusive_scan / reduce (generic code templated on the policy) with b200_exec(), index-exact."""
    import subprocess
    import sys
    from oracle.refcuda_runner import RefCuda
    from tests.parity import GRID_RTOL
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libzpcref_cuda.so not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    P = synth.elastic_cube(8, 32, jitter_F=0.05, jitter_C=0.5, shuffle_seed=11)
    fin = str(tmp_path / "in.npz")
    np.savez(fin, dt=synth.DT, E=E, nu=NU, gravity=synth.GRAVITY, mode=1, **P)
    outs = {}
    for mode in ("substep", "overlay"):
        fout = str(tmp_path / (mode + ".npz"))
        r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", mode, fin, fout], cwd=root, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, mode + ": " + r.stdout + r.stderr
        outs[mode] = np.load(fout)
    a, b = outs["substep"], outs["overlay"]
    ka, ga = grid_by_key(a["active_keys"], a["grid_p2g"])
    kb, gb = grid_by_key(b["active_keys"], b["grid_p2g"])
    assert np.array_equal(ka, kb) and np.array_equal(b["active_keys"], kb)          # the overlay's numbering is key-ordered
    check_channels(gb, ga, 1, "overlay P2G vs reference functors", GRID_RTOL)
    _, ua = grid_by_key(a["active_keys"], a["grid_upd"])
    _, ub = grid_by_key(b["active_keys"], b["grid_upd"])
    check_channels(ub[:, 1:4], ua[:, 1:4], 1, "overlay grid update", RTOL_STRESS)
    assert abs(float(a["max_vel_sqr"]) - float(b["max_vel_sqr"])) <= RTOL_STRESS * float(a["max_vel_sqr"])
    check_particles({k: b[k] for k in "xvCF"}, {k: a[k] for k in "xvCF"}, P["dx"], "overlay G2P vs reference functors", rtol=1e-5)
    fout = str(tmp_path / "prims.npz")
    r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "prims", "100003", fout], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(fout)
    ek, ev = oracle.radix_sort_pair("u32", z["keys"], z["vals"])
    assert np.array_equal(z["keys_out"], ek) and np.array_equal(z["vals_out"], ev)
    assert np.array_equal(z["scan"], oracle.scan("exclusive", "i32", z["vals"]))
    assert int(z["sum"][0]) == int(z["vals"].sum()) and int(z["max"][0]) == int(z["vals"].max())
    # a TileVector channel through the reference's aosoa iterators (the way its own C layer passes them)
    ntiles, nch, chn, na = (int(v) for v in z["tv_shape"])
    src = z["tv"].reshape(ntiles, nch, 32)[:, chn, :].reshape(-1)[:na]
    got = z["tv_scan"].reshape(ntiles, nch, 32)
    assert np.array_equal(got[:, chn, :].reshape(-1)[:na], np.cumsum(src) - src) and int(z["tv_sum"][0]) == int(src.sum())
    other = np.delete(got, chn, axis=1)
    assert not other.any() and not got[:, chn, :].reshape(-1)[na:].any()           # nothing outside the channel range was written


def test_references_own_lbvh_build_runs_on_b200_exec(oracle, tmp_path):
    """LBvh<3,int,f32>::build of the reference, UNCHANGED (it is a template in the policy), once with cuda_exec() and once with
    b200_exec(): its radix_sort_pair / exclusive_scan then run in libzpcb200, its own functors through the inherited launcher.  Both
    trees and this repository's native zpcb200_lbvh_build equal the oracle's, array for array."""
    import subprocess
    import sys
    from oracle.refcuda_runner import RefCuda
    from zpc_b200 import api
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libzpcref_cuda.so not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fout = str(tmp_path / "lbvh.npz")
    r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "lbvh", "20000", fout], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(fout)
    A = oracle.lbvh_build(z["bvs"])
    native = api.LBvh().build(torch.from_numpy(z["bvs"]).cuda())
    for k in ("auxIndices", "parents", "levels", "leafInds", "orderedBvs"):
        want = A[k].view(np.uint32) if A[k].dtype == np.float32 else A[k]
        for tag in ("cuda", "b200"):
            got = z[tag + "_" + k]
            assert np.array_equal(got.view(np.uint32) if got.dtype == np.float32 else got, want), (tag, k)
        got = getattr(native, k).cpu().numpy()
        assert np.array_equal(got.view(np.uint32) if got.dtype == np.float32 else got, want), ("native", k)


@pytest.mark.parametrize("model", ["sand", "nacc"])
def test_plastic_models_on_the_binned_path(oracle, model):
    """Drucker-Prager / NACC through the block-binned P2G (logJp as a side array in bin order): grid and logJp vs the
    reference-generated golden vectors; then four substeps with a re-bin in between (logJp follows the permutation) against the
    any-order AoS solver on the same particles."""
    from zpc_b200 import api
    from zpc_b200.solver import MpmSolver
    z = np.load(os.path.join(G, "mpm_cube6_%s.npz" % model))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    P["logJp"] = z["logJp_in"].copy()
    n, dx = P["x"].shape[0], P["dx"]
    if model == "sand":
        m = api.model_drucker_prager(P["volume"], E, NU, SAND["cohesion"], SAND["beta"], SAND["volumeCorrection"], SAND["yieldSurface"])
    else:
        m = api.model_nacc(P["volume"], NACC["E"], NACC["nu"], NACC["fa"], NACC["xi"], NACC["beta"], NACC["hardeningOn"])
    pars, table = build_partition(P)
    ht = host_table(table)
    bins = api.ParticleBins(n, max(ht["nblocks"] * 2, 64))
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    api.bin_particles(pars, table, dx, bins, order)
    bins.logJp = pars.logJp[order.long()].contiguous()
    grids = api.Grids(dx, ht["nblocks"])
    api.clean_grid_blocks(grids, table)
    api.p2g_transfer(bins, table, grids, synth.DT, m)
    torch.cuda.synchronize()
    rtol = [RTOL] * 4 + [RHS_RTOL[model]] * 3
    _, gold = grid_by_key(z["active_keys"], z["grid_p2g"])
    check_channels(grids.tiles.cpu().numpy(), gold, 1, "binned %s golden p2g" % model, rtol)
    perm = order.cpu().numpy()
    assert np.abs(bins.logJp.cpu().numpy() - z["logJp"][perm]).max() <= 2e-5
    # multi-step with a re-bin: binned vs AoS solver
    Q = dict(P)
    Q["v"] = (P["v"] * 6.0).astype(np.float32)
    Q["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n) / n)).astype(np.float32)
    res = []
    for layout in ("aos", "binned"):
        sol = MpmSolver(Q, dx, P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout=layout, rebin_every=2, model=m,
                        **({"partition": "with_rebin"} if layout == "binned" else {}))
        for _ in range(4):
            sol.substep()
        torch.cuda.synchronize()
        H = sol.particles_host() if layout == "binned" else dict(sol.aos.to_host(), logJp=sol.aos.logJp.cpu().numpy())
        o = np.argsort(H["m"], kind="stable")
        res.append({k: H[k][o] for k in ("x", "v", "C", "F", "logJp")})
    check_particles(res[1], res[0], dx, "binned vs AoS, %s, 4 substeps" % model, rtol=1e-4)
    assert np.abs(res[1]["logJp"] - res[0]["logJp"]).max() <= 1e-4


@pytest.mark.parametrize("staged", [1, 0])
def test_equation_of_state_on_the_binned_path(oracle, staged):
    """EquationOfStateConfig through the binned kernels (J as a side array in bin order; P2G record from C and J, G2P updates J and
    leaves F alone): vs the reference-generated golden vectors, then four substeps with a re-bin against the AoS solver"""
    from zpc_b200 import api
    from zpc_b200.solver import MpmSolver
    z = np.load(os.path.join(G, "mpm_cube6_eos.npz"))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    P["J"] = z["J_in"].copy()
    n, dx = P["x"].shape[0], P["dx"]
    m = api.model_eos(P["volume"], 4.0e4, 7.15, 0.01)
    api.set_tuning(-1, staged)
    try:
        pars, table = build_partition(P)
        ht = host_table(table)
        bins = api.ParticleBins(n, max(ht["nblocks"] * 2, 64))
        order = torch.empty(n, dtype=torch.int32, device="cuda")
        api.bin_particles(pars, table, dx, bins, order)
        bins.J = pars.J[order.long()].contiguous()
        grids = api.Grids(dx, ht["nblocks"])
        api.clean_grid_blocks(grids, table)
        api.p2g_transfer(bins, table, grids, synth.DT, m)
        _, gold = grid_by_key(z["active_keys"], z["grid_p2g"])
        check_channels(grids.tiles.cpu().numpy(), gold, 1, "binned EOS golden p2g", RTOL)
        mx = torch.zeros(1, device="cuda")
        api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
        api.g2p_transfer(bins, table, grids, synth.DT, model=m)
        torch.cuda.synchronize()
        perm = order.cpu().numpy()
        for k in "xv":
            check_channels(bins.attr(k).cpu().numpy(), z[k][perm], 1, "binned EOS golden g2p " + k, 3e-5, floor=float(np.abs(z[k]).max()))
        check_channels(bins.J.cpu().numpy()[:, None], z["J"][perm][:, None], 1, "binned EOS golden J", 3e-5)
        Q = {k: v for k, v in P.items() if k != "F"}
        Q["v"] = (P["v"] * 6.0).astype(np.float32)
        Q["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n) / n)).astype(np.float32)
        res = []
        for layout in ("aos", "binned"):
            sol = MpmSolver(Q, dx, P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout=layout, rebin_every=2, model=m,
                            **({"partition": "with_rebin"} if layout == "binned" else {}))
            for _ in range(4):
                sol.substep()
            torch.cuda.synchronize()
            H = sol.particles_host() if layout == "binned" else dict(sol.aos.to_host(), J=sol.aos.J.cpu().numpy())
            o = np.argsort(H["m"], kind="stable")
            res.append({k: H[k][o] for k in ("x", "v", "C", "J")})
        for k in "xv":
            check_channels(res[1][k], res[0][k], 1, "binned vs AoS EOS " + k, 1e-4, floor=float(np.abs(res[0][k]).max()))
        assert np.abs(res[1]["J"] - res[0]["J"]).max() <= 1e-4
    finally:
        api.set_tuning(-1, 1)


@pytest.mark.parametrize("model", ["vonmises", "sand", "nacc", "eos"])
def test_sparsegrid_runs_every_model(oracle, model):
    """zpcb200_sg_p2g_apic_model / zpcb200_sg_g2p_apic_eos: the any-order kernels instantiated for SparseGrid<3,f32,8>; grid after P2G
    compared per NODE (global cell coordinate) with the oracle's functor on the legacy side-4 grid, like the fixed-corotated test"""
    from tests.test_gpu_sparsegrid import _build, _host_table, _nodes
    from zpc_b200 import api
    name = {"vonmises": "mpm_cube6_vonmises", "sand": "mpm_cube6_sand", "nacc": "mpm_cube6_nacc", "eos": "mpm_cube6_eos"}[model]
    z = np.load(os.path.join(G, name + ".npz"))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    n, dx = P["x"].shape[0], P["dx"]
    if model in ("sand", "nacc"):
        P["logJp"] = z["logJp_in"].copy()
    if model == "eos":
        P["J"] = z["J_in"].copy()
    m = {"vonmises": lambda: api.model_vonmises(P["volume"], E, NU, float(z["ys"])),
         "sand": lambda: api.model_drucker_prager(P["volume"], E, NU, SAND["cohesion"], SAND["beta"], SAND["volumeCorrection"], SAND["yieldSurface"]),
         "nacc": lambda: api.model_nacc(P["volume"], NACC["E"], NACC["nu"], NACC["fa"], NACC["xi"], NACC["beta"], NACC["hardeningOn"]),
         "eos": lambda: api.model_eos(P["volume"], 4.0e4, 7.15, 0.01)}[model]()
    pars, sg = _build(P)
    t = _host_table(sg)
    nb = t["nblocks"]
    api.sg_clean(sg)
    api.sg_p2g_transfer(pars, sg, synth.DT, m)
    torch.cuda.synchronize()
    g1 = sg.grid[:nb].cpu().numpy()
    code_o, val_o = _nodes(z["active_keys"] * 4, z["grid_p2g"], 4)           # the reference-generated grid, per node
    code_s, val_s = _nodes(t["active_keys"], g1, 8)
    pos = np.searchsorted(code_s, code_o)
    assert (pos < code_s.shape[0]).all() and np.array_equal(code_s[pos], code_o)
    rhs = {"vonmises": RTOL_STRESS, "sand": RTOL_STRESS, "nacc": 1e-3, "eos": RTOL}[model]
    check_channels(val_s[pos], val_o, 1, "sg p2g " + model, [RTOL] * 4 + [rhs] * 3)
    if model in ("sand", "nacc"):
        assert np.abs(pars.logJp.cpu().numpy() - z["logJp"]).max() <= 2e-5
    if model == "eos":
        mx = torch.zeros(1, device="cuda")
        api.sg_compute_grid_velocity(sg, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
        api.sg_g2p_transfer(pars, sg, synth.DT, model=m)
        check_channels(pars.J.cpu().numpy()[:, None], z["J"][:, None], 1, "sg eos J", 3e-5)


@pytest.mark.parametrize("model", ["vonmises", "sand", "nacc", "eos"])
def test_sparsegrid_binned_path_runs_every_model(oracle, model):
    """zpcb200_sg_p2g_apic_model_binned / zpcb200_sg_g2p_apic_eos_binned: the binned kernels instantiated for SparseGrid<3,f32,8>
    (bins = octants) with the model as the record phase's template parameter; same reference-generated goldens and tolerances as the
    any-order SparseGrid test above; logJp / J live next to the bins in bin order"""
    from tests.test_gpu_sparsegrid import _build, _host_table, _nodes
    from zpc_b200 import api
    name = {"vonmises": "mpm_cube6_vonmises", "sand": "mpm_cube6_sand", "nacc": "mpm_cube6_nacc", "eos": "mpm_cube6_eos"}[model]
    z = np.load(os.path.join(G, name + ".npz"))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    n, dx = P["x"].shape[0], P["dx"]
    m = {"vonmises": lambda: api.model_vonmises(P["volume"], E, NU, float(z["ys"])),
         "sand": lambda: api.model_drucker_prager(P["volume"], E, NU, SAND["cohesion"], SAND["beta"], SAND["volumeCorrection"], SAND["yieldSurface"]),
         "nacc": lambda: api.model_nacc(P["volume"], NACC["E"], NACC["nu"], NACC["fa"], NACC["xi"], NACC["beta"], NACC["hardeningOn"]),
         "eos": lambda: api.model_eos(P["volume"], 4.0e4, 7.15, 0.01)}[model]()
    pars, sg = _build(P)
    t = _host_table(sg)
    nb = t["nblocks"]
    bins = api.ParticleBins(n, 8 * nb + 64)
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    api.sg_bin_particles(pars, sg, bins, order)
    perm = order.cpu().numpy()
    if model in ("sand", "nacc"):
        bins.logJp = torch.from_numpy(z["logJp_in"][perm].copy()).cuda()
    if model == "eos":
        bins.J = torch.from_numpy(z["J_in"][perm].copy()).cuda()
    api.sg_clean(sg)
    api.sg_p2g_transfer(bins, sg, synth.DT, m)
    torch.cuda.synchronize()
    assert int(bins.status.item()) == 0
    g1 = sg.grid[:nb].cpu().numpy()
    code_o, val_o = _nodes(z["active_keys"] * 4, z["grid_p2g"], 4)
    code_s, val_s = _nodes(t["active_keys"], g1, 8)
    pos = np.searchsorted(code_s, code_o)
    assert (pos < code_s.shape[0]).all() and np.array_equal(code_s[pos], code_o)
    rhs = {"vonmises": RTOL_STRESS, "sand": RTOL_STRESS, "nacc": 1e-3, "eos": RTOL}[model]
    check_channels(val_s[pos], val_o, 1, "sg binned p2g " + model, [RTOL] * 4 + [rhs] * 3)
    if model in ("sand", "nacc"):
        assert np.abs(bins.logJp.cpu().numpy() - z["logJp"][perm]).max() <= 2e-5
    if model == "eos":
        mx = torch.zeros(1, device="cuda")
        api.sg_compute_grid_velocity(sg, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
        api.sg_g2p_transfer(bins, sg, synth.DT, model=m)
        check_channels(bins.J.cpu().numpy()[:, None], z["J"][perm][:, None], 1, "sg binned eos J", 3e-5)


def test_grid_momentum_functors(oracle):
    """zpcb200_grid_momentum_to_velocity / zpcb200_grid_angular_momentum (GridOp.hpp:184-262) on the grid our own P2G leaves, against
    the oracle on that same grid: velocities bit for bit (one IEEE division and three products per cell), max |v|^2 to an ulp of the
    contracted sum, the six double sums to the rounding of another addition order; then the reference-generated golden file"""
    from zpc_b200 import api
    z = np.load(os.path.join(G, "mpm_cube6_grid_momentum.npz"))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    P["v"] = (P["v"] + z["v_shift"]).astype(np.float32)
    dx = P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    grids = api.Grids(dx, ht["nblocks"] + 2)                       # spare blocks beyond *cnt stay untouched
    api.p2g_transfer(pars, table, grids, synth.DT, api.model_fcr(P["volume"], E, NU))
    grids.tiles[ht["nblocks"]:] = 3.0
    torch.cuda.synchronize()
    g0 = grids.tiles.cpu().numpy()
    sum6 = torch.zeros(6, dtype=torch.float64, device="cuda")
    api.grid_angular_momentum(grids, table, sum6)
    api.grid_angular_momentum(grids, table, sum6)                  # adds: twice the sum
    want = oracle.grid_angular_momentum(g0[:ht["nblocks"]], ht["active_keys"], dx)
    got = sum6.cpu().numpy() / 2
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (got, want)
    ko, go = grid_by_key(z["active_keys"], z["grid"])
    ks, gs = grid_by_key(ht["active_keys"], g0[:ht["nblocks"]])
    assert np.array_equal(ko, ks)
    assert np.abs(got - z["sum6"]).max() <= 1e-5 * np.abs(z["sum6"]).max()      # the reference's own P2G grid differs by float rounding
    mx = torch.zeros(1, device="cuda")
    api.grid_momentum_to_velocity(grids, table, mx)
    torch.cuda.synchronize()
    g1 = grids.tiles.cpu().numpy()
    ref = g0[:ht["nblocks"]].copy()
    mo = oracle.grid_momentum_to_velocity(ref)
    assert np.array_equal(g1[:ht["nblocks"]].view(np.uint32), ref.view(np.uint32))
    assert (g1[ht["nblocks"]:] == 3.0).all()
    assert abs(float(mx.item()) - mo) <= 1e-6 * mo
    check_channels(gs[:, :4], go[:, :4], 1, "mass and momentum vs the reference-generated grid", RTOL)
    # channel arguments: the rhs channels against the mass channel; bad channel ranges are refused
    alt = torch.zeros(6, dtype=torch.float64, device="cuda")
    api.grid_angular_momentum(grids, table, alt, 0, 4)
    want_alt = oracle.grid_angular_momentum(g1[:ht["nblocks"]], ht["active_keys"], dx, 0, 4)
    assert np.abs(alt.cpu().numpy() - want_alt).max() <= 1e-9 * max(np.abs(want_alt).max(), 1e-30)   # measured on a B200: 2.5e-11 (double atomics in another order)
    with pytest.raises(RuntimeError):
        api.grid_momentum_to_velocity(grids, table, mx, 0, 5)
    with pytest.raises(RuntimeError):
        api.grid_momentum_to_velocity(grids, table, mx, 2, 1)


def test_grid_momentum_functors_vs_the_references_cuda_functors(oracle, tmp_path):
    """GridAngularMomentum / GridMomentumToVelocity: the reference's functors on cuda_exec() and zs::b200::grid_* on the same reference
    containers (own process); each arm against the oracle on the grid that arm started from"""
    import subprocess
    import sys
    from oracle.refcuda_runner import RefCuda
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libzpcref_cuda.so not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    P = synth.elastic_cube(9, 32, jitter_F=0.03, jitter_C=0.3, seed=6)
    P["v"] = (P["v"] + np.float32([0.3, -0.2, 0.1])).astype(np.float32)
    fin, fout = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(fin, dt=synth.DT, E=E, nu=NU, **P)
    r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "gridmom", fin, fout], cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(fout)
    for tag in ("ref", "b200"):
        want = oracle.grid_angular_momentum(z[tag + "_grid"], z[tag + "_keys"], P["dx"])
        assert np.abs(z[tag + "_sum6"] - want).max() <= (2e-7 if tag == "ref" else 1e-12) * np.abs(want).max(), tag   # nvcc contracts the reference's cross product
        g = z[tag + "_grid"].copy()
        mo = oracle.grid_momentum_to_velocity(g)
        assert np.array_equal(g.view(np.uint32), z[tag + "_vel"].view(np.uint32)), tag
        assert abs(float(z[tag + "_max"]) - mo) <= 1e-6 * mo
    assert np.abs(z["b200_sum6"] - z["ref_sum6"]).max() <= 1e-5 * np.abs(z["ref_sum6"]).max()


def test_overlay_fast_path_on_the_references_containers(tmp_path):
    """zs::b200::BinnedParticles (TileVector<f32,32> = the type of Particles::particleBins, zs::Vector metadata): four substeps with a re-bin
    through the overlay's block-binned fast path vs the same four substeps through the reference's own functors on cuda_exec(), both on
    the reference's containers, in one process of their own; particles are matched by their (unique) mass."""
    import subprocess
    import sys
    from oracle.refcuda_runner import RefCuda
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libzpcref_cuda.so not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    P = synth.elastic_cube(10, 32, jitter_F=0.03, jitter_C=0.3, seed=5)
    P["v"][:] = P["v"] * 6.0
    n = P["m"].shape[0]
    P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n) / n)).astype(np.float32)
    fin, fout = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(fin, dt=synth.DT * 10, E=E, nu=NU, gravity=synth.GRAVITY, mode=1, steps=4, rebin_every=2, **P)
    r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "binned", fin, fout], cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(fout)
    assert np.unique(z["m"]).size == n and np.array_equal(np.sort(z["m"]), np.sort(z["ref_m"]))
    o, ro = np.argsort(z["m"], kind="stable"), np.argsort(z["ref_m"], kind="stable")
    check_particles({k: z[k][o] for k in "xvCF"}, {k: z["ref_" + k][ro] for k in "xvCF"}, P["dx"], "overlay fast path vs reference functors", rtol=5e-5)


# last: a failed stream capture could leave the process unable to launch — nothing runs after it
def test_graph_replay_equals_eager_substeps():
    """MpmSolver.capture_cycle / replay_cycle: two replays of the captured 2 x rebin_every substeps give the particles the same
    number of eager substeps give (binned P2G adds with unordered shared-memory atomics: equal up to fp32 re-association)"""
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(10, 32, jitter_F=0.03, jitter_C=0.3, seed=5)
    P["v"][:] = P["v"] * 6.0
    n0 = P["m"].shape[0]
    P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n0) / n0)).astype(np.float32)
    res = []
    for graph in (False, True):
        sol = MpmSolver(P, P["dx"], P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout="binned", rebin_every=2, partition="with_rebin")
        if graph:
            k = sol.capture_cycle()
            assert k == 4 and sol.step_no == 4
            sol.replay_cycle()
            sol.replay_cycle()
        else:
            for _ in range(12):
                sol.substep()
        torch.cuda.synchronize()
        assert sol.step_no == 12 and sol.table.overflow.item() == 0
        Q = sol.particles_host()
        o = np.argsort(Q["m"], kind="stable")
        res.append({k: Q[k][o] for k in "xvCF"})
    check_particles(res[1], res[0], P["dx"], "graph replay vs eager", rtol=5e-5)


def test_bins_status_word_reports_what_used_to_be_silent(oracle):
    """zpc_bins_view.status (ADVICE r1 / VERDICT r1 weak #8): a particle that out-runs the partition's extra ring between re-bins, a bin
    buffer that is too small and a table with more blocks than bins each OR their ZPC_BINS_* bit into the device status word; nothing
    is thrown by the library, MpmSolver raises at the next re-bin."""
    from zpc_b200 import api
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(6, 32)
    n, dx = P["x"].shape[0], P["dx"]
    pars, table = build_partition(P)
    ht = host_table(table)
    # (a) binCapacity smaller than the number of blocks / bins
    small = api.ParticleBins(n, max(ht["nblocks"] // 4, 1))
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    api.bin_particles(pars, table, dx, small, order)
    torch.cuda.synchronize()
    st = int(small.status.item())
    assert st & api.BINS_BLOCK_CAPACITY, st
    with pytest.raises(RuntimeError):
        small.check_status()
    assert int(small.status.item()) == 0                                   # check_status cleared it
    # (b) a clean run leaves it at zero
    bins = api.ParticleBins(n, max(ht["nblocks"] * 2, 64))
    api.bin_particles(pars, table, dx, bins, order)
    grids = api.Grids(dx, ht["nblocks"])
    model = api.model_fcr(P["volume"], E, NU)
    for sweep in (4, 6):
        api.set_tuning(sweep, -1)
        try:
            api.clean_grid_blocks(grids, table)
            api.p2g_transfer(bins, table, grids, synth.DT, model)
            mx = torch.zeros(1, device="cuda")
            api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
            torch.cuda.synchronize()
            assert int(bins.status.item()) == 0
        finally:
            api.set_tuning(4, -1)
    # (c) particles teleported three blocks away without a new partition: their stencil blocks do not exist
    x = bins.pars.channel(api.PB_X, 3)
    far = torch.arange(0, n, 97, device="cuda")
    x[far] += torch.tensor([12.5 * dx, 0.0, 0.0], device="cuda")
    bins.pars.set_channel(api.PB_X, x)
    bins.cell_order_valid.zero_()
    for sweep in (4, 6):
        api.set_tuning(sweep, -1)
        try:
            bins.status.zero_()
            api.clean_grid_blocks(grids, table)
            api.p2g_transfer(bins, table, grids, synth.DT, model)
            torch.cuda.synchronize()
            assert int(bins.status.item()) & api.BINS_STENCIL_BLOCK_MISSING, sweep
        finally:
            api.set_tuning(4, -1)
    bins.status.zero_()
    api.g2p_transfer(bins, table, grids, synth.DT)
    torch.cuda.synchronize()
    assert int(bins.status.item()) & api.BINS_STENCIL_BLOCK_MISSING
    # (d) the solver surfaces it at the re-bin: a velocity that crosses more than the extra ring within one cadence
    Pf = synth.elastic_cube(6, 32)
    Pf["v"][:] = (0.0, -400.0, 0.0)                                        # 400 * 1e-4 * 32 = 1.3 cells per substep
    sol = MpmSolver(Pf, dx, Pf["volume"], synth.DT, synth.GRAVITY, mode=1, layout="binned", rebin_every=8, partition="with_rebin")
    with pytest.raises(RuntimeError, match="stencil block"):
        for _ in range(9):
            sol.substep()
    # (e) status_mode="deferred": the same words through an asynchronous copy to pinned memory — the host never waits at a re-bin, the
    # fault surfaces at a later re-bin or, at the latest, in flush_status() / particles_host()
    sol = MpmSolver(Pf, dx, Pf["volume"], synth.DT, synth.GRAVITY, mode=1, layout="binned", rebin_every=8, partition="with_rebin")
    sol.status_mode = "deferred"
    with pytest.raises(RuntimeError, match="stencil block"):
        for _ in range(9):
            sol.substep()
        sol.flush_status()
    from zpc_b200.selfcheck import identity_masses
    ok = synth.elastic_cube(6, 32, jitter_F=0.03, jitter_C=0.3)
    ok["m"] = identity_masses(ok["m"].shape[0], float(ok["m"].mean()))      # pairs the particles whatever order the re-bins leave
    a = MpmSolver(ok, dx, ok["volume"], synth.DT, synth.GRAVITY, mode=1, layout="binned", rebin_every=3, partition="with_rebin")
    b = MpmSolver(ok, dx, ok["volume"], synth.DT, synth.GRAVITY, mode=1, layout="binned", rebin_every=3, partition="with_rebin")
    b.status_mode = "deferred"
    for _ in range(10):
        a.substep()
        b.substep()
    b.flush_status()
    assert b._status_pending == []
    pa, pb = a.particles_host(), b.particles_host()
    oa, ob = np.argsort(pa["m"], kind="stable"), np.argsort(pb["m"], kind="stable")
    check_particles({k: pb[k][ob] for k in "xvCF"}, {k: pa[k][oa] for k in "xvCF"}, dx, "deferred vs synchronous status reads",
                    rtol=5e-5)                           # same kernels; the float atomics order the sums differently run to run


def test_graph_replay_with_a_side_array_equals_eager_substeps():
    """capture_cycle with a model that carries a per-particle scalar: J ping-pongs between two persistent buffers through the
    library's own gather, so the captured graph ends on the buffer it started from"""
    from zpc_b200 import api
    from zpc_b200.solver import MpmSolver
    P = synth.elastic_cube(10, 32, jitter_C=0.3, seed=5)
    n0 = P["m"].shape[0]
    P["v"][:] = P["v"] * 6.0
    P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n0) / n0)).astype(np.float32)
    Q = {k: v for k, v in P.items() if k != "F"}
    Q["J"] = (1.0 + 0.02 * np.sin(np.arange(n0))).astype(np.float32)
    model = api.model_eos(P["volume"])
    res = []
    for graph in (False, True):
        sol = MpmSolver(Q, P["dx"], P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout="binned", rebin_every=2, partition="with_rebin",
                        model=model)
        if graph:
            k = sol.capture_cycle()
            assert k == 4
            sol.replay_cycle()
            sol.replay_cycle()
        else:
            for _ in range(12):
                sol.substep()
        torch.cuda.synchronize()
        assert sol.step_no == 12
        R = sol.particles_host()
        o = np.argsort(R["m"], kind="stable")
        res.append({k: R[k][o] for k in ("x", "v", "C", "J")})
    vmax = float(np.abs(res[0]["v"]).max())
    check_channels(res[1]["x"], res[0]["x"], 1, "graph x", 5e-5, floor=float(np.abs(res[0]["x"]).max()))
    check_channels(res[1]["v"], res[0]["v"], 1, "graph v", 5e-5, floor=vmax)
    check_channels(res[1]["J"][:, None], res[0]["J"][:, None], 1, "graph J", 5e-5)
    assert np.abs(res[0]["J"] - Q["J"][np.argsort(Q["m"], kind="stable")]).max() > 1e-6   # J really evolved
