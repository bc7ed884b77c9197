"""Host-side control flow of MpmSolver without a GPU: the C ABI is replaced by a recorder that returns 0 (and a size for the
two-phase queries), tensors live on the CPU.  Checks the order of the functor calls of a substep for every layout / model /
collider combination, the re-bin cadence and the argument checks — a typo in the Python plumbing fails here, not on the box."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from zpc_b200 import api, synth  # noqa: E402
from zpc_b200.solver import MpmSolver  # noqa: E402


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("zpcb200_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls.append(name)
            # two-phase entries: (temp, &bytes, ...) with temp == NULL -> report a size
            if len(args) >= 2 and args[0] is None and hasattr(args[1], "_obj"):
                args[1]._obj.value = 1024
            return 0
        fn.__name__ = name
        return fn


@pytest.fixture
def recorder(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(api, "lib", lambda: rec)
    monkeypatch.setattr(api, "_stream_ptr", lambda stream=None: C.c_void_p(None))
    class _CpuScratch:
        def get(self, nbytes, device):
            return torch.empty(max(int(nbytes), 256), dtype=torch.uint8)
    monkeypatch.setattr(api, "_scratch", _CpuScratch())
    monkeypatch.setattr(api, "vec3_port", lambda x: api.zpc_port(x.data_ptr(), 0, 0, 0, 3))      # the real one insists on a CUDA tensor
    return rec


def _names(rec):
    out = [c.replace("zpcb200_", "") for c in rec.calls]
    rec.calls.clear()
    return out


def test_binned_substep_sequence_and_rebin_cadence(recorder):
    P = synth.elastic_cube(4, 16)
    sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="binned", rebin_every=3, device="cpu", partition="with_rebin")
    assert _names(recorder) == ["partition_build", "partition_build", "bin_particles", "bin_particles"]     # size query + run, each
    step = ["clean_grid", "p2g_apic_fcr_binned", "grid_update", "g2p_apic_binned"]
    for i in range(7):
        sol.substep()
        got = _names(recorder)
        if i in (3, 6):                                   # step_no 3 and 6: partition + re-bin first
            assert got == ["partition_build"] * 2 + ["rebin_particles"] * 2 + step, (i, got)
        else:
            assert got == step, (i, got)
    sol2 = MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="binned", rebin_every=0, device="cpu", partition="every_step")
    _names(recorder)
    sol2.substep()
    assert _names(recorder) == ["partition_build"] * 2 + step


def test_models_and_colliders_pick_the_right_entries(recorder):
    P = synth.elastic_cube(4, 16)
    n = P["x"].shape[0]
    cols = [api.plane_collider((0, 0.1, 0), (0, 1, 0), api.COLLIDER_SEPARATE), api.cuboid_collider((0, 0, 0), (0.1, 0.1, 0.1))]
    vm = api.model_vonmises(P["volume"])
    sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="binned", device="cpu", model=vm, colliders=cols)
    _names(recorder)
    sol.substep()
    assert _names(recorder) == ["partition_build"] * 2 + ["clean_grid", "p2g_apic_vonmises_binned", "grid_update_bc", "g2p_apic_binned"]
    for model, p2g, g2p, extra in ((api.model_drucker_prager(P["volume"]), "p2g_apic_drucker_prager", "g2p_apic", "logJp"),
                                   (api.model_nacc(P["volume"]), "p2g_apic_nacc", "g2p_apic", "logJp"),
                                   (api.model_eos(P["volume"]), "p2g_apic_eos", "g2p_apic_eos", "J"),
                                   (vm, "p2g_apic_vonmises", "g2p_apic", None), (None, "p2g_apic_fcr", "g2p_apic", None)):
        Q = dict(P)
        if extra:
            Q[extra] = np.zeros(n, np.float32)
        s = MpmSolver(Q, P["dx"], P["volume"], synth.DT, layout="aos", device="cpu", model=model)
        _names(recorder)
        s.substep()
        assert _names(recorder) == ["partition_build"] * 2 + ["clean_grid", p2g, "grid_update", g2p]
    with pytest.raises(ValueError):                       # the equation of state needs J
        MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="binned", device="cpu", model=api.model_eos(P["volume"]))
    Qj = {k: v for k, v in P.items() if k != "F"}
    Qj["J"] = np.ones(n, np.float32)
    s = MpmSolver(Qj, P["dx"], P["volume"], synth.DT, layout="binned", device="cpu", model=api.model_eos(P["volume"]), rebin_every=2,
                  partition="with_rebin")
    _names(recorder)
    for _ in range(3):
        s.substep()
    step = ["clean_grid", "p2g_apic_eos_binned", "grid_update", "g2p_apic_eos_binned"]
    # the side array follows the re-bin through the library's own gather (no eager torch indexing on that path)
    assert _names(recorder) == step * 2 + ["partition_build"] * 2 + ["rebin_particles_ordered"] * 2 + ["gather_f32"] + step
    assert s.bins.J is not None and s.bins_alt.J is not None and s.bins.J.data_ptr() != s.bins_alt.J.data_ptr()
    assert s.bins.J is not None and "J" in s.particles_host()
    # plastic models on the binned layout: logJp rides next to the bins and follows every re-bin
    Q = dict(P, logJp=np.linspace(-1, 0, n).astype(np.float32))
    s = MpmSolver(Q, P["dx"], P["volume"], synth.DT, layout="binned", device="cpu", model=api.model_nacc(P["volume"]), rebin_every=2,
                  partition="with_rebin")
    assert s.bins.logJp is not None and s.bins.logJp.shape == (n,)
    _names(recorder)
    for _ in range(3):
        s.substep()
    step = ["clean_grid", "p2g_apic_nacc_binned", "grid_update", "g2p_apic_binned"]
    assert _names(recorder) == step * 2 + ["partition_build"] * 2 + ["rebin_particles_ordered"] * 2 + ["gather_f32"] + step
    assert s.bins.logJp is not None and "logJp" in s.particles_host()
    with pytest.raises(ValueError):
        MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="binned", device="cpu", model=api.model_nacc(P["volume"]))   # no logJp given
    with pytest.raises(ValueError):                       # plastic model without logJp
        s = MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="aos", device="cpu", model=api.model_nacc(P["volume"]))
        s.substep()
    with pytest.raises(ValueError):
        MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="aos", device="cpu", colliders=cols * 3)


def test_sparsegrid_solver_sequence_for_every_model(recorder):
    """SgMpmSolver (SparseGrid<3,f32,8>, octant bins): the functor sequence of a substep, the re-bin cadence and — for the models that
    carry a per-particle scalar — the side array following every re-bin through the library's own gather"""
    from zpc_b200.sg_solver import SgMpmSolver
    P = synth.elastic_cube(4, 16)
    n = P["x"].shape[0]
    def _names_sg():                                        # without the host helpers the SparseGrid constructor calls
        return [c for c in _names(recorder) if c not in ("bht_table_size", "bht_params")]
    sol = SgMpmSolver(P, P["dx"], P["volume"], synth.DT, rebin_every=2, device="cpu")
    assert _names_sg() == ["sg_partition_build"] * 2 + ["sg_bin_particles"] * 2
    step = ["sg_clean", "sg_p2g_apic_fcr_binned", "sg_grid_update", "sg_g2p_apic_binned"]
    for i in range(5):
        sol.substep()
        got = _names_sg()
        assert got == (["sg_partition_build"] * 2 + ["sg_rebin_particles"] * 2 if i in (2, 4) else []) + step, (i, got)
    assert set(sol.particles_host()) == {"x", "v", "m", "C", "F"}
    Pj = {k: v for k, v in P.items() if k != "F"}
    Pj["J"] = np.ones(n, np.float32)
    cases = ((api.model_vonmises(P["volume"]), P, None, "sg_g2p_apic_binned"),
             (api.model_nacc(P["volume"]), dict(P, logJp=np.zeros(n, np.float32)), "logJp", "sg_g2p_apic_binned"),
             (api.model_drucker_prager(P["volume"]), dict(P, logJp=np.zeros(n, np.float32)), "logJp", "sg_g2p_apic_binned"),
             (api.model_eos(P["volume"]), Pj, "J", "sg_g2p_apic_eos_binned"))
    for model, Q, side, g2p in cases:
        s = SgMpmSolver(Q, P["dx"], P["volume"], synth.DT, rebin_every=2, device="cpu", model=model)
        assert _names_sg() == ["sg_partition_build"] * 2 + ["sg_bin_particles"] * 2 + (["gather_f32"] if side else [])
        step = ["sg_clean", "sg_p2g_apic_model_binned", "sg_grid_update", g2p]
        for _ in range(3):
            s.substep()
        assert _names_sg() == step * 2 + ["sg_partition_build"] * 2 + ["sg_rebin_particles"] * 2 + (["gather_f32"] if side else []) + step
        if side:
            a, b = getattr(s.bins, side), getattr(s.bins_alt, side)
            assert a is not None and b is not None and a.data_ptr() != b.data_ptr() and side in s.particles_host()
    with pytest.raises(ValueError):
        SgMpmSolver(P, P["dx"], P["volume"], synth.DT, device="cpu", model=api.model_nacc(P["volume"]))   # no logJp given


def test_status_words_are_read_at_every_rebin_and_after_every_graph_replay(recorder):
    """the device status words (a stencil block missing from the partition, bin capacity, table overflow, whatever the owner
    registered) reach the host at every re-bin and — the replayed re-bins cannot read back — once after every graph replay"""
    P = synth.elastic_cube(4, 16)
    sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, layout="binned", rebin_every=2, device="cpu", partition="with_rebin")
    sol.substep()
    sol.substep()
    sol.bins.status.fill_(8)                              # raised by a P2G / G2P since the last re-bin
    with pytest.raises(RuntimeError, match="stencil block"):
        sol.substep()                                     # step_no 2: the re-bin reads the retired bins' word
    assert int(sol.bins_alt.status.item()) == 0           # ... and clears it
    extra = torch.zeros(1, dtype=torch.int32)
    sol.extra_status.append((extra, "halo maps"))

    class _Graph:
        replays = 0

        def replay(self):
            self.replays += 1
    sol._graph, sol._graph_len = _Graph(), 4
    n0 = sol.step_no
    sol.replay_cycle()
    assert sol._graph.replays == 1 and sol.step_no == n0 + 4
    sol.table.overflow.fill_(1)
    with pytest.raises(RuntimeError, match="overflow"):
        sol.replay_cycle()
    sol.table.overflow.zero_()
    extra.fill_(3)
    with pytest.raises(RuntimeError, match=r"halo maps \(status 3\)"):
        sol.replay_cycle()
    assert int(extra.item()) == 0
    sol.check_status = False                              # opt-out: nothing is read
    sol.bins.status.fill_(2)
    sol.replay_cycle()
    sol.bins.status.zero_()
    # deferred mode: the words are copied out asynchronously and looked at when the copy has landed (CPU tensors: at once)
    sol.check_status, sol.status_mode = True, "deferred"
    sol.replay_cycle()
    assert sol._status_pending == []
    sol.bins.status.fill_(8)
    with pytest.raises(RuntimeError, match="since the last re-bin: a stencil block"):
        sol.flush_status()
    assert int(sol.bins.status.item()) == 0 and sol._status_pending == []
    sol.flush_status()
    sol.bins_alt.status.fill_(1)
    with pytest.raises(RuntimeError, match="home block"):
        sol.particles_host()                              # a host read of the results flushes first


def test_particle_range_views_are_pointer_offsets():
    P = synth.elastic_cube(3, 16)
    P["logJp"] = np.zeros(P["x"].shape[0], np.float32)
    pars = api.Particles(P, device="cpu")
    full, part = pars.view(), pars.range(10, 50).view()
    assert part.count == 40 and full.count == pars.n
    assert part.X - full.X == 4 * 3 * 10 and part.F - full.F == 4 * 9 * 10 and part.M - full.M == 4 * 10 and part.logJp - full.logJp == 40
    assert not part.J and not full.J
    with pytest.raises(ValueError):
        pars.view(5, pars.n + 1)


# ---- DistMpmSolver control flow over gloo, same recording ABI ------------------------------------------------------------
def _dist_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zpc_b200.dist_solver import DistMpmSolver, HaloExchange
        rec = _Recorder()

        class _CpuScratch:
            def get(self, nbytes, device):
                return torch.empty(max(int(nbytes), 256), dtype=torch.uint8)
        api.lib = lambda: rec
        api._stream_ptr = lambda stream=None: C.c_void_p(None)
        api._scratch = _CpuScratch()
        api.vec3_port = lambda x: api.zpc_port(x.data_ptr(), 0, 0, 0, 3)
        P = synth.elastic_cube_slab(4, 16, rank, world)
        halo = HaloExchange(None, 7, "cpu", lambda g, ids, buf: None, lambda g, ids, buf: None)
        sol = DistMpmSolver(P, P["dx"], P["volume"], synth.DT, layout="aos", device="cpu", halo=halo)
        assert sol.transport == "nccl"
        hin = {k: torch.from_numpy(P[k]) for k in ("x", "v", "m", "C", "F")}
        hout = {k: torch.empty_like(hin[k]) for k in ("x", "v", "C", "F")}
        rec.calls.clear()
        mx = sol.substep_host(hin, hout)
        names = [c.replace("zpcb200_", "") for c in rec.calls]
        assert names == ["partition_build"] * 2 + ["clean_grid", "p2g_apic_fcr", "grid_update", "g2p_apic"], names
        assert mx == 0.0 and torch.equal(hout["x"], hin["x"])        # the recorder computes nothing: buffers pass through
        # the binned solver's substep on the same stand-ins
        sol2 = DistMpmSolver(P, P["dx"], P["volume"], synth.DT, device="cpu", halo=halo, rebin_every=2)
        rec.calls.clear()
        for _ in range(3):
            sol2.substep()
        names = [c.replace("zpcb200_", "") for c in rec.calls]
        step = ["clean_grid", "p2g_apic_fcr_binned", "grid_update", "g2p_apic_binned"]
        assert names == step * 2 + ["partition_build"] * 2 + ["rebin_particles"] * 2 + step, names
        sol2.max_vel_sqr()
        # graph replay reads the status words afterwards (deferred: through an asynchronous copy; bench.py --gpus N runs on it)
        class _Graph:
            def replay(self):
                pass
        sol2._graph, sol2._graph_len, sol2.local.status_mode = _Graph(), 4, "deferred"
        n0 = sol2.local.step_no
        sol2.replay_cycle()
        assert sol2.local.step_no == n0 + 4 and sol2.local._status_pending == []
        sol2.local.table.overflow.fill_(1)
        try:
            sol2.replay_cycle()
            raise AssertionError("the overflow flag went unnoticed")
        except RuntimeError as ex:
            assert "overflow" in str(ex)
        sol2.local.table.overflow.zero_()
        sol2.local.flush_status()
        # a collider and a per-particle scalar on the multi-GPU fast path: the fused boundary update and the J variant of the G2P
        # (this path used to call the plain update and the F variant whatever the solver was built with)
        Pj = {k: v for k, v in P.items() if k != "F"}
        Pj["J"] = np.ones(P["x"].shape[0], np.float32)
        sol3 = DistMpmSolver(Pj, P["dx"], P["volume"], synth.DT, device="cpu", halo=halo, rebin_every=2, model=api.model_eos(P["volume"]),
                             colliders=[api.plane_collider((0.0, 0.1, 0.0), (0.0, 1.0, 0.0))])
        rec.calls.clear()
        for _ in range(3):
            sol3.substep()
        names = [c.replace("zpcb200_", "") for c in rec.calls]
        step = ["clean_grid", "p2g_apic_eos_binned", "grid_update_bc", "g2p_apic_eos_binned"]
        assert names == step * 2 + ["partition_build"] * 2 + ["rebin_particles_ordered"] * 2 + ["gather_f32"] + step, names
        sol3.max_vel_sqr()
        q.put(rank)
    finally:
        dist.destroy_process_group()


def test_dist_solver_control_flow_over_gloo():
    import socket
    import torch.multiprocessing as mp
    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=10) for _ in range(2)) == [0, 1]


def test_self_check_cloud_and_identity_tags():
    """zpc_b200/selfcheck.py: the check's cloud keeps every slab >= 3 blocks wide for 2..8 ranks, and its identity tags (masses) are
    exact and distinct at the largest size — with the round-1 tag (a relative step of 0.1 / n) the 7 M-particle cloud of an 8-rank
    run held duplicates and the comparison paired different particles"""
    import numpy as np
    from zpc_b200 import selfcheck, synth
    for world in (2, 4, 8):
        s, G = selfcheck.check_cloud(world)
        assert s // world >= 12 and G >= s + 16
        c0, c1 = synth.slab_cell_range(s, world - 1, world)
        assert (c1 - c0) // (s * s) >= 12
    s, G = selfcheck.check_cloud(8)
    n0 = 8 * s ** 3
    m = selfcheck.identity_masses(n0, 2.4e-4)
    assert m.dtype == np.float32 and np.unique(m).size == n0 and 0.5 <= m[0] / 2.4e-4 <= 2.0 and m[-1] < 2 * m[0] * 1.0001
    old = (np.float32(2.4e-4) * (1.0 + 0.1 * np.arange(n0) / n0)).astype(np.float32)
    assert np.unique(old).size < n0

