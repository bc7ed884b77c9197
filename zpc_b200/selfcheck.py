"""Multi-GPU self-check: N-GPU sharded substeps against the single-GPU solver on the same cloud (no oracle involved — both sides are
this library; the single-GPU path is what the parity tests pin against the reference).  Collective: every rank of the default process
group calls `multi_gpu_parity`; rank 0 holds the full cloud and runs the single-GPU solver.

Used by `bench.py --gpus N` (N > 1: the line carries the outcome, so a scaling number never stands without it), by
tests/dist_check.py (torchrun) and by tests/test_gpu_mpm.py::test_multi_gpu_substeps_match_single_gpu (needs >= 2 GPUs)."""
import numpy as np
import torch
import torch.distributed as dist

from . import synth
from .dist_solver import DistMpmSolver
from .solver import MpmSolver


def _rel(a, b, floor):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)).max())


def identity_masses(n0, m0):
    """Identity tag in the mass: consecutive float32 values (2^23 + i) x a power of two near the physical mass m0 — exact and distinct
    for up to 2^23 particles.  (A relative step of 0.1 / n0 collapses into duplicates beyond ~1.6 M particles, and the stable argsort
    that pairs the N-GPU result with the single-GPU one then pairs different particles: the 96^3 cloud of an 8-rank check did.)"""
    if n0 > (1 << 23):
        raise ValueError("identity tags cover 2^23 particles")
    return ((1 << 23) + np.arange(n0, dtype=np.int64)).astype(np.float32) * np.float32(2.0 ** np.round(np.log2(m0 / (1 << 23))))


def check_cloud(world):
    """(s, G): cells per side of the check's cube and of its domain for a world size — every x-slab at least 12 cells = 3 blocks wide"""
    s = max(24, 12 * world)
    G = 64
    while G < s + 16:
        G *= 2
    return s, G


def multi_gpu_parity(s=None, G=None, steps=6, rebin_every=3, migrate=False, e2e=False, transport="auto", rtol=5e-5, graph=False):
    """-> dict(ok, world, steps, max_err per attribute, max_vel_err, transport, shared_blocks_rank0, migrated)

    The cloud grows with the world size: every x-slab is at least 12 cells = 3 blocks wide (24^3 cells up to 2 ranks, 48^3 at 4, 96^3 at
    8), so that a block and its ring are shared with at most ZPCB200_HALO_K = 4 other ranks — the fused halo's documented limit; thinner
    slabs raise ZPC_HALO_TOO_MANY_PEERS (a 24^3 cloud on 8 ranks did).  dt scales with dx: the same CFL numbers and the same motion in cells
    per substep at every size."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if s is None or G is None:
        s, G = check_cloud(world)
    full = synth.elastic_cube(s, G, jitter_F=0.03, jitter_C=0.3)
    full["v"] *= 6.0                                  # particles cross cells, blocks and the slab cut
    n0 = full["m"].shape[0]
    full["m"] = identity_masses(n0, float(full["m"].mean()))
    c0, c1 = synth.slab_cell_range(s, rank, world)
    P = {k: (np.ascontiguousarray(v[8 * c0:8 * c1]) if isinstance(v, np.ndarray) else v) for k, v in full.items()}
    dt = synth.DT * 10 * 64.0 / G
    # blocks of a slab with the partition's extra ring (EnlargeSparsity{-1,3}): thin slabs hold many more blocks per particle than
    # the solver's default capacity (n / 256) assumes
    w_cells = (c1 - c0) // (s * s)
    eb = max((w_cells // 4 + 6) * (s // 4 + 6) ** 2, 1024)
    sol = DistMpmSolver(P, P["dx"], P["volume"], dt, synth.GRAVITY, mode=1, rebin_every=rebin_every, transport=transport,
                        layout="aos" if e2e else "binned", expected_blocks=eb)
    if e2e:
        hin = {k: torch.from_numpy(P[k].copy()).pin_memory() for k in ("x", "v", "m", "C", "F")}
        hout = {k: torch.empty_like(hin[k]).pin_memory() for k in ("x", "v", "C", "F")}
    ownership, moved = None, 0
    if migrate:
        from .dist_solver import BlockOwnership, shard_by_blocks
        _, cuts, keys = shard_by_blocks(full["x"], full["dx"], world)
        ownership = BlockOwnership(keys, cuts)
    if graph:   # the same substeps as CUDA graph replays (DistMpmSolver.capture_cycle): eager up to a cycle boundary, then one replay
        k = sol.capture_cycle()
        sol.replay_cycle()
        torch.cuda.synchronize()
        steps = sol.local.step_no
    for i in range(0 if graph else steps):
        if ownership is not None and i == (steps // 2 // rebin_every) * rebin_every and i > 0:     # at a re-bin boundary
            moved = sol.migrate(ownership)
        if e2e:
            sol.substep_host(hin, hout)
            torch.cuda.synchronize()
            for k in ("x", "v", "C", "F"):
                hin[k], hout[k] = hout[k], hin[k]
        else:
            sol.substep()
    torch.cuda.synchronize()
    mine = {k: hin[k].numpy() for k in ("x", "v", "m", "C", "F")} if e2e else sol.local.particles_host()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)       # only rank 0 compares
    mx = float(sol.max_vel_sqr().item())
    moved_all = torch.tensor([moved], device="cuda", dtype=torch.int64)
    dist.all_reduce(moved_all)
    out = dict(ok=True, world=world, steps=steps, transport=sol.transport, migrated=int(moved_all.item()) if migrate else None,
               path="substep_host (AoS)" if e2e else "substep (binned)" + (", CUDA graph replay" if graph else ""))
    if rank == 0:
        got = {k: np.concatenate([g[k] for g in gathered]) for k in ("x", "v", "m", "C", "F")}
        one = MpmSolver(full, full["dx"], full["volume"], dt, synth.GRAVITY, mode=1, layout="binned", rebin_every=rebin_every,
                        partition="with_rebin", expected_blocks=max((s // 4 + 6) ** 3, 1024))
        for _ in range(steps):
            one.substep()
        want = one.particles_host()
        og, ow = np.argsort(got["m"], kind="stable"), np.argsort(want["m"], kind="stable")
        vmax = float(np.abs(want["v"]).max())
        floors = dict(x=float(np.abs(want["x"]).max()), v=vmax, C=4.0 / full["dx"] * vmax, F=float(np.abs(want["F"]).max()))
        errs = {k: _rel(got[k][og], want[k][ow], floors[k]) for k in "xvCF"}
        mv = abs(mx - float(one.max_vel_sqr.item())) / max(mx, 1e-30)
        out.update(max_err=errs, max_vel_err=mv, shared_blocks_rank0=sol.halo.shared_blocks(), particles=n0,
                   ok=bool(got["m"].shape[0] == n0 and max(errs.values()) <= rtol and mv <= 1e-4))
        del one
    flag = torch.tensor([1 if out["ok"] else 0], device="cuda")
    dist.broadcast(flag, 0)
    out["ok"] = bool(flag.item() == 1)
    del sol
    torch.cuda.empty_cache()
    return out
