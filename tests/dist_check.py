"""torchrun --nproc-per-node N tests/dist_check.py : N-GPU sharded substeps against the single-GPU solver on
the same cloud (rank 0 holds both).  Exit code 0 = parity within the multi-step tolerance.
ZPC_MIGRATE=1: after half of the substeps every particle is handed to the rank that owns its current home block
(DistMpmSolver.migrate; ownership = BlockOwnership over shard_by_blocks of the initial cloud) — results must not change.
ZPC_E2E=1: the substeps go through DistMpmSolver.substep_host (host buffers per rank, AoS kernels) instead."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.parity import check_particles  # noqa: E402
from zpc_b200 import synth  # noqa: E402
from zpc_b200.dist_solver import DistMpmSolver  # noqa: E402
from zpc_b200.solver import MpmSolver  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    s, G, steps = 24, 64, 6
    kw = dict(jitter_F=0.03, jitter_C=0.3)
    P = synth.elastic_cube_slab(s, G, rank, world)   # jitter needs the global stream: apply it from the full cloud below
    full = synth.elastic_cube(s, G, **kw)
    full["v"] *= 6.0                                  # particles cross cells, blocks and the slab cut
    n0 = full["m"].shape[0]
    full["m"] = (full["m"] * (1.0 + 0.1 * np.arange(n0) / n0)).astype(np.float32)   # identity tag
    c0, c1 = synth.slab_cell_range(s, rank, world)
    P = {k: (np.ascontiguousarray(v[8 * c0:8 * c1]) if isinstance(v, np.ndarray) else v) for k, v in full.items()}
    e2e = os.environ.get("ZPC_E2E") == "1"
    sol = DistMpmSolver(P, P["dx"], P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, rebin_every=3,
                        transport=os.environ.get("ZPC_HALO", "auto"), layout="aos" if e2e else "binned")
    if e2e:
        hin = {k: torch.from_numpy(P[k].copy()).pin_memory() for k in ("x", "v", "m", "C", "F")}
        hout = {k: torch.empty_like(hin[k]).pin_memory() for k in ("x", "v", "C", "F")}
    ownership = None
    if os.environ.get("ZPC_MIGRATE") == "1":
        from zpc_b200.dist_solver import BlockOwnership, shard_by_blocks
        _, cuts, keys = shard_by_blocks(full["x"], full["dx"], world)
        ownership = BlockOwnership(keys, cuts)
    for i in range(steps):
        if ownership is not None and i == steps // 2:       # a re-bin boundary (rebin_every = 3, steps = 6)
            moved = sol.migrate(ownership)
            print("rank %d: migrated %d particles away, now holds %d" % (rank, moved, sol.n))
        if e2e:
            sol.substep_host(hin, hout)
            torch.cuda.synchronize()
            for k in ("x", "v", "C", "F"):
                hin[k], hout[k] = hout[k], hin[k]
        else:
            sol.substep()
    torch.cuda.synchronize()
    mine = {k: hin[k].numpy() for k in ("x", "v", "m", "C", "F")} if e2e else sol.local.particles_host()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    mx = float(sol.max_vel_sqr().item())
    ok = True
    if rank == 0:
        got = {k: np.concatenate([g[k] for g in gathered]) for k in ("x", "v", "m", "C", "F")}
        one = MpmSolver(full, full["dx"], full["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout="binned", rebin_every=3,
                        partition="with_rebin")
        for _ in range(steps):
            one.substep()
        want = one.particles_host()

        def canon(Q):
            o = np.argsort(Q["m"], kind="stable")
            return {k: Q[k][o] for k in "xvCF"}
        try:
            check_particles(canon(got), canon(want), full["dx"], "%d-GPU vs 1-GPU (%d substeps)" % (world, steps), rtol=5e-5)
            assert abs(mx - float(one.max_vel_sqr.item())) <= 1e-4 * mx
            print("dist_check ok: world %d, transport %s, shared blocks on rank 0: %d" % (world, sol.transport, sol.halo.shared_blocks()))
        except AssertionError as e:
            print("dist_check FAILED:", e)
            ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
