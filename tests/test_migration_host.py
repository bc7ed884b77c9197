"""Particle migration between ranks on CPU tensors over gloo: ownership of a block is a function of its key
(BlockOwnership over the cuts of shard_by_blocks), migrate_particles is one all_to_all of counts and one of 100-byte
records.  Checks: nothing is lost or duplicated, every particle ends on the rank that owns its home block, attributes travel
with their particle, a second migration without motion moves nothing."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zpc_b200 import synth
        from zpc_b200.dist_solver import BlockOwnership, migrate_particles, shard_by_blocks
        P = synth.elastic_cube(12, 32, jitter_C=0.3, jitter_F=0.02, shuffle_seed=8)
        n, dx = P["x"].shape[0], P["dx"]
        P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n) / n)).astype(np.float32)          # unique masses = identity
        owner, cuts, keys = shard_by_blocks(P["x"], dx, world)
        own = BlockOwnership(keys, cuts)
        # ownership as a function of position agrees with the sharding it came from
        o2 = own.owner_of_positions(torch.from_numpy(P["x"]), dx).numpy()
        assert np.array_equal(o2, owner)
        mine = owner == rank
        attrs = {k: torch.from_numpy(np.ascontiguousarray(P[k][mine])) for k in ("x", "v", "m", "C", "F")}
        # the cloud drifts diagonally by ~1.5 blocks: many particles change owner, some enter blocks nobody listed before
        attrs["x"] = attrs["x"] + torch.tensor([6.3 * dx, -2.1 * dx, 0.7 * dx])
        dest = own.owner_of_positions(attrs["x"], dx)
        n_leave = int((dest != rank).sum())
        new = migrate_particles(attrs, dest)
        assert all(new[k].shape[0] == new["x"].shape[0] for k in new) and new["m"].dim() == 1 and new["C"].shape[1] == 9
        assert bool((own.owner_of_positions(new["x"], dx) == rank).all())
        # a second pass moves nothing
        again = migrate_particles(new, own.owner_of_positions(new["x"], dx))
        assert all(torch.equal(again[k], new[k]) for k in new)
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: v.numpy() for k, v in new.items()})
        leave = [None] * world
        dist.all_gather_object(leave, n_leave)
        if rank == 0:
            allm = np.concatenate([g["m"] for g in gathered])
            assert allm.size == n and np.unique(allm).size == n                          # nothing lost, nothing duplicated
            o = np.argsort(allm, kind="stable")
            ref_o = np.argsort(P["m"], kind="stable")
            shift = np.float32([6.3 * dx, -2.1 * dx, 0.7 * dx])
            for k in ("v", "C", "F"):
                assert np.array_equal(np.concatenate([g[k] for g in gathered])[o], P[k][ref_o]), k
            assert np.array_equal(np.concatenate([g["x"] for g in gathered])[o], (torch.from_numpy(P["x"][ref_o]) + torch.from_numpy(shift)).numpy())
            q.put((sum(leave), [g["m"].size for g in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_migration_conserves_particles_and_respects_ownership(world):
    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    moved, sizes = q.get(timeout=10)
    assert moved > 100 and sum(sizes) == 8 * 12 ** 3
