"""N>1 host logic on CPU: world_size-2 and -4 gloo runs of the slab sharding + one-ring grid-block halo exchange
(zpc_b200/dist_solver.py::HaloExchange).  The per-rank P2G is computed with the oracle (this is a test of the
exchange plumbing, not of the kernels); pack / unpack are the torch-indexing stand-ins defined HERE, the product
uses the C-ABI kernels zpcb200_halo_pack / zpcb200_halo_unpack_add (covered by tests/test_gpu_mpm.py)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CpuGrids:
    nch = 7

    def __init__(self, tiles):
        self.tiles = tiles


def cpu_pack(grids, ids, buf):
    buf.copy_(grids.tiles[ids.long()])


def cpu_unpack_add(grids, ids, buf):
    grids.tiles[ids.long()] += buf


def sorted_table(oracle, x, dx, enlarge_extra):
    """oracle partition with rows renumbered in ascending key order (what zpcb200_partition_build produces)"""
    n = x.shape[0]
    tab = oracle.partition_build(x, dx, oracle.table_size_for(max(n // 8, 64)))
    keys = tab["active_keys"]
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    rank_of = np.empty(len(order), np.int32)
    rank_of[order] = np.arange(len(order), dtype=np.int32)
    occ = tab["indices"] >= 0
    tab["indices"][occ] = rank_of[tab["indices"][occ]]
    tab["active_keys"] = np.ascontiguousarray(keys[order])
    return tab


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        from zpc_b200 import synth
        from zpc_b200.dist_solver import HaloExchange, pack_keys
        o = Oracle()
        s, G = (8, 32) if world == 2 else (16, 32)
        full = synth.elastic_cube(s, G, jitter_F=0.04, jitter_C=0.4)
        c0, c1 = synth.slab_cell_range(s, rank, world)
        mine = {k: (np.ascontiguousarray(v[8 * c0:8 * c1]) if isinstance(v, np.ndarray) else v) for k, v in full.items()}
        part = synth.elastic_cube_slab(s, G, rank, world, jitter_F=0.0)
        assert np.array_equal(part["x"], mine["x"])                       # the shard generator reproduces the slice
        dx = full["dx"]
        tab = sorted_table(o, mine["x"], dx, 0)
        grid = o.p2g(mine, tab, dx, synth.DT, 5e4, 0.4, full["volume"])
        tiles = torch.from_numpy(grid.copy())
        halo = HaloExchange(None, 7, "cpu", cpu_pack, cpu_unpack_add)
        peers = halo.build(torch.from_numpy(tab["active_keys"]))
        if world == 2:
            assert len(peers) == 1 and peers[0][0] == 1 - rank
        else:
            # 4-cell slabs that do not line up with the 4-cell blocks: a rank shares blocks with its neighbours AND, through
            # the one-ring of the stencil, with ranks two slabs away
            assert [p[0] for p in peers] == sorted(p[0] for p in peers) and rank not in [p[0] for p in peers]
            assert {rank - 1, rank + 1} & set(range(world)) <= {p[0] for p in peers}
        # every pair of ranks lists the same shared keys in the same (ascending) order
        mykeys = torch.from_numpy(tab["active_keys"])
        per_peer = {}
        for q_, ids, _, _ in peers:
            codes = pack_keys(mykeys[ids.long()])
            assert bool((codes[1:] > codes[:-1]).all())
            per_peer[q_] = codes
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: v.tolist() for k, v in per_peer.items()})
        for q_, codes in per_peer.items():
            assert gathered[q_][rank] == codes.tolist()
        shared_keys = mykeys[torch.cat([ids for _, ids, _, _ in peers]).long().unique()]
        halo.exchange_add(CpuGrids(tiles))
        mx = torch.tensor([float(rank + 1)])
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        assert mx.item() == world
        # against the single-domain oracle P2G, block by block (by key)
        tab_f = sorted_table(o, full["x"], dx, 0)
        g_full = o.p2g(full, tab_f, dx, synth.DT, 5e4, 0.4, full["volume"])
        full_row = {tuple(k): i for i, k in enumerate(tab_f["active_keys"].tolist())}
        shared = set(map(tuple, shared_keys.tolist()))
        worst = 0.0
        scale = np.abs(g_full).max(axis=(0, 2), keepdims=True)[0]
        n_checked = 0
        for i, k in enumerate(map(tuple, tab["active_keys"].tolist())):
            ref_tile = g_full[full_row[k]]
            got = tiles[i].numpy()
            # a block that is NOT shared belongs to this rank's particles only -> must already be complete
            err = np.abs(got - ref_tile) / np.maximum(scale, 1e-30)
            worst = max(worst, float(err.max()))
            n_checked += 1
        assert worst < 1e-5, worst
        q.put((rank, len(shared), n_checked, worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_matches_single_domain(world):
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
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert all(r[1] > 0 for r in res)          # every rank shares blocks with somebody
    if world == 2:
        assert res[0][1] == res[1][1]          # the same number on both sides
