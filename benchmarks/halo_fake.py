"""Single-GPU stand-in for the fused halo kernels (timing / ncu only, results are meaningless): the maps declare a slab of blocks shared
with a 'rank 1' whose receive buffer is this GPU's own, so the P2G write-back issues its extra bulk reduce-adds and the grid update
runs its receive branch, without a second process.   python benchmarks/halo_fake.py [--config C2]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zpc_b200 import api, synth  # noqa: E402
from zpc_b200.solver import MpmSolver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--frac", type=float, default=0.25, help="fraction of the blocks declared shared")
    a = ap.parse_args()
    G, s = synth.CONFIGS[a.config]
    P = synth.elastic_cube(s, G)
    sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, layout="binned", rebin_every=0, partition="with_rebin")
    for _ in range(2):
        sol.substep()
    torch.cuda.synchronize()
    nb, cap, K = sol.table.size(), sol.block_cap, api.HALO_K
    seg = cap
    buf = torch.zeros(2 * 2 * seg * 448, dtype=torch.float32, device="cuda")
    peer = torch.full((cap, K), -1, dtype=torch.int32, device="cuda")
    pos = torch.zeros(cap, K, dtype=torch.int32, device="cuda")
    nshared = int(nb * a.frac)
    peer[:nshared, 0] = 1
    pos[:nshared, 0] = torch.arange(nshared, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ptrs = (C.c_void_p * api.HALO_MAX_PEERS)(*([buf.data_ptr(), buf.data_ptr()] + [None] * (api.HALO_MAX_PEERS - 2)))

    def view(on, half):
        return api.zpc_halo_view(peer.data_ptr() if on else None, pos.data_ptr(), 2, 1, seg, half, buf.data_ptr(), ptrs, status.data_ptr())
    L = sol
    res = {}
    for on in (False, True, False, True):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        tp = tu = 0.0
        for it in range(6):
            api.clean_grid_blocks(L.grids, L.table)
            ev[0].record()
            api.p2g_transfer_halo(L.bins, L.table, L.grids, L.dt, L.model, view(on, it & 1))
            ev[1].record()
            L.max_vel_sqr.zero_()
            ev[2].record()
            api.grid_update_halo(L.grids, L.table, L.dt, L.extf, L.mode, [], L.max_vel_sqr, view(on, it & 1))
            ev[3].record()
            torch.cuda.synchronize()
            if it >= 2:
                tp += ev[0].elapsed_time(ev[1]) / 4
                tu += ev[2].elapsed_time(ev[3]) / 4
        # the plain update for comparison
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        api.compute_grid_block_velocity(L.grids, L.table, L.dt, L.extf, L.mode, L.max_vel_sqr)
        e1.record()
        torch.cuda.synchronize()
        print("halo %s: p2g %.3f ms, grid_update_halo %.3f ms (plain grid_update %.3f ms), blocks %d shared %d" % (on, tp, tu, e0.elapsed_time(e1), nb, nshared if on else 0))


if __name__ == "__main__":
    main()
