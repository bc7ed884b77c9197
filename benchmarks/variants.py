#!/usr/bin/env python
"""Times the kernel variants of the binned P2G / G2P (zpcb200_set_tuning) on one resident workload.

  python benchmarks/variants.py [--config C3] [--steps 8] [--warmup 3]

One solver is built once; for every (p2g_sweep, g2p_staged) combination the same substeps are replayed with per-stage
CUDA events.  Prints one JSON line per combination (ms per stage per substep, algorithmic GB/s, fraction of HBM peak)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--combos", default="3:0,4:0,4:1,5:1,4:64,4:256", help="comma list of p2g_sweep:g2p_staged")
    ap.add_argument("--tag", default=os.environ.get("ZPCB200_LIB", ""))
    args = ap.parse_args()
    import torch
    from bench import BYTES_PER_PARTICLE, peaks
    from zpc_b200 import api, synth
    from zpc_b200.solver import MpmSolver
    G, s = synth.CONFIGS[args.config]
    P = synth.elastic_cube(s, G)
    hbm, _ = peaks()
    n = P["x"].shape[0]
    for sweep, staged in [tuple(int(x) for x in c.split(":")) for c in args.combos.split(",")]:
        api.set_tuning(sweep, staged)
        # a fresh solver per combination: every variant sees the same particle state (no accumulated drift)
        sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, layout="binned", rebin_every=0, partition="with_rebin")
        for _ in range(args.warmup):
            sol.substep()
        torch.cuda.synchronize()
        sol.stage_events = []
        for _ in range(args.steps):
            sol.substep()
        torch.cuda.synchronize()
        st = {k: v / args.steps for k, v in sol.stage_times_ms().items()}
        sol.stage_events = None
        out = dict(tag=args.tag, config=args.config, n=n, p2g_sweep=sweep, g2p_staged=staged, ms=st)
        for k in ("p2g", "g2p"):
            gbs = BYTES_PER_PARTICLE[k] * n / (st[k] * 1e-3) / 1e9
            out[k + "_gbps"] = gbs
            out[k + "_frac"] = gbs / hbm
        fused = sum(st[k] for k in ("clean", "p2g", "grid_update", "g2p"))
        out["fused_ms"] = fused
        out["fused_frac"] = 257.5 * n / (fused * 1e-3) / 1e9 / hbm
        print(json.dumps(out), flush=True)
        del sol
        torch.cuda.empty_cache()
    api.set_tuning(4, 1)


if __name__ == "__main__":
    main()
