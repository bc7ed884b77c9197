#!/usr/bin/env python
"""SgMpmSolver (SparseGrid<3,f32,8>, block-binned fast path): parity of a few substeps against MpmSolver on the legacy grid (same
library, same particles: particle for particle), then timed substeps at a BASELINE config.  One JSON line.
  python benchmarks/sg_solver_bench.py [--config C3] [--steps 16] [--warmup 4]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    args = ap.parse_args()
    import torch
    from bench import peaks
    from zpc_b200 import synth
    from zpc_b200.sg_solver import SgMpmSolver
    from zpc_b200.solver import MpmSolver
    out = {}
    # ---- parity at a small size: 12 substeps with re-bins every 2 (0.4 cells per substep: inside the one-cell reach of a bin between re-bins), particles move across cells and blocks
    P = synth.elastic_cube(20, 64, jitter_F=0.03, jitter_C=0.3, shuffle_seed=3)
    P["v"] *= 6.0
    n0 = P["m"].shape[0]
    from zpc_b200.selfcheck import identity_masses
    P["m"] = identity_masses(n0, float(P["m"].mean()))
    dt = synth.DT * 10
    a = SgMpmSolver(P, P["dx"], P["volume"], dt, synth.GRAVITY, rebin_every=2)
    b = MpmSolver(P, P["dx"], P["volume"], dt, synth.GRAVITY, mode=1, layout="binned", rebin_every=2, partition="with_rebin")
    for _ in range(12):
        a.substep(); b.substep()
    torch.cuda.synchronize()
    ga, gb = a.particles_host(), b.particles_host()
    oa, ob = np.argsort(ga["m"], kind="stable"), np.argsort(gb["m"], kind="stable")
    vmax = float(np.abs(gb["v"]).max())
    floors = dict(x=float(np.abs(gb["x"]).max()), v=vmax, C=4.0 / P["dx"] * vmax, F=float(np.abs(gb["F"]).max()))
    errs = {k: float((np.abs(ga[k][oa].astype(np.float64) - gb[k][ob]) / np.maximum(np.maximum(np.abs(ga[k][oa]), np.abs(gb[k][ob])), floors[k])).max())
            for k in "xvCF"}
    out["parity_vs_legacy_grid_solver"] = dict(particles=n0, substeps=12, rebin_every=2, max_err=errs, ok=bool(max(errs.values()) <= 5e-5),
                                               max_vel_sqr=[float(a.max_vel_sqr.item()), float(b.max_vel_sqr.item())])
    del a, b
    torch.cuda.empty_cache()
    # ---- timing
    G, s = synth.CONFIGS[args.config]
    P = synth.elastic_cube(s, G)
    n = P["x"].shape[0]
    hbm, _ = peaks()
    sol = SgMpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, rebin_every=8, expected_blocks=max(2 * n // 4096, 512) * 2)
    for _ in range(args.warmup):
        sol.substep()
    torch.cuda.synchronize()
    sol.stage_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sol.substep()
    e1.record()
    torch.cuda.synchronize()
    st = {k: v / args.steps for k, v in sol.stage_times_ms().items()}
    ms = e0.elapsed_time(e1) / args.steps
    fused = sum(st.get(k, 0.0) for k in ("clean", "p2g", "grid_update", "g2p"))
    out.update(config=args.config, n=n, active_blocks=sol.sg.table.size(), ms_per_step=ms, value=n / (ms * 1e-3), stage_ms=st,
               fused_ms=fused, fused_frac=257.5 * n / (fused * 1e-3) / 1e9 / hbm)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
