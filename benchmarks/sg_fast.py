#!/usr/bin/env python
"""SparseGrid<3,f32,8>: the block-binned fast path (bins = octants of the side-8 blocks) against the any-order kernels on the
same cloud.  python benchmarks/sg_fast.py [--config C2] [--steps 8]   -> one JSON line (ms per stage per substep)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    from bench import BYTES_PER_PARTICLE, peaks
    from zpc_b200 import api, synth
    G, s = synth.CONFIGS[args.config]
    P = synth.elastic_cube(s, G)
    n, dx = P["x"].shape[0], P["dx"]
    hbm, _ = peaks()
    pars = api.Particles(P)
    sg = api.SparseGrid(7, max(n // 2048, 64) * 2)
    sg.scale(dx)
    api.sg_partition_for_particles(api.vec3_port(pars.x), n, sg)
    torch.cuda.synchronize()
    assert sg.table.overflow.item() == 0
    nb = sg.table.size()
    model = api.model_fcr(P["volume"], synth.MODEL["E"], synth.MODEL["nu"])
    mx = torch.zeros(1, device="cuda")
    bins = api.ParticleBins(n, 8 * nb + 64)
    api.sg_bin_particles(pars, sg, bins)
    torch.cuda.synchronize()
    assert int(bins.status.item()) == 0

    def run(target, steps):
        ev = []
        for _ in range(steps):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            e[0].record(); api.sg_clean(sg)
            e[1].record(); api.sg_p2g_transfer(target, sg, synth.DT, model)
            e[2].record(); api.sg_compute_grid_velocity(sg, synth.DT, (0.0, synth.GRAVITY, 0.0), 1, mx)
            e[3].record(); api.sg_g2p_transfer(target, sg, synth.DT)
            e[4].record()
            ev.append(e)
        torch.cuda.synchronize()
        names = ("clean", "p2g", "grid_update", "g2p")
        return {k: sum(e[i].elapsed_time(e[i + 1]) for e in ev) / len(ev) for i, k in enumerate(names)}

    out = dict(config=args.config, n=n, active_blocks=nb, bins=int(bins.num_bins.item()))
    for name, target in (("binned", bins), ("any_order", pars)):
        run(target, args.warmup)
        st = run(target, args.steps)
        out[name] = dict(ms=st, fused_ms=sum(st.values()),
                         p2g_frac=BYTES_PER_PARTICLE["p2g"] * n / (st["p2g"] * 1e-3) / 1e9 / hbm,
                         g2p_frac=BYTES_PER_PARTICLE["g2p"] * n / (st["g2p"] * 1e-3) / 1e9 / hbm)
    out["speedup"] = {k: out["any_order"]["ms"][k] / out["binned"]["ms"][k] for k in ("p2g", "g2p")}
    out["status"] = int(bins.status.item())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
