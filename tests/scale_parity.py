"""TEST INFRASTRUCTURE: one substep of this library (AoS drop-in path, binned fast path on every sweep variant) against the
reference's OWN implementations on the same input at BASELINE sizes — by block key, per channel, max and 99.9th percentile:

  * the reference CUDA path   (oracle/_ref/libzpcref_cuda.so through oracle/refcuda_runner.py, a process of its own): the primary
    GPU oracle of SURVEY §8(c);
  * the reference OpenMP path (oracle/_ref/libzpcref.so, the unmodified reference compiled by oracle/Makefile).

Used by tests/test_gpu_scale_parity.py (asserts) and by `python -m tests.scale_parity --size C2` (prints the table kept under
profiles/).  Error measure = SURVEY §8(c): |a - b| / max(|a|, |b|, s), s = 1e-3 x the channel's max-abs ("strict"), and the same
with s = the channel's max-abs ("scale").
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from zpc_b200 import synth  # noqa: E402

E, NU = synth.MODEL["E"], synth.MODEL["nu"]
GRID_CH = ["m", "mv_x", "mv_y", "mv_z", "rhs_x", "rhs_y", "rhs_z"]


def _err(a, b, scale_floor):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale_floor)


def channel_errors(got, want, names, floors=None):
    """got, want: [..., nch] arrays (channel last).  -> {name: dict(scale, max, p999, strict_max, strict_p999)}"""
    out = {}
    for c, name in enumerate(names):
        a, b = got[..., c].ravel(), want[..., c].ravel()
        scale = float(max(np.abs(a).max(), np.abs(b).max()))
        if floors is not None:
            scale = max(scale, float(floors[c]))
        if scale == 0.0:
            out[name] = dict(scale=0.0, max=0.0, p999=0.0, strict_max=0.0, strict_p999=0.0)
            continue
        e, es = _err(a, b, scale), _err(a, b, 1e-3 * scale)
        out[name] = dict(scale=scale, max=float(e.max()), p999=float(np.quantile(e, 0.999)), strict_max=float(es.max()),
                         strict_p999=float(np.quantile(es, 0.999)))
    return out


def particle_errors(got, want, dx):
    vmax = float(np.abs(want["v"]).max())
    floors = dict(x=float(np.abs(want["x"]).max()), v=vmax, C=4.0 / dx * vmax, F=float(np.abs(want["F"]).max()))
    out = {}
    for k in "xvCF":
        e = _err(got[k], want[k], floors[k])
        out[k] = dict(scale=floors[k], max=float(e.max()), p999=float(np.quantile(e, 0.999)))
    return out


def make_input(s, G, jitter=True, seed=7):
    kw = dict(jitter_F=0.03, jitter_C=0.3) if jitter else {}
    return synth.elastic_cube(s, G, seed=seed, **kw)


def reference_cuda(P, mode=1, timeout=1800):
    """one substep on the reference's own CUDA functors, in a process of its own"""
    d = tempfile.mkdtemp(prefix="zpc_refcuda_")
    fin, fout = os.path.join(d, "in.npz"), os.path.join(d, "out.npz")
    np.savez(fin, dt=synth.DT, E=E, nu=NU, gravity=synth.GRAVITY, mode=mode, **P)
    r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "substep", fin, fout], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("reference CUDA path failed: " + (r.stdout + r.stderr)[-2000:])
    z = dict(np.load(fout))
    for f in (fin, fout):
        os.remove(f)
    os.rmdir(d)
    return z


def reference_omp(P, mode=1, threads=None):
    """one substep on the reference's OpenMP policy (oracle/_ref/libzpcref.so)"""
    from oracle.pyoracle import Ref
    ref = Ref()
    n = P["x"].shape[0]
    sim = ref.mpm(n, P["dx"], nthreads=threads or ref.max_threads(), expected_blocks=max(n // 8, 64))
    try:
        sim.set_particles(P)
        sim.partition()
        sim.clean_grid()
        sim.p2g(synth.DT, E, NU, P["volume"])
        keys = sim.keys()
        g1 = sim.grid()
        mx = sim.grid_update(synth.DT, synth.GRAVITY, mode)
        g2 = sim.grid()
        sim.g2p(synth.DT)
        Q = sim.get_particles()
    finally:
        sim.close()
    return dict(active_keys=keys, grid_p2g=g1, grid_upd=g2, max_vel_sqr=np.float32(mx), nblocks=keys.shape[0], **{k: Q[k] for k in "xvCF"})


def ours(P, variant, mode=1):
    """variant: "aos" | 3 | 4 | 6 (binned P2G sweep).  -> grids by key after P2G / update, particles in the input order"""
    import torch
    from tests.parity import grid_by_key
    from zpc_b200 import api
    n, dx = P["x"].shape[0], P["dx"]
    pars = api.Particles(P)
    table = api.HashTable(max(n // 8, 64))
    api.partition_for_particles(api.vec3_port(pars.x), n, dx, table)
    torch.cuda.synchronize()
    assert table.overflow.item() == 0
    nb = int(table.cnt.item())
    keys = table.active_keys[:nb].cpu().numpy()
    grids = api.Grids(dx, nb)
    api.clean_grid_blocks(grids, table)
    model = api.model_fcr(P["volume"], E, NU)
    if variant == "aos":
        src = pars
    else:
        api.set_tuning(int(variant), 1)
        src = api.ParticleBins(n, max(nb * 2, 64))
        order = torch.empty(n, dtype=torch.int32, device="cuda")
        api.bin_particles(pars, table, dx, src, order)
    try:
        api.p2g_transfer(src, table, grids, synth.DT, model)
        torch.cuda.synchronize()
        _, g1 = grid_by_key(keys, grids.tiles[:nb].cpu().numpy())
        mx = torch.zeros(1, device="cuda")
        api.compute_grid_block_velocity(grids, table, synth.DT, (0.0, synth.GRAVITY, 0.0), mode, mx)
        ks, g2 = grid_by_key(keys, grids.tiles[:nb].cpu().numpy())
        api.g2p_transfer(src, table, grids, synth.DT)
        torch.cuda.synchronize()
    finally:
        if variant != "aos":
            api.set_tuning(4, 1)
    if variant == "aos":
        Q = pars.to_host()
    else:
        perm = order.cpu().numpy()
        Q = {}
        for k in "xvCF":
            a = np.empty_like(P[k])
            a[perm] = src.attr(k).cpu().numpy()
            Q[k] = a
    return dict(keys=ks, grid_p2g=g1, grid_upd=g2, max_vel_sqr=float(mx.item()), **Q)


def compare(ref, got, dx):
    from tests.parity import grid_by_key
    kr, g1r = grid_by_key(ref["active_keys"], ref["grid_p2g"])
    _, g2r = grid_by_key(ref["active_keys"], ref["grid_upd"])
    assert np.array_equal(kr, got["keys"]), "block sets differ"
    out = dict(p2g=channel_errors(np.moveaxis(got["grid_p2g"], 1, -1), np.moveaxis(g1r, 1, -1), GRID_CH),
               update=channel_errors(np.moveaxis(got["grid_upd"][:, 1:4], 1, -1), np.moveaxis(g2r[:, 1:4], 1, -1), ["v_x", "v_y", "v_z"]),
               g2p=particle_errors(got, ref, dx),
               max_vel_sqr=abs(got["max_vel_sqr"] - float(ref["max_vel_sqr"])) / max(float(ref["max_vel_sqr"]), 1e-30))
    return out


def table_md(title, res):
    lines = ["### " + title, "", "| stage | channel | scale | max | 99.9 pct | strict max | strict 99.9 pct |", "|---|---|---|---|---|---|---|"]
    for stage in ("p2g", "update", "g2p"):
        for ch, e in res[stage].items():
            lines.append("| %s | %s | %.3e | %.2e | %.2e | %s | %s |" % (stage, ch, e["scale"], e["max"], e["p999"],
                                                                          "%.2e" % e["strict_max"] if "strict_max" in e else "-",
                                                                          "%.2e" % e["strict_p999"] if "strict_p999" in e else "-"))
    lines.append("| update | max_vel_sqr | - | %.2e | - | - | - |" % res["max_vel_sqr"])
    return "\n".join(lines) + "\n"


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="C2", help="C1 | C2 | C3 | <s>x<G>")
    ap.add_argument("--no-jitter", action="store_true", help="the bench workload itself (F = I, C = 0): the stress channels are ~0")
    ap.add_argument("--variants", default="aos,4,6")
    ap.add_argument("--refs", default="cuda,omp")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    G, s = synth.CONFIGS[a.size] if a.size in synth.CONFIGS else tuple(int(v) for v in a.size.split("x"))[::-1]
    P = make_input(s, G, jitter=not a.no_jitter)
    n = P["x"].shape[0]
    refs = {}
    if "cuda" in a.refs:
        refs["reference CUDA path (cuda_exec, libzpcref_cuda.so)"] = reference_cuda(P)
    if "omp" in a.refs:
        refs["reference OpenMP path (omp_exec, libzpcref.so)"] = reference_omp(P)
    md = ["# Achieved error of one substep vs the reference's own implementations: %d particles, %d^3 cells in a %d^3 domain%s" %
          (n, s, G, "" if a.no_jitter else ", F = I + U(-0.03, 0.03), C = U(-0.3, 0.3), v jittered"), "",
          "Error = |a - b| / max(|a|, |b|, scale) per entry; `strict` uses 1e-3 x scale as the floor (SURVEY §8(c)).  Grids compared by block key.", ""]
    js = {}
    for v in a.variants.split(","):
        got = ours(P, v if v == "aos" else int(v))
        for rname, ref in refs.items():
            res = compare(ref, got, P["dx"])
            title = "%s vs %s" % ("AoS drop-in kernels" if v == "aos" else "binned fast path, sweep %s" % v, rname)
            md.append(table_md(title, res))
            js[title] = res
    if len(refs) == 2:
        (n0, r0), (n1, r1) = list(refs.items())
        from tests.parity import grid_by_key
        k1, g1 = grid_by_key(r1["active_keys"], r1["grid_p2g"])
        _, g2 = grid_by_key(r1["active_keys"], r1["grid_upd"])
        res = compare(r0, dict(keys=k1, grid_p2g=g1, grid_upd=g2, max_vel_sqr=float(r1["max_vel_sqr"]), **{k: r1[k] for k in "xvCF"}), P["dx"])
        title = "the reference against itself: %s vs %s" % (n1, n0)
        md.append(table_md(title, res))
        js[title] = res
    text = "\n".join(md)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
        with open(os.path.splitext(a.out)[0] + ".json", "w") as f:
            json.dump(js, f, indent=1)


if __name__ == "__main__":
    main()
