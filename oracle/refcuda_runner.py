"""TEST / BASELINE INFRASTRUCTURE — NOT PRODUCT CODE.

Runs the reference's OWN CUDA path (oracle/_ref/libzpcref_cuda.so: the unmodified reference headers + CUDA backend compiled by
`make -C oracle refcuda`) in a process of its own — no torch, no libzpcb200 — so that a fault inside the reference cannot take a
test session or a bench run down with it.

  python -m oracle.refcuda_runner substep IN.npz OUT.npz   # x v m C F dx dt E nu volume gravity mode -> grids + particles
  python -m oracle.refcuda_runner bench G S STEPS WARMUP    # elastic cube of S^3 cells in a G^3 domain; prints one JSON line
  python -m oracle.refcuda_runner overlay IN.npz OUT.npz   # the same substep on the reference's containers through
                                                            # include/zpcb200/zs_overlay.cuh (b200_exec() + zs::b200::* launches)
  python -m oracle.refcuda_runner binned IN.npz OUT.npz     # IN also carries steps, rebin_every: the reference's functors for `steps`
                                                            # substeps vs the overlay's block-binned fast path (zs::b200::BinnedParticles)
  python -m oracle.refcuda_runner lbvh N OUT.npz            # the reference's own LBvh::build on cuda_exec() and on b200_exec()
  python -m oracle.refcuda_runner prims-bench LOG2N ITERS   # C5: the reference's CudaExecutionPolicy primitives (and the same generic
                                                            # calls on b200_exec()) timed on the GPU; one JSON line
  python -m oracle.refcuda_runner prims N OUT.npz           # zs::radix_sort_pair / exclusive_scan / reduce with b200_exec()
  python -m oracle.refcuda_runner g2p2g IN.npz OUT.npz      # the reference's own G2P2GTransfer (simulation/transfer/G2P2G.hpp:49-141) on
                                                            # cuda_exec(): IN carries the particles, model (0 fcr, 1 von Mises), prm;
                                                            # the grid velocity DOFs are node_field() of the node coordinates
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libzpcref_cuda.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def node_field(keys):
    """Grid velocity DOFs as a function of the node's integer coordinates only, so that two partitions with different block
    numberings (the reference's insertion order, ours by key rank) hold the same field: keys [nb, 3] block coordinates (the
    table's keys, SparsityOp.hpp:72-78: cell coordinate / 4) -> [nb * 64, 3] float32, cell id (x << 4) | (y << 2) | z inside a
    block (geometry/Structure.hpp:49-61)."""
    keys = np.asarray(keys, np.int64)
    c = np.arange(64)
    loc = np.stack([c >> 4, (c >> 2) & 3, c & 3], 1)                       # [64, 3]
    node = (keys[:, None, :] * 4 + loc[None, :, :]).reshape(-1, 3)            # [nb * 64, 3]
    out = np.empty((node.shape[0], 3), np.float32)
    for d in range(3):
        h = (node[:, 0] * 73856093 + node[:, 1] * 19349663 + node[:, 2] * 83492791 + d * 2654435761) & 0xFFFFFFFF
        h = (h ^ (h >> 15)) * 2246822519 & 0xFFFFFFFF
        h = (h ^ (h >> 13)) & 0xFFFFFF
        out[:, d] = (h.astype(np.float64) / float(0xFFFFFF) * 2.0 - 1.0).astype(np.float32)
    return out


class RefCuda:
    def __init__(self):
        self.L = C.CDLL(SO)
        self.L.zpcrefcuda_mpm_create.restype = C.c_void_p
        self.L.zpcrefcuda_mpm_partition.restype = C.c_int
        self.L.zpcrefcuda_mpm_grid_update.restype = C.c_float

    @staticmethod
    def available():
        return os.path.exists(SO)

    def substep(self, P, dt, E, nu, gravity, mode, steps=1, collect=True, expected_blocks=None, overlay=False):
        n, dx = P["x"].shape[0], float(P["dx"])
        L = self.L
        if overlay:   # same containers, this repository's kernels behind the reference-side binding
            class _O:
                pass
            O = _O()
            for name in ("partition", "clean_grid", "p2g", "grid_update", "g2p"):
                setattr(O, "zpcrefcuda_mpm_" + name, getattr(L, "zpcrefcuda_overlay_" + name))
            O.zpcrefcuda_mpm_partition.restype = C.c_int
            O.zpcrefcuda_mpm_grid_update.restype = C.c_float
            for name in ("create", "set_particles", "get_keys", "get_grid", "get_particles", "destroy"):
                setattr(O, "zpcrefcuda_mpm_" + name, getattr(L, "zpcrefcuda_mpm_" + name))
            O.zpcrefcuda_sync = L.zpcrefcuda_sync
            L = O
        h = C.c_void_p(L.zpcrefcuda_mpm_create(C.c_int(n), C.c_float(dx), C.c_int(expected_blocks or max(n // 8, 64))))
        L.zpcrefcuda_mpm_set_particles(h, _p(P["x"]), _p(P["v"]), _p(P["m"]), _p(P["C"]), _p(P["F"]))
        out = {}
        times = []
        stages = {k: [] for k in ("partition", "clean", "p2g", "grid_update", "g2p")}

        def timed(name, fn, *a):
            t = time.perf_counter()
            r_ = fn(*a)
            L.zpcrefcuda_sync()
            stages[name].append(time.perf_counter() - t)
            return r_
        for _ in range(steps):
            L.zpcrefcuda_sync()
            t0 = time.perf_counter()
            nb = timed("partition", L.zpcrefcuda_mpm_partition, h)
            timed("clean", L.zpcrefcuda_mpm_clean_grid, h)
            timed("p2g", L.zpcrefcuda_mpm_p2g, h, C.c_float(dt), C.c_float(E), C.c_float(nu), C.c_float(P["volume"]))
            if collect and steps == 1:
                keys = np.empty((nb, 3), np.int32); g1 = np.empty((nb, 7, 64), np.float32)
                L.zpcrefcuda_mpm_get_keys(h, _p(keys)); L.zpcrefcuda_mpm_get_grid(h, _p(g1))
                out.update(active_keys=keys, grid_p2g=g1)
            mx = timed("grid_update", L.zpcrefcuda_mpm_grid_update, h, C.c_float(dt), C.c_float(gravity), C.c_int(mode))
            if collect and steps == 1:
                g2 = np.empty((nb, 7, 64), np.float32)
                L.zpcrefcuda_mpm_get_grid(h, _p(g2))
                out.update(grid_upd=g2, max_vel_sqr=np.float32(mx), nblocks=nb)
            timed("g2p", L.zpcrefcuda_mpm_g2p, h, C.c_float(dt))
            times.append(time.perf_counter() - t0)
        out["stage_ms"] = {k: [1e3 * t for t in v] for k, v in stages.items()}
        if collect:
            x = np.empty((n, 3), np.float32); v = np.empty((n, 3), np.float32)
            Cm = np.empty((n, 9), np.float32); F = np.empty((n, 9), np.float32)
            L.zpcrefcuda_mpm_get_particles(h, _p(x), _p(v), _p(Cm), _p(F))
            out.update(x=x, v=v, C=Cm, F=F)
        L.zpcrefcuda_mpm_destroy(h)
        out["times"] = np.array(times)
        return out


def main(argv):
    sys.path.insert(0, os.path.dirname(HERE))
    r = RefCuda()
    if argv[0] == "substep":
        z = np.load(argv[1])
        P = {k: np.ascontiguousarray(z[k]) for k in ("x", "v", "m", "C", "F")}
        P["dx"], P["volume"] = float(z["dx"]), float(z["volume"])
        out = r.substep(P, float(z["dt"]), float(z["E"]), float(z["nu"]), float(z["gravity"]), int(z["mode"]))
        out.pop("stage_ms", None)
        np.savez(argv[2], **out)
    elif argv[0] == "g2p2g":
        z = np.load(argv[1])
        P = {k: np.ascontiguousarray(z[k]) for k in ("x", "v", "m", "C", "F")}
        n, dx = P["x"].shape[0], float(z["dx"])
        L = r.L
        L.zpcrefcuda_mpm_g2p2g.restype = C.c_int
        h = C.c_void_p(L.zpcrefcuda_mpm_create(C.c_int(n), C.c_float(dx), C.c_int(max(n // 8, 64))))
        L.zpcrefcuda_mpm_set_particles(h, _p(P["x"]), _p(P["v"]), _p(P["m"]), _p(P["C"]), _p(P["F"]))
        nb = L.zpcrefcuda_mpm_partition(h)
        keys = np.empty((nb, 3), np.int32)
        L.zpcrefcuda_mpm_get_keys(h, _p(keys))
        gridv = np.ascontiguousarray(node_field(keys))
        gridr = np.zeros((nb * 64, 3), np.float32)
        prm = np.ascontiguousarray(z["prm"], np.float32)
        rc = L.zpcrefcuda_mpm_g2p2g(h, C.c_int(int(z["model"])), _p(prm), C.c_float(float(z["volume"])), C.c_float(float(z["dt"])), _p(gridv), _p(gridr))
        assert rc == 0, rc
        L.zpcrefcuda_mpm_destroy(h)
        np.savez(argv[2], active_keys=keys, gridr=gridr.reshape(nb, 64, 3), nblocks=nb)
    elif argv[0] == "overlay":
        z = np.load(argv[1])
        P = {k: np.ascontiguousarray(z[k]) for k in ("x", "v", "m", "C", "F")}
        P["dx"], P["volume"] = float(z["dx"]), float(z["volume"])
        out = r.substep(P, float(z["dt"]), float(z["E"]), float(z["nu"]), float(z["gravity"]), int(z["mode"]), overlay=True)
        out.pop("stage_ms", None)
        np.savez(argv[2], **out)
    elif argv[0] == "binned":
        z = np.load(argv[1])
        P = {k: np.ascontiguousarray(z[k]) for k in ("x", "v", "m", "C", "F")}
        P["dx"], P["volume"] = float(z["dx"]), float(z["volume"])
        steps, every = int(z["steps"]), int(z["rebin_every"])
        ref = r.substep(P, float(z["dt"]), float(z["E"]), float(z["nu"]), float(z["gravity"]), 1, steps=steps, collect=True)
        n = P["x"].shape[0]
        L = r.L
        h = C.c_void_p(L.zpcrefcuda_mpm_create(C.c_int(n), C.c_float(P["dx"]), C.c_int(max(n // 8, 64))))
        L.zpcrefcuda_mpm_set_particles(h, _p(P["x"]), _p(P["v"]), _p(P["m"]), _p(P["C"]), _p(P["F"]))
        order = np.empty(n, np.int32)
        L.zpcrefcuda_overlay_binned_substeps(h, C.c_int(steps), C.c_int(every), C.c_float(float(z["dt"])), C.c_float(float(z["E"])),
                                             C.c_float(float(z["nu"])), C.c_float(P["volume"]), C.c_float(float(z["gravity"])), _p(order))
        x = np.empty((n, 3), np.float32); v = np.empty((n, 3), np.float32); Cm = np.empty((n, 9), np.float32); F = np.empty((n, 9), np.float32)
        m = np.empty(n, np.float32)
        L.zpcrefcuda_mpm_get_particles(h, _p(x), _p(v), _p(Cm), _p(F))
        L.zpcrefcuda_mpm_get_mass(h, _p(m))
        L.zpcrefcuda_mpm_destroy(h)
        np.savez(argv[2], ref_x=ref["x"], ref_v=ref["v"], ref_C=ref["C"], ref_F=ref["F"], ref_m=P["m"], x=x, v=v, C=Cm, F=F, m=m)
    elif argv[0] == "gridmom":   # GridAngularMomentum + GridMomentumToVelocity: the reference's functors on cuda_exec() vs the overlay
        z = np.load(argv[1])
        P = {k: np.ascontiguousarray(z[k]) for k in ("x", "v", "m", "C", "F")}
        P["dx"], P["volume"] = float(z["dx"]), float(z["volume"])
        n = P["x"].shape[0]
        L = r.L
        out = {}
        for tag, fn in (("ref", L.zpcrefcuda_mpm_grid_momentum), ("b200", L.zpcrefcuda_overlay_grid_momentum)):
            fn.restype = C.c_float
            h = C.c_void_p(L.zpcrefcuda_mpm_create(C.c_int(n), C.c_float(P["dx"]), C.c_int(max(n // 8, 64))))
            L.zpcrefcuda_mpm_set_particles(h, _p(P["x"]), _p(P["v"]), _p(P["m"]), _p(P["C"]), _p(P["F"]))
            nb = L.zpcrefcuda_mpm_partition(h)            # the reference's own partition and P2G for both: identical tables,
            L.zpcrefcuda_mpm_clean_grid(h)                # grids equal up to the order of the P2G atomics
            L.zpcrefcuda_mpm_p2g(h, C.c_float(float(z["dt"])), C.c_float(float(z["E"])), C.c_float(float(z["nu"])), C.c_float(P["volume"]))
            g0 = np.empty((nb, 7, 64), np.float32); keys = np.empty((nb, 3), np.int32)
            L.zpcrefcuda_mpm_get_grid(h, _p(g0)); L.zpcrefcuda_mpm_get_keys(h, _p(keys))
            sum6 = np.zeros(6, np.float64)
            mx = fn(h, _p(sum6))
            g1 = np.empty((nb, 7, 64), np.float32)
            L.zpcrefcuda_mpm_get_grid(h, _p(g1))
            L.zpcrefcuda_mpm_destroy(h)
            out.update({tag + "_grid": g0, tag + "_keys": keys, tag + "_sum6": sum6, tag + "_max": np.float32(mx), tag + "_vel": g1})
        np.savez(argv[2], **out)
    elif argv[0] == "lbvh":
        n = int(argv[1])
        rs = np.random.RandomState(77)
        c = rs.uniform(0, 1, (n, 3)).astype(np.float32); hw = rs.uniform(0.001, 0.03, (n, 3)).astype(np.float32)
        c[n // 3: 2 * n // 3] = c[n // 3]
        bvs = np.ascontiguousarray(np.concatenate([c - hw, c + hw], 1), np.float32)
        out = dict(bvs=bvs)
        r.L.zpcrefcuda_lbvh_build.restype = C.c_int
        for tag, use in (("cuda", 0), ("b200", 1)):
            nn = 2 * n - 1
            ob = np.zeros((nn, 6), np.float32); aux = np.full(nn, -7, np.int32); par = np.full(nn, -7, np.int32)
            lev = np.full(nn, -7, np.int32); li = np.full(n, -7, np.int32)
            got = r.L.zpcrefcuda_lbvh_build(C.c_int(use), C.c_int(n), _p(bvs), _p(ob), _p(aux), _p(par), _p(lev), _p(li))
            assert got == nn
            out.update({tag + "_orderedBvs": ob, tag + "_auxIndices": aux, tag + "_parents": par, tag + "_levels": lev, tag + "_leafInds": li})
        np.savez(argv[2], **out)
    elif argv[0] == "prims-bench":
        lg, iters = int(argv[1]), int(argv[2])
        n = 1 << lg
        row = dict(log2n=lg, impl="reference CudaExecutionPolicy (cuda_exec(), CUB underneath) vs the same calls on b200_exec()")
        for tag, use in (("ref_cuda", 0), ("overlay", 1)):
            ms = (C.c_double * 3)()
            r.L.zpcrefcuda_prims_bench(C.c_int(use), C.c_size_t(n), C.c_int(iters), ms)
            row.update({tag + "_sort_pair_ms": ms[0], tag + "_scan_ms": ms[1], tag + "_reduce_ms": ms[2],
                        tag + "_sort_pair_gbps": 68 * n / ms[0] / 1e6, tag + "_scan_gbps": 8 * n / ms[1] / 1e6, tag + "_reduce_gbps": 4 * n / ms[2] / 1e6})
        print(json.dumps(row))
    elif argv[0] == "prims":
        n = int(argv[1])
        rs = np.random.RandomState(12345)
        keys = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
        vals = rs.randint(-1000, 1000, size=n).astype(np.int32)
        ko, vo, sc = np.empty_like(keys), np.empty_like(vals), np.empty_like(vals)
        sm, mx = np.zeros(1, np.int32), np.zeros(1, np.int32)
        r.L.zpcrefcuda_overlay_prims(_p(keys), _p(vals), _p(ko), _p(vo), _p(sc), _p(sm), _p(mx), C.c_size_t(n))
        # a TileVector<int, 32> channel through the reference's aosoa iterators
        ntiles, nch, chn, na = 41, 3, 1, 41 * 32 - 5
        tv = rs.randint(-50, 50, size=ntiles * nch * 32).astype(np.int32)
        tvo, tsum = np.zeros_like(tv), np.zeros(1, np.int32)
        r.L.zpcrefcuda_overlay_aosoa(_p(tv), C.c_int(ntiles), C.c_int(nch), C.c_int(chn), C.c_size_t(na), _p(tvo), _p(tsum))
        np.savez(argv[2], keys=keys, vals=vals, keys_out=ko, vals_out=vo, scan=sc, sum=sm, max=mx, tv=tv, tv_scan=tvo, tv_sum=tsum,
                 tv_shape=np.array([ntiles, nch, chn, na]))
    elif argv[0] == "bench":
        from zpc_b200 import synth
        G, s, steps, warmup = (int(a) for a in argv[1:5])
        P = synth.elastic_cube(s, G)
        n = P["x"].shape[0]
        out = r.substep(P, synth.DT, synth.MODEL["E"], synth.MODEL["nu"], synth.GRAVITY, 1, steps=steps + warmup, collect=False,
                        expected_blocks=max(n // 64, 1024))
        t = out["times"][warmup:]
        stage = {k: float(np.mean(v[warmup:])) for k, v in out["stage_ms"].items()}
        print(json.dumps(dict(impl="reference-cuda", n=n, ms_per_step=float(t.mean() * 1e3), ms_min=float(t.min() * 1e3), stage_ms=stage,
                              value=n / float(t.mean()), unit="particle-substeps/s", steps=steps, warmup=warmup,
                              note="the reference's own CUDA functors on cuda_exec(), wall clock around a synchronised substep "
                                   "(partition + clean + P2G + mv+=rhs + update + G2P), particles resident in HBM")))
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main(sys.argv[1:])
