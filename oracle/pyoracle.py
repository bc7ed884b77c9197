"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes bindings for the two checkers:
  * ``Oracle``  -> oracle/libzpcoracle.so  (plain-C restatement, oracle/*.c)
  * ``Ref``     -> oracle/_ref/libzpcref.so (the unmodified reference compiled by oracle/Makefile;
                   present only if it was built where /root/reference is mounted)
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libzpcoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libzpcref.so")

_f = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    srcs = [os.path.join(HERE, f) for f in ("mpm_oracle.c", "prims_oracle.c", "sparse_oracle.c", "bvh_oracle.c", "oracle.h")]
    if (not force and os.path.exists(ORACLE_SO)
            and all(os.path.getmtime(ORACLE_SO) >= os.path.getmtime(s) for s in srcs)):
        return ORACLE_SO
    subprocess.check_call(["make", "-s", "-C", HERE, "-B", "oracle"])
    return ORACLE_SO


def build_ref():
    """(Re)build oracle/_ref from /root/reference when it is mounted; no-op otherwise."""
    if os.path.isdir("/root/reference/include/zensim"):
        subprocess.check_call(["make", "-s", "-C", HERE, "-j8", "ref"])
    return REF_SO if os.path.exists(REF_SO) else None


def build_ref_cuda():
    """(Re)build oracle/_ref/libzpcref_cuda.so — the reference's own CUDA path, compiled for sm_100 — when /root/reference is
    mounted and the library is older than its driver; no-op otherwise (it can only RUN on a GPU box)."""
    so = os.path.join(HERE, "_ref", "libzpcref_cuda.so")
    srcs = [os.path.join(HERE, f) for f in ("ref_driver_cuda.cu", "Makefile", "../include/zpcb200/zs_overlay.cuh", "../include/zpcb200.h")]
    if os.path.isdir("/root/reference/include/zensim") and not (
            os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(f) for f in srcs)):
        subprocess.check_call(["make", "-s", "-C", HERE, "-j8", "refcuda"])
    return so if os.path.exists(so) else None


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


_KT = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64}
_ST = {"i32": np.int32, "u32": np.uint32, "i64": np.int64, "u64": np.uint64, "f32": np.float32, "f64": np.float64}


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.zo_table_size_for.restype = C.c_int
        L.zo_partition_build.restype = C.c_int
        L.zo_table_query.restype = C.c_int
        L.zo_hash_slot0.restype = C.c_int

    # ---- hash / partition ----
    def table_size_for(self, expected):
        return int(self.lib.zo_table_size_for(C.c_int(expected)))

    def hash_slot0(self, key, table_size):
        k = np.ascontiguousarray(key, np.int32)
        return int(self.lib.zo_hash_slot0(_ptr(k), C.c_int(table_size)))

    def partition_build(self, x, dx, table_size):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        keys = np.empty((table_size, 3), np.int32)
        indices = np.empty(table_size, np.int32)
        status = np.empty(table_size, np.int32)
        active = np.zeros((table_size, 3), np.int32)
        cnt = np.zeros(1, np.int32)
        nb = self.lib.zo_partition_build(C.c_int(n), _ptr(x), C.c_float(dx), C.c_int(table_size),
                                         _ptr(keys), _ptr(indices), _ptr(status), _ptr(active),
                                         _ptr(cnt))
        return dict(keys=keys, indices=indices, status=status, active_keys=active[:nb].copy(),
                    nblocks=int(nb), table_size=table_size)

    def index_buckets(self, x, dx, displacement, table_size):
        """-> dict(table arrays, nbuckets, counts[nb+1], offsets[nb+1], indices[n])"""
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        keys = np.empty((table_size, 3), np.int32); idx = np.empty(table_size, np.int32); st = np.empty(table_size, np.int32)
        ak = np.zeros((table_size, 3), np.int32); cnt = np.zeros(1, np.int32)
        counts = np.zeros(n + 2, np.int32); offsets = np.zeros(n + 2, np.int32); indices = np.full(max(n, 1), -1, np.int32)
        self.lib.zo_index_buckets.restype = C.c_int
        nb = self.lib.zo_index_buckets(C.c_int(n), _ptr(x), C.c_float(dx), C.c_float(displacement), C.c_int(table_size), _ptr(keys),
                                       _ptr(idx), _ptr(st), _ptr(ak), _ptr(cnt), _ptr(counts), _ptr(offsets), _ptr(indices))
        return dict(keys=keys, indices=idx, table_size=table_size, nblocks=nb, active_keys=ak[:nb].copy(), counts=counts[:nb + 1].copy(),
                    offsets=offsets[:nb + 1].copy(), ids=indices[:n].copy())

    def table_query(self, key, tab):
        k = np.ascontiguousarray(key, np.int32)
        return int(self.lib.zo_table_query(_ptr(k), C.c_int(tab["table_size"]), _ptr(tab["keys"]),
                                           _ptr(tab["indices"])))

    # ---- bht<i32,3,int,16> + SparseGrid<3,f32,8> accessors (sparse_oracle.c) ----
    def bht_params(self):
        hf = np.zeros(6, np.uint32)
        self.lib.zo_bht_params(_ptr(hf))
        return hf

    def bht_table_size(self, expected):
        self.lib.zo_bht_table_size.restype = C.c_int
        return int(self.lib.zo_bht_table_size(C.c_int(expected)))

    def bht_hash(self, hx, hy, key):
        self.lib.zo_bht_hash.restype = C.c_uint32
        k = np.ascontiguousarray(key, np.int32)
        return int(self.lib.zo_bht_hash(C.c_uint32(int(hx)), C.c_uint32(int(hy)), _ptr(k)))

    def bht_new(self, expected):
        ts = self.bht_table_size(expected)
        t = dict(table_size=ts, hf=self.bht_params(), keys16=np.empty((ts, 4), np.int32), indices=np.empty(ts, np.int32),
                 status=np.empty(ts, np.int32), active_keys=np.zeros((max(ts, 1), 3), np.int32), cnt=np.zeros(1, np.int32))
        self.lib.zo_bht_clear(C.c_int(ts), _ptr(t["keys16"]), _ptr(t["indices"]), _ptr(t["status"]), _ptr(t["cnt"]))
        return t

    def bht_insert(self, t, keys):
        self.lib.zo_bht_insert.restype = C.c_int
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        out = np.empty(keys.shape[0], np.int32)
        for i in range(keys.shape[0]):
            out[i] = self.lib.zo_bht_insert(_ptr(keys[i]), C.c_int(t["table_size"]), _ptr(t["hf"]), _ptr(t["keys16"]),
                                            _ptr(t["indices"]), _ptr(t["active_keys"]), _ptr(t["cnt"]))
        return out

    def bht_query(self, t, keys):
        self.lib.zo_bht_query.restype = C.c_int
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        out = np.empty(keys.shape[0], np.int32)
        for i in range(keys.shape[0]):
            out[i] = self.lib.zo_bht_query(_ptr(keys[i]), C.c_int(t["table_size"]), _ptr(t["hf"]), _ptr(t["keys16"]),
                                           _ptr(t["indices"]))
        return out

    def sg_partition_build(self, x, dx, expected):
        self.lib.zo_sg_partition_build.restype = C.c_int
        t = self.bht_new(expected)
        x = np.ascontiguousarray(x, np.float32)
        nb = self.lib.zo_sg_partition_build(C.c_int(x.shape[0]), _ptr(x), C.c_float(dx), C.c_int(t["table_size"]), _ptr(t["hf"]),
                                            _ptr(t["keys16"]), _ptr(t["indices"]), _ptr(t["status"]), _ptr(t["active_keys"]),
                                            _ptr(t["cnt"]))
        t["nblocks"] = int(nb)
        return t

    def sg_value_or(self, t, grid, chn, coords, dflt):
        self.lib.zo_sg_value_or.restype = C.c_float
        coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
        grid = np.ascontiguousarray(grid, np.float32)
        nch = grid.shape[1]
        out = np.empty(coords.shape[0], np.float32)
        for i in range(coords.shape[0]):
            out[i] = self.lib.zo_sg_value_or(C.c_int(chn), _ptr(coords[i]), C.c_float(dflt), C.c_int(t["table_size"]), _ptr(t["hf"]),
                                             _ptr(t["keys16"]), _ptr(t["indices"]), _ptr(grid), C.c_int(nch))
        return out

    def sg_coords(self, t, m16, bno, cno):
        m16 = np.ascontiguousarray(m16, np.float32)
        ic = np.empty((len(bno), 3), np.int32)
        wc = np.empty((len(bno), 3), np.float32)
        for i in range(len(bno)):
            self.lib.zo_sg_coords(C.c_int(int(bno[i])), C.c_int(int(cno[i])), _ptr(t["active_keys"]), _ptr(m16), _ptr(ic[i]),
                                  _ptr(wc[i]))
        return ic, wc

    def tilevector_reorder_tiles(self, tiles, map_, scatter):
        tiles = np.ascontiguousarray(tiles, np.float32)
        m = np.ascontiguousarray(map_, np.int32)
        out = np.zeros_like(tiles)
        self.lib.zo_tilevector_reorder_tiles(_ptr(tiles), _ptr(out), C.c_int(int(np.prod(tiles.shape[1:]))), _ptr(m),
                                             C.c_int(m.size), C.c_int(int(scatter)))
        return out

    def bht_reorder(self, t, map_, scatter):
        """returns a renumbered copy of table dict t (indices + active_keys)"""
        m = np.ascontiguousarray(map_, np.int32)
        n = m.size
        out = dict(t)
        out["indices"] = t["indices"].copy()
        ok = np.zeros_like(np.ascontiguousarray(t["active_keys"][:n], np.int32))
        ak = np.ascontiguousarray(t["active_keys"][:n], np.int32)
        self.lib.zo_bht_reorder(C.c_int(t["table_size"]), _ptr(np.ascontiguousarray(t["hf"], np.uint32)), _ptr(t["keys16"]),
                                _ptr(out["indices"]), _ptr(ak), C.c_int(n), _ptr(m), C.c_int(int(scatter)), _ptr(ok))
        out["active_keys"] = ok
        return out

    # ---- per particle math ----
    def lame(self, E, nu):
        mu, lam = C.c_float(), C.c_float()
        self.lib.zo_lame(C.c_float(E), C.c_float(nu), C.byref(mu), C.byref(lam))
        return mu.value, lam.value

    def svd3(self, F):
        F = np.ascontiguousarray(F, np.float32)
        U = np.empty(9, np.float32); S = np.empty(3, np.float32); V = np.empty(9, np.float32)
        self.lib.zo_svd3(_ptr(F), _ptr(U), _ptr(S), _ptr(V))
        return U, S, V

    def stress_fixedcorotated(self, volume, E, nu, F):
        mu, lam = self.lame(E, nu)
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        self.lib.zo_stress_fixedcorotated(C.c_float(volume), C.c_float(mu), C.c_float(lam), _ptr(F),
                                          _ptr(PF))
        return PF

    # ---- transfer ----
    def p2g(self, P, tab, dx, dt, E, nu, volume, grid=None):
        nb = tab["nblocks"]
        if grid is None:
            grid = np.zeros((nb, 7, 64), np.float32)
        n = P["x"].shape[0]
        self.lib.zo_p2g_fcr(C.c_int(n), _ptr(P["x"]), _ptr(P["v"]), _ptr(P["m"]), _ptr(P["C"]),
                            _ptr(P["F"]), C.c_float(dx), C.c_float(dt), C.c_float(E), C.c_float(nu),
                            C.c_float(volume), C.c_int(tab["table_size"]), _ptr(tab["keys"]),
                            _ptr(tab["indices"]), _ptr(grid))
        return grid

    def stress_vonmises(self, volume, E, nu, yield_stress, F):
        mu, lam = self.lame(E, nu)
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        self.lib.zo_stress_vonmises(C.c_float(volume), C.c_float(mu), C.c_float(lam), C.c_float(yield_stress), _ptr(F), _ptr(PF))
        return PF

    def p2g_vonmises(self, P, tab, dx, dt, E, nu, yield_stress, volume, grid=None):
        nb = tab["nblocks"]
        if grid is None:
            grid = np.zeros((nb, 7, 64), np.float32)
        n = P["x"].shape[0]
        self.lib.zo_p2g_vonmises(C.c_int(n), _ptr(P["x"]), _ptr(P["v"]), _ptr(P["m"]), _ptr(P["C"]), _ptr(P["F"]),
                                 C.c_float(dx), C.c_float(dt), C.c_float(E), C.c_float(nu), C.c_float(yield_stress),
                                 C.c_float(volume), C.c_int(tab["table_size"]), _ptr(tab["keys"]), _ptr(tab["indices"]),
                                 _ptr(grid))
        return grid

    # DruckerPragerConfig / NACCConfig.  sand = dict(cohesion, beta, yieldSurface, volumeCorrection),
    # nacc = dict(fa, xi, beta, hardeningOn); both return (PF, logJp_after)
    def math_sqrt(self, x):
        self.lib.zo_math_sqrt.restype = C.c_float
        return float(self.lib.zo_math_sqrt(C.c_float(x)))

    def nacc_consts(self, E, nu, fa, dim=3):
        self.lib.zo_nacc_bulk.restype = C.c_float
        self.lib.zo_nacc_msqr.restype = C.c_float
        return (float(self.lib.zo_nacc_bulk(C.c_float(E), C.c_float(nu))),
                float(self.lib.zo_nacc_msqr(C.c_float(fa), C.c_int(dim))))

    def stress_sand(self, volume, E, nu, sand, logJp, F):
        mu, lam = self.lame(E, nu)
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        lj = C.c_float(logJp)
        self.lib.zo_stress_sand(C.c_float(volume), C.c_float(mu), C.c_float(lam), C.c_float(sand["cohesion"]),
                                C.c_float(sand["beta"]), C.c_float(sand["yieldSurface"]),
                                C.c_int(int(sand["volumeCorrection"])), C.byref(lj), _ptr(F), _ptr(PF))
        return PF, lj.value

    def stress_nacc(self, volume, E, nu, nacc, logJp, F):
        mu, lam = self.lame(E, nu)
        bm, msqr = self.nacc_consts(E, nu, nacc["fa"])
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        lj = C.c_float(logJp)
        self.lib.zo_stress_nacc(C.c_float(volume), C.c_float(mu), C.c_float(lam), C.c_float(bm), C.c_float(nacc["xi"]),
                                C.c_float(nacc["beta"]), C.c_float(msqr), C.c_int(int(nacc["hardeningOn"])),
                                C.byref(lj), _ptr(F), _ptr(PF))
        return PF, lj.value

    def p2g_sand(self, P, tab, dx, dt, E, nu, sand, volume, grid=None):
        """P["logJp"] is updated in place (P2G.hpp:101)."""
        nb = tab["nblocks"]
        if grid is None:
            grid = np.zeros((nb, 7, 64), np.float32)
        n = P["x"].shape[0]
        self.lib.zo_p2g_sand(C.c_int(n), _ptr(P["x"]), _ptr(P["v"]), _ptr(P["m"]), _ptr(P["C"]), _ptr(P["F"]),
                             _ptr(P["logJp"]), C.c_float(dx), C.c_float(dt), C.c_float(E), C.c_float(nu),
                             C.c_float(sand["cohesion"]), C.c_float(sand["beta"]), C.c_float(sand["yieldSurface"]),
                             C.c_int(int(sand["volumeCorrection"])), C.c_float(volume), C.c_int(tab["table_size"]),
                             _ptr(tab["keys"]), _ptr(tab["indices"]), _ptr(grid))
        return grid

    def p2g_nacc(self, P, tab, dx, dt, E, nu, nacc, volume, grid=None):
        """P["logJp"] is updated in place (P2G.hpp:101)."""
        nb = tab["nblocks"]
        if grid is None:
            grid = np.zeros((nb, 7, 64), np.float32)
        n = P["x"].shape[0]
        self.lib.zo_p2g_nacc(C.c_int(n), _ptr(P["x"]), _ptr(P["v"]), _ptr(P["m"]), _ptr(P["C"]), _ptr(P["F"]),
                             _ptr(P["logJp"]), C.c_float(dx), C.c_float(dt), C.c_float(E), C.c_float(nu),
                             C.c_float(nacc["fa"]), C.c_float(nacc["xi"]), C.c_float(nacc["beta"]),
                             C.c_int(int(nacc["hardeningOn"])), C.c_int(3), C.c_float(volume), C.c_int(tab["table_size"]),
                             _ptr(tab["keys"]), _ptr(tab["indices"]), _ptr(grid))
        return grid

    def p2g_eos(self, P, tab, dx, dt, bulk, viscosity, volume, grid=None):
        nb = tab["nblocks"]
        if grid is None:
            grid = np.zeros((nb, 7, 64), np.float32)
        n = P["x"].shape[0]
        self.lib.zo_p2g_eos(C.c_int(n), _ptr(P["x"]), _ptr(P["v"]), _ptr(P["m"]), _ptr(P["C"]), _ptr(P["J"]),
                            C.c_float(dx), C.c_float(dt), C.c_float(bulk), C.c_float(viscosity), C.c_float(volume),
                            C.c_int(tab["table_size"]), _ptr(tab["keys"]), _ptr(tab["indices"]), _ptr(grid))
        return grid

    def g2p_eos(self, P, tab, grid, dx, dt):
        n = P["x"].shape[0]
        self.lib.zo_g2p_eos(C.c_int(n), _ptr(P["x"]), _ptr(P["v"]), _ptr(P["C"]), _ptr(P["J"]), C.c_float(dx),
                            C.c_float(dt), C.c_int(tab["table_size"]), _ptr(tab["keys"]), _ptr(tab["indices"]),
                            _ptr(grid))

    def grid_update(self, grid, dt, extf, mode):
        e = np.ascontiguousarray(extf, np.float32)
        mx = np.zeros(1, np.float32)
        self.lib.zo_grid_update(C.c_int(grid.shape[0]), _ptr(grid), C.c_float(dt), _ptr(e),
                                C.c_int(mode), _ptr(mx))
        return float(mx[0])

    def grid_momentum_to_velocity(self, grid, m_chn=0, mv_chn=1):
        mx = np.zeros(1, np.float32)
        self.lib.zo_grid_momentum_to_velocity(C.c_int(grid.shape[0]), C.c_int(grid.shape[1]), _ptr(grid), C.c_int(m_chn),
                                              C.c_int(mv_chn), _ptr(mx))
        return float(mx[0])

    def grid_angular_momentum(self, grid, active_keys, dx, m_chn=0, mv_chn=1):
        k = np.ascontiguousarray(active_keys, np.int32)
        out = np.zeros(6, np.float64)
        self.lib.zo_grid_angular_momentum(C.c_int(grid.shape[0]), C.c_int(grid.shape[1]), _ptr(k), _ptr(grid), C.c_float(dx),
                                          C.c_int(m_chn), C.c_int(mv_chn), _ptr(out))
        return out

    def apply_boundary(self, grid, active_keys, dx, geom, ctype, p0, p1, motion=None):
        """motion: None (static) or 20 floats b[3], dbdt[3], R[9] row-major, omega[3], s, dsdt"""
        k = np.ascontiguousarray(active_keys, np.int32)
        a = np.ascontiguousarray(p0, np.float32); b = np.ascontiguousarray(p1, np.float32)
        if motion is None:
            self.lib.zo_apply_boundary(C.c_int(grid.shape[0]), _ptr(k), _ptr(grid), C.c_float(dx), C.c_int(geom),
                                       C.c_int(ctype), _ptr(a), _ptr(b))
        else:
            m = np.ascontiguousarray(motion, np.float32)
            assert m.size == 20
            self.lib.zo_apply_boundary_moving(C.c_int(grid.shape[0]), _ptr(k), _ptr(grid), C.c_float(dx), C.c_int(geom),
                                              C.c_int(ctype), _ptr(a), _ptr(b), _ptr(m))

    # ---- LBvh<3,int,f32> ----
    def lbvh_build(self, bvs, refit=True):
        bvs = np.ascontiguousarray(bvs, np.float32).reshape(-1, 6)
        n = bvs.shape[0]
        nn = 2 * n - 1 if n > 2 else n
        out = dict(n=n, orderedBvs=np.zeros((nn, 6), np.float32), auxIndices=np.full(nn, -7, np.int32),
                   parents=np.full(nn, -7, np.int32), levels=np.full(nn, -7, np.int32), leafInds=np.full(n, -7, np.int32))
        self.lib.zo_lbvh_build(C.c_int(n), _ptr(bvs), _ptr(out["orderedBvs"]), _ptr(out["auxIndices"]), _ptr(out["parents"]),
                               _ptr(out["levels"]), _ptr(out["leafInds"]), C.c_int(int(refit)))
        return out

    def lbvh_refit(self, bvh, bvs):
        bvs = np.ascontiguousarray(bvs, np.float32).reshape(-1, 6)
        self.lib.zo_lbvh_refit(C.c_int(bvh["n"]), _ptr(bvs), _ptr(bvh["orderedBvs"]), _ptr(bvh["auxIndices"]), _ptr(bvh["parents"]),
                               _ptr(bvh["levels"]), _ptr(bvh["leafInds"]))

    def lbvh_whole_box_and_codes(self, bvs):
        bvs = np.ascontiguousarray(bvs, np.float32).reshape(-1, 6)
        box = np.empty(6, np.float32)
        self.lib.zo_lbvh_whole_box(C.c_int(bvs.shape[0]), _ptr(bvs), _ptr(box))
        self.lib.zo_lbvh_morton.restype = C.c_uint32
        return box, np.array([self.lib.zo_lbvh_morton(_ptr(box), _ptr(b)) for b in bvs], np.uint32)

    def lbvh_iter_neighbors(self, bvh, bv, cap=4096):
        bv = np.ascontiguousarray(bv, np.float32)
        out = np.empty(cap, np.int32)
        self.lib.zo_lbvh_iter_neighbors.restype = C.c_int
        c = self.lib.zo_lbvh_iter_neighbors(C.c_int(bvh["n"]), _ptr(bvh["orderedBvs"]), _ptr(bvh["auxIndices"]), _ptr(bvh["levels"]),
                                            _ptr(bv), _ptr(out), C.c_int(cap))
        assert c <= cap
        return out[:c].copy()

    def g2p2g(self, model, prm, P, tab, dx, dt, E, nu, volume, gridv):
        """G2P2GTransfer (pinned against the reference's own functor on the GPU, tests/test_gpu_models.py): gridv [nb*64, 3] -> gridr [nb*64, 3]"""
        n = P["x"].shape[0]
        prm = np.ascontiguousarray(prm, np.float32)
        gridv = np.ascontiguousarray(gridv, np.float32)
        gridr = np.zeros_like(gridv)
        nul = C.c_void_p(None)
        self.lib.zo_g2p2g(C.c_int(model), _ptr(prm), C.c_int(n), _ptr(P["x"]), _ptr(P["F"]) if "F" in P else nul,
                          _ptr(P["J"]) if "J" in P else nul, _ptr(P["logJp"]) if "logJp" in P else nul, C.c_float(dx), C.c_float(dt),
                          C.c_float(E), C.c_float(nu), C.c_float(volume), C.c_int(tab["table_size"]), _ptr(tab["keys"]),
                          _ptr(tab["indices"]), _ptr(gridv), _ptr(gridr))
        return gridr

    def cuboid(self, x, mn, mx):
        x = np.ascontiguousarray(x, np.float32)
        mn = np.ascontiguousarray(mn, np.float32); mx = np.ascontiguousarray(mx, np.float32)
        sdf = np.empty(x.shape[0], np.float32); nm = np.empty((x.shape[0], 3), np.float32)
        self.lib.zo_cuboid(C.c_int(x.shape[0]), _ptr(x), _ptr(mn), _ptr(mx), _ptr(sdf), _ptr(nm))
        return sdf, nm

    def g2p(self, P, tab, grid, dx, dt):
        n = P["x"].shape[0]
        self.lib.zo_g2p(C.c_int(n), _ptr(P["x"]), _ptr(P["v"]), _ptr(P["C"]), _ptr(P["F"]),
                        C.c_float(dx), C.c_float(dt), C.c_int(tab["table_size"]), _ptr(tab["keys"]),
                        _ptr(tab["indices"]), _ptr(grid))

    def substep(self, P, dx, dt, E, nu, volume, gravity, mode, expected_blocks=None):
        """Composed explicit APIC substep of SURVEY §3.1; P is updated in place."""
        n = P["x"].shape[0]
        ts = self.table_size_for(expected_blocks or max(n // 8, 1))
        tab = self.partition_build(P["x"], dx, ts)
        grid = self.p2g(P, tab, dx, dt, E, nu, volume)
        grid_p2g = grid.copy()
        mx = self.grid_update(grid, dt, (0.0, gravity, 0.0), mode)
        self.g2p(P, tab, grid, dx, dt)
        return dict(tab=tab, grid_p2g=grid_p2g, grid=grid, max_vel_sqr=mx)

    # ---- primitives ----
    def radix_sort_pair(self, kind, keys, vals, sbit=0, ebit=None):
        kt = _KT[kind]
        keys = np.ascontiguousarray(keys, kt); vals = np.ascontiguousarray(vals, np.int32)
        ko = np.empty_like(keys); vo = np.empty_like(vals)
        ebit = keys.itemsize * 8 if ebit is None else ebit
        getattr(self.lib, "zo_radix_sort_pair_" + kind)(_ptr(keys), _ptr(vals), _ptr(ko), _ptr(vo),
                                                        C.c_size_t(keys.size), C.c_int(sbit),
                                                        C.c_int(ebit))
        return ko, vo

    def radix_sort(self, kind, keys, sbit=0, ebit=None):
        kt = _KT[kind]
        keys = np.ascontiguousarray(keys, kt)
        ko = np.empty_like(keys)
        ebit = keys.itemsize * 8 if ebit is None else ebit
        getattr(self.lib, "zo_radix_sort_" + kind)(_ptr(keys), _ptr(ko), C.c_size_t(keys.size),
                                                   C.c_int(sbit), C.c_int(ebit))
        return ko

    def scan(self, which, kind, a):
        a = np.ascontiguousarray(a, _ST[kind])
        out = np.empty_like(a)
        getattr(self.lib, "zo_%s_scan_sum_%s" % (which, kind))(_ptr(a), _ptr(out),
                                                               C.c_size_t(a.size))
        return out

    def reduce(self, op, kind, a):
        a = np.ascontiguousarray(a, _ST[kind])
        out = np.zeros(1, _ST[kind])
        getattr(self.lib, "zo_reduce_%s_%s" % (op, kind))(_ptr(a), _ptr(out), C.c_size_t(a.size))
        return out[0]

    def merge_sort_pair(self, kind, keys, vals):
        k = np.array(keys, _ST[kind]); v = np.array(vals, np.int32)
        getattr(self.lib, "zo_merge_sort_pair_" + kind)(_ptr(k), _ptr(v), C.c_size_t(k.size))
        return k, v


class Ref:
    """The unmodified reference (seq_exec when nthreads == 0, omp_exec().threads(n) otherwise)."""

    def __init__(self, path=None):
        path = path or REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.zpcref_mpm_create.restype = C.c_void_p
        L.zpcref_mpm_partition.restype = C.c_int
        L.zpcref_mpm_table_size.restype = C.c_int
        L.zpcref_mpm_get_maxvel.restype = C.c_float
        L.zpcref_max_threads.restype = C.c_int

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def max_threads(self):
        return int(self.lib.zpcref_max_threads())

    class Bht:
        """the reference's bht<int,3,int,16> on the host (container/Bht.hpp)"""

        def __init__(self, ref, expected, handle=None):
            self.L = ref.lib
            self.L.zpcref_bht_create.restype = C.c_void_p
            self.own = handle is None
            self.h = C.c_void_p(self.L.zpcref_bht_create(C.c_int(expected))) if handle is None else handle

        def close(self):
            if self.h and self.own:
                self.L.zpcref_bht_destroy(self.h)
            self.h = None

        def info(self):
            info = np.zeros(3, np.int32)
            hf = np.zeros(6, np.uint32)
            self.L.zpcref_bht_info(self.h, _ptr(info), _ptr(hf))
            return dict(table_size=int(info[0]), num_buckets=int(info[1]), cnt=int(info[2]), hf=hf)

        def insert(self, keys):
            keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
            out = np.empty(keys.shape[0], np.int32)
            self.L.zpcref_bht_insert(self.h, _ptr(keys), C.c_int(keys.shape[0]), _ptr(out))
            return out

        def query(self, keys):
            keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
            out = np.empty(keys.shape[0], np.int32)
            self.L.zpcref_bht_query(self.h, _ptr(keys), C.c_int(keys.shape[0]), _ptr(out))
            return out

        def arrays(self):
            i = self.info()
            ts = i["table_size"]
            k16 = np.empty((ts, 4), np.int32); idx = np.empty(ts, np.int32); st = np.empty(ts, np.int32)
            ak = np.empty((max(i["cnt"], 1), 3), np.int32)
            self.L.zpcref_bht_get(self.h, _ptr(k16), _ptr(idx), _ptr(st), _ptr(ak))
            return dict(keys16=k16, indices=idx, status=st, active_keys=ak[: i["cnt"]], cnt=i["cnt"], table_size=ts, hf=i["hf"])

        def reorder(self, map_, scatter):
            m = np.ascontiguousarray(map_, np.int32)
            self.L.zpcref_bht_reorder(self.h, _ptr(m), C.c_int(int(scatter)))

        def load(self, keys16, indices, active_keys, cnt):
            """overwrite the container's arrays (e.g. with a table built on the GPU); query() then runs the
            reference's own BHTView::query over them"""
            k = np.ascontiguousarray(keys16, np.int32); ix = np.ascontiguousarray(indices, np.int32)
            ak = np.ascontiguousarray(active_keys, np.int32)
            assert k.shape[0] == self.info()["table_size"]
            self.L.zpcref_bht_load(self.h, _ptr(k), _ptr(ix), _ptr(ak), C.c_int(int(cnt)))

    class SparseGrid:
        """the reference's SparseGrid<3,f32,8> on the host (geometry/SparseGrid.hpp)"""

        def __init__(self, ref, nblocks, nch):
            self.ref, self.L = ref, ref.lib
            self.L.zpcref_sg_create.restype = C.c_void_p
            self.L.zpcref_sg_table.restype = C.c_void_p
            self.h = C.c_void_p(self.L.zpcref_sg_create(C.c_int(nblocks), C.c_int(nch)))
            self.table = Ref.Bht(ref, 0, handle=C.c_void_p(self.L.zpcref_sg_table(self.h)))
            self.nch = nch

        def close(self):
            if self.h:
                self.L.zpcref_sg_destroy(self.h)
                self.h = None

        def scale(self, s):
            self.L.zpcref_sg_scale(self.h, C.c_float(s))

        def translate(self, t):
            t = np.ascontiguousarray(t, np.float32)
            self.L.zpcref_sg_translate(self.h, _ptr(t))

        def transform(self):
            m = np.empty(16, np.float32)
            self.L.zpcref_sg_get_transform(self.h, _ptr(m))
            return m

        def set_background(self, b):
            self.L.zpcref_sg_set_background(self.h, C.c_float(b))

        def load_grid(self, data):
            data = np.ascontiguousarray(data, np.float32)
            self.L.zpcref_sg_load_grid(self.h, _ptr(data), C.c_int(data.shape[0]))

        def reorder_tiles(self, map_, scatter):
            m = np.ascontiguousarray(map_, np.int32)
            out = np.empty((m.size, self.nch, 512), np.float32)
            self.L.zpcref_sg_reorder_tiles(self.h, _ptr(m), C.c_int(m.size), C.c_int(int(scatter)), _ptr(out))
            return out

        def value_or(self, chn, coords, dflt):
            coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
            out = np.empty(coords.shape[0], np.float32)
            self.L.zpcref_sg_value_or(self.h, C.c_int(chn), _ptr(coords), C.c_int(coords.shape[0]), C.c_float(dflt), _ptr(out))
            return out

        def coords(self, bno, cno):
            bno = np.ascontiguousarray(bno, np.int32); cno = np.ascontiguousarray(cno, np.int32)
            ic = np.empty((bno.shape[0], 3), np.int32); wc = np.empty((bno.shape[0], 3), np.float32)
            self.L.zpcref_sg_coords(self.h, _ptr(bno), _ptr(cno), C.c_int(bno.shape[0]), _ptr(ic), _ptr(wc))
            return ic, wc

    class Mpm:
        def __init__(self, ref, n, dx, nthreads, expected_blocks):
            self.L = ref.lib
            self.n = n
            self.h = C.c_void_p(self.L.zpcref_mpm_create(C.c_int(n), C.c_float(dx),
                                                         C.c_int(nthreads),
                                                         C.c_int(expected_blocks)))
            self.nblocks = 0

        def close(self):
            if self.h:
                self.L.zpcref_mpm_destroy(self.h)
                self.h = None

        def set_particles(self, P):
            self.L.zpcref_mpm_set_particles(self.h, _ptr(P["x"]), _ptr(P["v"]), _ptr(P["m"]),
                                            _ptr(P["C"]), _ptr(P["F"]))

        def get_particles(self):
            n = self.n
            x = np.empty((n, 3), np.float32); v = np.empty((n, 3), np.float32)
            Cm = np.empty((n, 9), np.float32); F = np.empty((n, 9), np.float32)
            self.L.zpcref_mpm_get_particles(self.h, _ptr(x), _ptr(v), _ptr(Cm), _ptr(F))
            return dict(x=x, v=v, C=Cm, F=F)

        def partition(self):
            self.nblocks = int(self.L.zpcref_mpm_partition(self.h))
            return self.nblocks

        def keys(self):
            k = np.empty((self.nblocks, 3), np.int32)
            self.L.zpcref_mpm_get_keys(self.h, _ptr(k))
            return k

        def table(self):
            ts = int(self.L.zpcref_mpm_table_size(self.h))
            keys = np.empty((ts, 3), np.int32); idx = np.empty(ts, np.int32)
            self.L.zpcref_mpm_get_table(self.h, _ptr(keys), _ptr(idx))
            return dict(keys=keys, indices=idx, table_size=ts, nblocks=self.nblocks,
                        active_keys=self.keys())

        def clean_grid(self):
            self.L.zpcref_mpm_clean_grid(self.h)

        def p2g(self, dt, E, nu, volume):
            self.L.zpcref_mpm_p2g(self.h, C.c_float(dt), C.c_float(E), C.c_float(nu),
                                  C.c_float(volume))

        def p2g_vonmises(self, dt, E, nu, yield_stress, volume):
            self.L.zpcref_mpm_p2g_vonmises(self.h, C.c_float(dt), C.c_float(E), C.c_float(nu), C.c_float(yield_stress),
                                           C.c_float(volume))

        def grid_update(self, dt, gravity, mode):
            self.L.zpcref_mpm_grid_update(self.h, C.c_float(dt), C.c_float(gravity), C.c_int(mode))
            return float(self.L.zpcref_mpm_get_maxvel(self.h))

        def g2p(self, dt):
            self.L.zpcref_mpm_g2p(self.h, C.c_float(dt))

        def momentum_to_velocity(self):
            self.L.zpcref_mpm_momentum_to_velocity(self.h)
            return float(self.L.zpcref_mpm_get_maxvel(self.h))

        def angular_momentum(self):
            out = np.zeros(6, np.float64)
            self.L.zpcref_mpm_angular_momentum(self.h, _ptr(out))
            return out

        def apply_boundary(self, geom, ctype, p0, p1, motion=None):
            a = np.ascontiguousarray(p0, np.float32); b = np.ascontiguousarray(p1, np.float32)
            if motion is None:
                self.L.zpcref_mpm_apply_boundary(self.h, C.c_int(geom), C.c_int(ctype), _ptr(a), _ptr(b))
            else:
                m = np.ascontiguousarray(motion, np.float32)
                self.L.zpcref_mpm_apply_boundary_moving(self.h, C.c_int(geom), C.c_int(ctype), _ptr(a), _ptr(b), _ptr(m))

        def set_J(self, J):
            self.L.zpcref_mpm_set_J(self.h, _ptr(np.ascontiguousarray(J, np.float32)))

        def get_J(self):
            J = np.empty(self.n, np.float32)
            self.L.zpcref_mpm_get_J(self.h, _ptr(J))
            return J

        def set_logJp(self, logJp):
            self.L.zpcref_mpm_set_logJp(self.h, _ptr(np.ascontiguousarray(logJp, np.float32)))

        def get_logJp(self):
            a = np.empty(self.n, np.float32)
            self.L.zpcref_mpm_get_logJp(self.h, _ptr(a))
            return a

        def p2g_sand(self, dt, E, nu, sand, volume):
            self.L.zpcref_mpm_p2g_sand(self.h, C.c_float(dt), C.c_float(E), C.c_float(nu), C.c_float(sand["cohesion"]),
                                       C.c_float(sand["beta"]), C.c_float(sand["yieldSurface"]),
                                       C.c_int(int(sand["volumeCorrection"])), C.c_float(volume))

        def p2g_nacc(self, dt, E, nu, nacc, volume):
            self.L.zpcref_mpm_p2g_nacc(self.h, C.c_float(dt), C.c_float(E), C.c_float(nu), C.c_float(nacc["fa"]),
                                       C.c_float(nacc["xi"]), C.c_float(nacc["beta"]), C.c_int(int(nacc["hardeningOn"])),
                                       C.c_float(volume))

        def p2g_eos(self, dt, bulk, gamma, viscosity, volume):
            self.L.zpcref_mpm_p2g_eos(self.h, C.c_float(dt), C.c_float(bulk), C.c_float(gamma), C.c_float(viscosity),
                                      C.c_float(volume))

        def g2p_eos(self, dt):
            self.L.zpcref_mpm_g2p_eos(self.h, C.c_float(dt))

        def grid(self):
            g = np.empty((self.nblocks, 7, 64), np.float32)
            self.L.zpcref_mpm_get_grid(self.h, _ptr(g))
            return g

    def mpm(self, n, dx, nthreads=0, expected_blocks=None):
        return Ref.Mpm(self, n, dx, nthreads, expected_blocks or max(n // 8, 1))

    def radix_sort_pair(self, kind, keys, vals, sbit=0, ebit=None, nthreads=0):
        kt = _KT[kind]
        keys = np.ascontiguousarray(keys, kt); vals = np.ascontiguousarray(vals, np.int32)
        ko = np.empty_like(keys); vo = np.empty_like(vals)
        ebit = keys.itemsize * 8 if ebit is None else ebit
        getattr(self.lib, "zpcref_radix_sort_pair_" + kind)(C.c_int(nthreads), _ptr(keys),
                                                            _ptr(vals), _ptr(ko), _ptr(vo),
                                                            C.c_size_t(keys.size), C.c_int(sbit),
                                                            C.c_int(ebit))
        return ko, vo

    def radix_sort(self, kind, keys, sbit=0, ebit=None, nthreads=0):
        kt = _KT[kind]
        keys = np.ascontiguousarray(keys, kt)
        ko = np.empty_like(keys)
        ebit = keys.itemsize * 8 if ebit is None else ebit
        getattr(self.lib, "zpcref_radix_sort_" + kind)(C.c_int(nthreads), _ptr(keys), _ptr(ko),
                                                       C.c_size_t(keys.size), C.c_int(sbit),
                                                       C.c_int(ebit))
        return ko

    def scan(self, which, kind, a, nthreads=0):
        a = np.ascontiguousarray(a, _ST[kind])
        out = np.empty_like(a)
        getattr(self.lib, "zpcref_%s_scan_sum_%s" % (which, kind))(C.c_int(nthreads), _ptr(a),
                                                                   _ptr(out), C.c_size_t(a.size))
        return out

    def reduce(self, op, kind, a, nthreads=0):
        a = np.ascontiguousarray(a, _ST[kind])
        out = np.zeros(1, _ST[kind])
        getattr(self.lib, "zpcref_reduce_%s_%s" % (op, kind))(C.c_int(nthreads), _ptr(a), _ptr(out),
                                                              C.c_size_t(a.size))
        return out[0]

    def stress_vonmises(self, volume, E, nu, yield_stress, F):
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        self.lib.zpcref_stress_vonmises(C.c_float(volume), C.c_float(E), C.c_float(nu), C.c_float(yield_stress), _ptr(F), _ptr(PF))
        return PF

    def stress_sand(self, volume, E, nu, sand, logJp, F):
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        lj = C.c_float(logJp)
        self.lib.zpcref_stress_sand(C.c_float(volume), C.c_float(E), C.c_float(nu), C.c_float(sand["cohesion"]),
                                    C.c_float(sand["beta"]), C.c_float(sand["yieldSurface"]),
                                    C.c_int(int(sand["volumeCorrection"])), C.byref(lj), _ptr(F), _ptr(PF))
        return PF, lj.value

    def stress_nacc(self, volume, E, nu, nacc, logJp, F):
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        lj = C.c_float(logJp)
        self.lib.zpcref_stress_nacc(C.c_float(volume), C.c_float(E), C.c_float(nu), C.c_float(nacc["fa"]),
                                    C.c_float(nacc["xi"]), C.c_float(nacc["beta"]), C.c_int(int(nacc["hardeningOn"])),
                                    C.byref(lj), _ptr(F), _ptr(PF))
        return PF, lj.value

    def index_buckets(self, x, dx, displacement, nthreads=0):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        ak = np.zeros((n, 3), np.int32); counts = np.zeros(n + 2, np.int32); offsets = np.zeros(n + 2, np.int32)
        indices = np.full(max(n, 1), -1, np.int32)
        self.lib.zpcref_index_buckets.restype = C.c_int
        nb = self.lib.zpcref_index_buckets(C.c_int(nthreads), C.c_int(n), _ptr(x), C.c_float(dx), C.c_float(displacement), _ptr(ak),
                                           _ptr(counts), _ptr(offsets), _ptr(indices))
        return dict(nblocks=nb, active_keys=ak[:nb].copy(), counts=counts[:nb + 1].copy(), offsets=offsets[:nb + 1].copy(), ids=indices[:n].copy())

    def lbvh_build(self, bvs, refit=True, nthreads=0):
        bvs = np.ascontiguousarray(bvs, np.float32).reshape(-1, 6)
        n = bvs.shape[0]
        nn = 2 * n - 1 if n > 2 else n
        out = dict(n=n, orderedBvs=np.zeros((nn, 6), np.float32), auxIndices=np.full(nn, -7, np.int32),
                   parents=np.full(nn, -7, np.int32), levels=np.full(nn, -7, np.int32), leafInds=np.full(n, -7, np.int32))
        self.lib.zpcref_lbvh_build.restype = C.c_int
        got = self.lib.zpcref_lbvh_build(C.c_int(nthreads), C.c_int(n), _ptr(bvs), _ptr(out["orderedBvs"]), _ptr(out["auxIndices"]),
                                         _ptr(out["parents"]), _ptr(out["levels"]), _ptr(out["leafInds"]), C.c_int(int(refit)))
        assert got == nn
        return out

    def lbvh_build_then_refit(self, bvs0, bvs1, nthreads=0):
        a = np.ascontiguousarray(bvs0, np.float32).reshape(-1, 6)
        b = np.ascontiguousarray(bvs1, np.float32).reshape(-1, 6)
        n = a.shape[0]
        out = np.zeros((2 * n - 1 if n > 2 else n, 6), np.float32)
        self.lib.zpcref_lbvh_build_then_refit(C.c_int(nthreads), C.c_int(n), _ptr(a), _ptr(b), _ptr(out))
        return out

    def math_sqrt(self, x):
        self.lib.zpcref_math_sqrt.restype = C.c_float
        return float(self.lib.zpcref_math_sqrt(C.c_float(x)))

    def nacc_consts(self, E, nu, fa):
        b, m = C.c_float(), C.c_float()
        self.lib.zpcref_nacc_consts(C.c_float(E), C.c_float(nu), C.c_float(fa), C.byref(b), C.byref(m))
        return b.value, m.value

    def merge_sort_pair(self, kind, keys, vals, nthreads=0):
        k = np.array(keys, _ST[kind]); v = np.array(vals, np.int32)
        getattr(self.lib, "zpcref_merge_sort_pair_" + kind)(C.c_int(nthreads), _ptr(k), _ptr(v), C.c_size_t(k.size))
        return k, v

    def svd3(self, F):
        F = np.ascontiguousarray(F, np.float32)
        U = np.empty(9, np.float32); S = np.empty(3, np.float32); V = np.empty(9, np.float32)
        self.lib.zpcref_svd3(_ptr(F), _ptr(U), _ptr(S), _ptr(V))
        return U, S, V

    def stress_fixedcorotated(self, volume, E, nu, F):
        F = np.ascontiguousarray(F, np.float32)
        PF = np.empty(9, np.float32)
        self.lib.zpcref_stress_fixedcorotated(C.c_float(volume), C.c_float(E), C.c_float(nu),
                                              _ptr(F), _ptr(PF))
        return PF
