"""Host-side mirror of the reference interface for the MPM transfer path, over the C ABI in
include/zpcb200.h (ctypes).  torch is used for device memory and streams only.

Mirrors (reference paths relative to /root/reference/include/zensim/):
  CudaExecutionPolicy  <- cuda/execution/ExecutionPolicy.cuh:362-912 (reduce / *_scan / radix_sort(_pair),
                          chained setters device().stream().sync(), sync defaults to True)
  HashTable            <- container/HashTable.hpp:15-206        (keys/indices/status/_activeKeys/_cnt)
  Grids                <- geometry/Structure.hpp:140-260        (collocated TileVector<f32,64>, {m,v,rhs})
  Particles            <- geometry/Structurefree.hpp:22-224     (AoS per attribute)
  TileVector           <- container/TileVector.hpp:14-561       (AoSoA, length-32 tiles)
  partition_for_particles, CleanGridBlocks, P2GTransfer, ComputeGridBlockVelocity, G2PTransfer
                       <- simulation/{sparsity,grid,transfer}/*.hpp as free functions taking a policy.
There is no CPU fallback: a missing libzpcb200.so or a missing GPU raises.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZPCB200_LIB") or os.path.join(HERE, "libzpcb200.so")

BIN_MAX = 1024
PB_M, PB_X, PB_V, PB_C, PB_F, PB_NCH = 0, 1, 4, 7, 16, 25


class ZpcError(RuntimeError):
    pass


class zpc_port(C.Structure):
    _fields_ = [("base", C.c_void_p), ("idx", C.c_uint32), ("numTileBits", C.c_uint32),
                ("tileMask", C.c_uint32), ("numChns", C.c_uint32)]


class zpc_particles_view(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("M", "X", "V", "Dinv", "J", "F", "C", "logJp")] + [("count", C.c_size_t)]


class zpc_hashtable_view(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("indices", C.c_void_p), ("status", C.c_void_p),
                ("activeKeys", C.c_void_p), ("tableSize", C.c_int), ("cnt", C.c_void_p)]


class zpc_grids_view(C.Structure):
    _fields_ = [("tiles", C.c_void_p), ("numBlocks", C.c_size_t), ("numChannels", C.c_int), ("dx", C.c_float)]


class zpc_tilevector_view(C.Structure):
    _fields_ = [("base", C.c_void_p), ("size", C.c_size_t), ("numChannels", C.c_int)]


class zpc_fixed_corotated(C.Structure):
    _fields_ = [("rho", C.c_float), ("volume", C.c_float), ("dim", C.c_int), ("E", C.c_float), ("nu", C.c_float)]


class zpc_vonmises_fixed_corotated(C.Structure):
    _fields_ = [("rho", C.c_float), ("volume", C.c_float), ("dim", C.c_int), ("E", C.c_float), ("nu", C.c_float),
                ("yieldStress", C.c_float)]


class zpc_drucker_prager(C.Structure):
    _fields_ = [("rho", C.c_float), ("volume", C.c_float), ("dim", C.c_int), ("E", C.c_float), ("nu", C.c_float),
                ("logJp0", C.c_float), ("fa", C.c_float), ("cohesion", C.c_float), ("beta", C.c_float),
                ("volumeCorrection", C.c_int), ("yieldSurface", C.c_float)]


class zpc_nacc(C.Structure):
    _fields_ = [("rho", C.c_float), ("volume", C.c_float), ("dim", C.c_int), ("E", C.c_float), ("nu", C.c_float),
                ("logJp0", C.c_float), ("fa", C.c_float), ("xi", C.c_float), ("beta", C.c_float), ("hardeningOn", C.c_int)]


class zpc_equation_of_state(C.Structure):
    _fields_ = [("rho", C.c_float), ("volume", C.c_float), ("dim", C.c_int), ("bulk", C.c_float), ("gamma", C.c_float),
                ("viscosity", C.c_float)]


class zpc_collider(C.Structure):
    _fields_ = [("geometry", C.c_int), ("type", C.c_int), ("origin", C.c_float * 3), ("normal", C.c_float * 3),
                ("b", C.c_float * 3), ("dbdt", C.c_float * 3), ("R", C.c_float * 9), ("omega", C.c_float * 3),
                ("s", C.c_float), ("dsdt", C.c_float)]


class zpc_lbvh_view(C.Structure):
    _fields_ = [("orderedBvs", C.c_void_p), ("auxIndices", C.c_void_p), ("parents", C.c_void_p), ("levels", C.c_void_p),
                ("leafInds", C.c_void_p)]


class zpc_bht_view(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("indices", C.c_void_p), ("status", C.c_void_p), ("activeKeys", C.c_void_p),
                ("tableSize", C.c_uint32), ("numBuckets", C.c_uint32), ("cnt", C.c_void_p), ("success", C.c_void_p),
                ("hf", C.c_uint32 * 6)]


class zpc_sparsegrid_view(C.Structure):
    _fields_ = [("table", zpc_bht_view), ("grid", C.c_void_p), ("numBlocks", C.c_size_t), ("numChannels", C.c_int),
                ("transform", C.c_float * 16), ("background", C.c_float)]


class zpc_bins_view(C.Structure):
    _fields_ = [("pars", zpc_tilevector_view), ("binStart", C.c_void_p), ("binKey", C.c_void_p),
                ("numBins", C.c_void_p), ("binCapacity", C.c_int), ("cellOrder", C.c_void_p),
                ("cellStart", C.c_void_p), ("cellOrderValid", C.c_void_p), ("status", C.c_void_p)]


HALO_K, HALO_MAX_PEERS = 4, 16


class zpc_halo_view(C.Structure):
    _fields_ = [("peer", C.c_void_p), ("pos", C.c_void_p), ("world", C.c_int), ("rank", C.c_int), ("seg", C.c_int), ("half", C.c_int),
                ("recv", C.c_void_p), ("peers", C.c_void_p * HALO_MAX_PEERS), ("status", C.c_void_p)]


# bits of zpc_bins_view.status (include/zpcb200.h)
BINS_HOME_BLOCK_MISSING, BINS_BIN_CAPACITY, BINS_BLOCK_CAPACITY, BINS_STENCIL_BLOCK_MISSING, BINS_TMA_TIMEOUT = 1, 2, 4, 8, 16


_lib = None


def lib():
    """The C-ABI library.  Fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZpcError("libzpcb200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or python zpc_b200/build.py); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.zpcb200_version.restype = C.c_char_p
        _lib.policy__b200.restype = C.c_void_p
        _lib.zpcb200_bht_table_size.restype = C.c_size_t
        _lib.zpcb200_bht_table_size.argtypes = [C.c_size_t]
    return _lib


def _check(rc, what):
    if rc != 0:
        raise ZpcError("%s failed with code %d" % (what, rc))


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def port(t, idx=0):
    """Iterator port over a contiguous 1-D device tensor (GenericIterator.hpp: aos, numTileBits 0)."""
    if t is None:
        return zpc_port(None, 0, 0, 0, 1)
    assert t.is_cuda and t.is_contiguous()
    return zpc_port(t.data_ptr(), idx, 0, 0, 1)


def vec3_port(x):
    """Port over an AoS [N,3] float tensor seen as vec3 elements."""
    assert x.is_cuda and x.is_contiguous() and x.shape[-1] == 3
    return zpc_port(x.data_ptr(), 0, 0, 0, 3)


_SUFFIX = {torch.int32: "i32", torch.float32: "f32", torch.int64: "i64", torch.uint32: "u32",
           torch.uint64: "u64", torch.float64: "f64"}


class TileVector:
    """AoSoA container (TileVector.hpp:108): element (chn,i) at base[(i//L*nch + chn)*L + i%L]."""

    def __init__(self, size, num_channels, L=32, device="cuda", dtype=torch.float32):
        self.size, self.nch, self.L = int(size), int(num_channels), L
        ntiles = (self.size + L - 1) // L
        self.data = torch.zeros(max(ntiles, 1) * self.nch * L, dtype=dtype, device=device)

    def view(self):
        return zpc_tilevector_view(self.data.data_ptr(), self.size, self.nch)

    def port(self, chn, idx=0):
        """get_iterator_1__tv_<T>_<L>(v, id, chnOffset) (py_interop/TileVectorInstantiations.cpp:24-120)"""
        L = self.L
        bits = L.bit_length() - 1
        return zpc_port(self.data.data_ptr() + chn * L * self.data.element_size(), idx, bits, L - 1, self.nch)

    def channel(self, chn, width=1):
        """host-side gather of channels [chn, chn+width) as a [size, width] tensor (for tests)."""
        L = self.L
        t = self.data.view(-1, self.nch, L)[:, chn:chn + width, :]          # [tiles, width, L]
        return t.permute(0, 2, 1).reshape(-1, width)[: self.size].contiguous()

    def set_channel(self, chn, values):
        L = self.L
        values = values.reshape(self.size, -1)
        width = values.shape[1]
        ntiles = self.data.numel() // (self.nch * L)
        pad = torch.zeros(ntiles * L, width, dtype=self.data.dtype, device=self.data.device)
        pad[: self.size] = values
        self.data.view(ntiles, self.nch, L)[:, chn:chn + width, :] = pad.view(ntiles, L, width).permute(0, 2, 1)


class CudaExecutionPolicy:
    """cuda_exec(): reduce / scans / radix sorts on B200.  Inputs are contiguous device tensors or
    (TileVector, channel) pairs; outputs likewise.  Scratch is cached per policy."""

    def __init__(self):
        self._device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self._stream = None
        self._sync = True
        self._scratch = None

    def device(self, i):
        self._device = int(i)
        return self

    def stream(self, s):
        self._stream = s
        return self

    def sync(self, b):
        self._sync = bool(b)
        return self

    # -- helpers
    def _tmp(self, nbytes):
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device="cuda:%d" % self._device)
        return self._scratch

    @staticmethod
    def _as_port(a):
        if isinstance(a, tuple):
            tv, chn = a
            return tv.port(chn), tv.data.dtype, tv.size
        return port(a), a.dtype, a.numel()

    def _two_phase(self, fn, *args):
        nbytes = C.c_size_t(0)
        st = _stream_ptr(self._stream)
        _check(fn(None, C.byref(nbytes), *args, st), fn.__name__ + "(size)")
        tmp = self._tmp(nbytes.value)
        cap = C.c_size_t(tmp.numel())
        _check(fn(C.c_void_p(tmp.data_ptr()), C.byref(cap), *args, st), fn.__name__)
        if self._sync:
            (self._stream or torch.cuda.current_stream()).synchronize()

    # -- primitives (ExecutionPolicy.cuh:552-866)
    def reduce(self, src, out, op="sum"):
        p, dt, n = self._as_port(src)
        po, _, _ = self._as_port(out)
        fn = getattr(lib(), "zpcb200_reduce_%s_%s" % (op, _SUFFIX[dt]))
        self._two_phase(fn, p, po, C.c_size_t(n))

    def exclusive_scan(self, src, out):
        p, dt, n = self._as_port(src)
        po, _, _ = self._as_port(out)
        self._two_phase(getattr(lib(), "zpcb200_exclusive_scan_sum_" + _SUFFIX[dt]), p, po, C.c_size_t(n))

    def inclusive_scan(self, src, out):
        p, dt, n = self._as_port(src)
        po, _, _ = self._as_port(out)
        self._two_phase(getattr(lib(), "zpcb200_inclusive_scan_sum_" + _SUFFIX[dt]), p, po, C.c_size_t(n))

    def radix_sort_pair(self, keys_in, vals_in, keys_out, vals_out, count=None, sbit=0, ebit=None, kind=None):
        pk, dt, n = self._as_port(keys_in)
        pv, _, _ = self._as_port(vals_in)
        pko, _, _ = self._as_port(keys_out)
        pvo, _, _ = self._as_port(vals_out)
        kind = kind or _SUFFIX[dt]
        n = n if count is None else count
        ebit = {"u32": 32, "i32": 32, "u64": 64}[kind] if ebit is None else ebit
        self._two_phase(getattr(lib(), "zpcb200_radix_sort_pair_" + kind), pk, pv, pko, pvo, C.c_size_t(n),
                        C.c_int(sbit), C.c_int(ebit))

    def radix_sort(self, keys_in, keys_out, sbit=0, ebit=None, kind=None):
        pk, dt, n = self._as_port(keys_in)
        pko, _, _ = self._as_port(keys_out)
        kind = kind or _SUFFIX[dt]
        ebit = {"u32": 32, "i32": 32, "u64": 64}[kind] if ebit is None else ebit
        self._two_phase(getattr(lib(), "zpcb200_radix_sort_" + kind), pk, pko, C.c_size_t(n), C.c_int(sbit),
                        C.c_int(ebit))


    def merge_sort_pair(self, keys, vals, count=None):
        """stable ascending sort of (keys, vals) IN PLACE (ExecutionPolicy.cuh:701-760); keys i32 | f32 | f64, vals i32"""
        pk, dt, n = self._as_port(keys)
        pv, _, _ = self._as_port(vals)
        n = n if count is None else count
        self._two_phase(getattr(lib(), "zpcb200_merge_sort_pair_" + _SUFFIX[dt]), pk, pv, C.c_size_t(n))

    def merge_sort(self, keys):
        pk, dt, n = self._as_port(keys)
        self._two_phase(getattr(lib(), "zpcb200_merge_sort_" + _SUFFIX[dt]), pk, C.c_size_t(n))


def cuda_exec():
    return CudaExecutionPolicy()


def next_2pow(n):
    p = 1
    while p < n:
        p <<= 1
    return p


class HashTable:
    """HashTable<i32,3,int>: tableSize = next_2pow(expected) * 16 (HashTable.hpp:70,87-90)."""

    def __init__(self, expected_entries, device="cuda"):
        self.table_size = next_2pow(int(expected_entries)) * 16
        ts = self.table_size
        self.keys = torch.empty(ts, 3, dtype=torch.int32, device=device)
        self.indices = torch.empty(ts, dtype=torch.int32, device=device)
        self.status = torch.empty(ts, dtype=torch.int32, device=device)
        self.active_keys = torch.zeros(ts, 3, dtype=torch.int32, device=device)
        self.cnt = torch.zeros(1, dtype=torch.int32, device=device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)

    def view(self):
        return zpc_hashtable_view(self.keys.data_ptr(), self.indices.data_ptr(), self.status.data_ptr(),
                                  self.active_keys.data_ptr(), self.table_size, self.cnt.data_ptr())

    def size(self):
        return int(self.cnt.item())  # the one D2H read the reference also does (HashTable.hpp:152)


class Grids:
    """Grids<f32,3,4> collocated grid: one [7][64] tile per block, channels {m, v(3), rhs(3)}."""

    def __init__(self, dx, num_blocks, num_channels=7, device="cuda"):
        self.dx, self.num_blocks, self.nch = float(dx), int(num_blocks), num_channels
        self.tiles = torch.zeros(self.num_blocks, num_channels, 64, dtype=torch.float32, device=device)

    def view(self):
        return zpc_grids_view(self.tiles.data_ptr(), self.num_blocks, self.nch, self.dx)


class Bht:
    """bht<i32,3,int,16> (container/Bht.hpp:16-272): buckets of 16, three universal hashes from std::mt19937(2),
    tableSize = evaluateTableSize(expected) (:154-158), 16-byte key slots."""

    def __init__(self, expected_entries, device="cuda"):
        self.table_size = int(lib().zpcb200_bht_table_size(int(expected_entries)))
        ts = max(self.table_size, 1)
        hf = (C.c_uint32 * 6)()
        lib().zpcb200_bht_params(hf)
        self.hf = [int(v) for v in hf]
        self.keys = torch.empty(ts, 4, dtype=torch.int32, device=device)
        self.indices = torch.empty(ts, dtype=torch.int32, device=device)
        self.status = torch.empty(ts, dtype=torch.int32, device=device)
        self.active_keys = torch.zeros(ts, 3, dtype=torch.int32, device=device)
        self.cnt = torch.zeros(1, dtype=torch.int32, device=device)
        self.success = torch.ones(1, dtype=torch.int32, device=device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)

    def view(self):
        return zpc_bht_view(self.keys.data_ptr(), self.indices.data_ptr(), self.status.data_ptr(), self.active_keys.data_ptr(),
                            self.table_size, self.table_size // 16, self.cnt.data_ptr(), self.success.data_ptr(),
                            (C.c_uint32 * 6)(*self.hf))

    def size(self):
        return int(self.cnt.item())


class SparseGrid:
    """SparseGrid<3,f32,8> (geometry/SparseGrid.hpp:16-188): bht keyed by block origins + TileVector<f32,512> + 4x4
    index-to-world transform (row-vector convention) + background value."""

    def __init__(self, num_channels, num_blocks, device="cuda"):
        self.nch, self.num_blocks = int(num_channels), int(num_blocks)
        self.table = Bht(num_blocks, device)
        self.grid = torch.zeros(max(self.num_blocks, 1), self.nch, 512, dtype=torch.float32, device=device)
        self.transform = [1.0 if i % 5 == 0 else 0.0 for i in range(16)]
        self.background = 0.0

    def scale(self, s):
        """SparseGrid::scale -> Transform::preScale (uniform)"""
        import numpy as np
        m = np.array(self.transform, np.float32).reshape(4, 4)
        sm = np.eye(4, dtype=np.float32)
        sm[0, 0] = sm[1, 1] = sm[2, 2] = np.float32(s)
        self.transform = [float(v) for v in (sm @ m).reshape(-1)]

    def translate(self, t):
        """SparseGrid::translate -> Transform::postTranslate"""
        for d in range(3):
            self.transform[12 + d] += float(t[d])

    def view(self):
        return zpc_sparsegrid_view(self.table.view(), self.grid.data_ptr(), self.num_blocks, self.nch,
                                   (C.c_float * 16)(*self.transform), self.background)

    def num_active_blocks(self):
        return self.table.size()


def reorder_tiles(src, dst, num_channels, tile_length, map_, scatter=False, stream=None):
    """TileVector::reorderTiles on raw tile storage (src, dst: float32 device tensors of whole tiles; map_: int32)"""
    _check(lib().zpcb200_tilevector_reorder_tiles(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), C.c_int(num_channels),
                                                  C.c_int(tile_length), C.c_void_p(map_.data_ptr()), C.c_size_t(map_.numel()),
                                                  C.c_int(int(scatter)), _stream_ptr(stream)), "reorder_tiles")


def bht_reorder(table, map_, scatter=False, stream=None):
    """bht::reorder: renumber the table through map_ (int32 device tensor of table.size() entries)"""
    n = map_.numel()
    ordered = torch.zeros_like(table.active_keys)
    _check(lib().zpcb200_bht_reorder(table.view(), C.c_void_p(map_.data_ptr()), C.c_int(n), C.c_int(int(scatter)),
                                     C.c_void_p(ordered.data_ptr()), _stream_ptr(stream)), "bht_reorder")
    table.active_keys = ordered


def sg_reorder_morton(sg, stream=None):
    """put the active blocks of a SparseGrid in Morton order (table numbering + grid tiles); returns the gather map"""
    n = sg.table.size()
    map_ = torch.empty(max(n, 1), dtype=torch.int32, device=sg.grid.device)
    _two_phase(lib().zpcb200_sg_morton_order, (sg.view(), C.c_int(n), C.c_void_p(map_.data_ptr()),
                                               C.c_void_p(sg.table.overflow.data_ptr())), (), stream)
    map_ = map_[:n]
    if n:
        bht_reorder(sg.table, map_, False, stream)
        new_grid = torch.zeros_like(sg.grid)
        reorder_tiles(sg.grid, new_grid, sg.nch, 512, map_, False, stream)
        sg.grid = new_grid
    return map_


def sg_partition_for_particles(x_port, n, sg, stream=None, enlarge=(0, 2), wide=False):
    _two_phase(lib().zpcb200_sg_partition_build_wide if wide else lib().zpcb200_sg_partition_build, (x_port, C.c_size_t(n), sg.view(), C.c_int(enlarge[0]), C.c_int(enlarge[1]),
                                                  C.c_void_p(sg.table.overflow.data_ptr())), (), stream)


def sg_clean(sg, stream=None):
    _check(lib().zpcb200_sg_clean(sg.view(), _stream_ptr(stream)), "sg_clean")


def sg_bin_particles(pars, sg, bins, order_out=None, stream=None):
    """block-binned fast path on the SparseGrid: bins = octants (4^3 cells) of the side-8 blocks"""
    _two_phase(lib().zpcb200_sg_bin_particles, (pars.view(), sg.view(), bins.view(),
                                                C.c_void_p(order_out.data_ptr() if order_out is not None else None)), (), stream)


def sg_rebin_particles(src, sg, dst, stream=None, order_out=None):
    _two_phase(lib().zpcb200_sg_rebin_particles, (src.view(), sg.view(), dst.view(),
                                                  C.c_void_p(order_out.data_ptr() if order_out is not None else None)), (), stream)


def sg_p2g_transfer(pars, sg, dt, model, stream=None):
    if isinstance(pars, ParticleBins):
        if isinstance(model, zpc_fixed_corotated):
            _check(lib().zpcb200_sg_p2g_apic_fcr_binned(pars.view(), sg.view(), C.c_float(dt), model, _stream_ptr(stream)), "sg_p2g(binned)")
            return
        kind = _MODEL_KIND[type(model)]
        side = pars.logJp if kind in (2, 3) else pars.J if kind == 4 else None
        if kind >= 2 and side is None:
            raise ValueError("this model needs its per-particle scalar (logJp / J) in bin order next to the bins")
        _check(lib().zpcb200_sg_p2g_apic_model_binned(pars.view(), C.c_void_p(side.data_ptr() if side is not None else None), sg.view(),
                                                      C.c_float(dt), C.c_int(kind), C.byref(model), _stream_ptr(stream)), "sg_p2g(model, binned)")
        return
    if isinstance(model, zpc_fixed_corotated):
        _check(lib().zpcb200_sg_p2g_apic_fcr(pars.view(), sg.view(), C.c_float(dt), model, _stream_ptr(stream)), "sg_p2g")
        return
    kind = {zpc_vonmises_fixed_corotated: 1, zpc_drucker_prager: 2, zpc_nacc: 3, zpc_equation_of_state: 4}[type(model)]
    _check(lib().zpcb200_sg_p2g_apic_model(pars.view(), sg.view(), C.c_float(dt), C.c_int(kind), C.byref(model), _stream_ptr(stream)),
           "sg_p2g(model)")


def sg_compute_grid_velocity(sg, dt, extf, mode, max_vel_sqr, stream=None):
    e = (C.c_float * 3)(*[float(v) for v in extf])
    _check(lib().zpcb200_sg_grid_update(sg.view(), C.c_float(dt), e, C.c_int(mode), C.c_void_p(max_vel_sqr.data_ptr()),
                                        _stream_ptr(stream)), "sg_grid_update")


def sg_g2p_transfer(pars, sg, dt, stream=None, model=None):
    if isinstance(pars, ParticleBins):
        if isinstance(model, zpc_equation_of_state):
            _check(lib().zpcb200_sg_g2p_apic_eos_binned(pars.view(), C.c_void_p(pars.J.data_ptr()), sg.view(), C.c_float(dt), _stream_ptr(stream)),
                   "sg_g2p(eos, binned)")
            return
        _check(lib().zpcb200_sg_g2p_apic_binned(pars.view(), sg.view(), C.c_float(dt), _stream_ptr(stream)), "sg_g2p(binned)")
        return
    if isinstance(model, zpc_equation_of_state):
        _check(lib().zpcb200_sg_g2p_apic_eos(pars.view(), sg.view(), C.c_float(dt), _stream_ptr(stream)), "sg_g2p(eos)")
        return
    _check(lib().zpcb200_sg_g2p_apic(pars.view(), sg.view(), C.c_float(dt), _stream_ptr(stream)), "sg_g2p")


def sg_value_or(sg, chn, coords, dflt, stream=None):
    """SparseGridView::valueOr(false_c, chn, coord, default) at an [n,3] int32 device tensor of index coordinates"""
    assert coords.is_cuda and coords.dtype == torch.int32 and coords.is_contiguous()
    out = torch.empty(coords.shape[0], dtype=torch.float32, device=coords.device)
    _check(lib().zpcb200_sg_value_or(sg.view(), C.c_int(chn), C.c_void_p(coords.data_ptr()), C.c_size_t(coords.shape[0]),
                                     C.c_float(dflt), C.c_void_p(out.data_ptr()), _stream_ptr(stream)), "sg_value_or")
    return out


def sg_cell_coords(sg, bno, cno, stream=None):
    """(iCoord, wCoord) of n (block, cell) pairs (int32 device tensors)"""
    n = bno.shape[0]
    ic = torch.empty(n, 3, dtype=torch.int32, device=bno.device)
    wc = torch.empty(n, 3, dtype=torch.float32, device=bno.device)
    _check(lib().zpcb200_sg_cell_coords(sg.view(), C.c_void_p(bno.data_ptr()), C.c_void_p(cno.data_ptr()), C.c_size_t(n),
                                        C.c_void_p(ic.data_ptr()), C.c_void_p(wc.data_ptr()), _stream_ptr(stream)), "sg_cell_coords")
    return ic, wc


class Particles:
    """Particles<f32,3>: AoS attribute arrays x,v (vec3), m, C,F (column-major vec9)."""

    def __init__(self, P, device="cuda"):
        self.n = int(P["x"].shape[0])
        self.x = torch.as_tensor(P["x"]).to(device).contiguous()
        self.v = torch.as_tensor(P["v"]).to(device).contiguous()
        self.m = torch.as_tensor(P["m"]).to(device).contiguous()
        self.C = torch.as_tensor(P["C"]).to(device).contiguous()
        self.F = (torch.as_tensor(P["F"]).to(device).contiguous() if "F" in P          # the equation of state carries J instead
                  else torch.eye(3, dtype=torch.float32, device=device).reshape(1, 9).repeat(self.n, 1))
        self.J = torch.as_tensor(P["J"]).to(device).contiguous() if "J" in P else None   # EquationOfStateConfig only
        self.logJp = torch.as_tensor(P["logJp"]).to(device).contiguous() if "logJp" in P else None  # DruckerPrager / NACC

    def view(self, lo=0, hi=None):
        """ParticlesView over particles [lo, hi) (default: all) — the attribute arrays are AoS, a range is a pointer offset"""
        hi = self.n if hi is None else hi
        if not (0 <= lo <= hi <= self.n):
            raise ValueError("particle range [%d, %d) outside [0, %d)" % (lo, hi, self.n))
        def p(t, w):
            return t.data_ptr() + 4 * w * lo if t is not None else None
        return zpc_particles_view(p(self.m, 1), p(self.x, 3), p(self.v, 3), None, p(self.J, 1), p(self.F, 9), p(self.C, 9),
                                  p(self.logJp, 1), hi - lo)

    def range(self, lo, hi):
        """a non-owning sub-range usable wherever Particles is (p2g_transfer / g2p_transfer on a chunk)"""
        return _ParticleRange(self, lo, hi)

    def to_host(self):
        return {k: getattr(self, k).cpu().numpy() for k in ("x", "v", "m", "C", "F")}


class _ParticleRange:
    def __init__(self, pars, lo, hi):
        self._p, self._lo, self._hi = pars, int(lo), int(hi)
        self.n = self._hi - self._lo
        self.J, self.logJp = pars.J, pars.logJp

    def view(self):
        return self._p.view(self._lo, self._hi)


class ParticleBins:
    """Block-binned AoSoA particles (include/zpcb200.h: zpc_bins_view)."""

    def __init__(self, n, bin_capacity, device="cuda", cell_order_cache=True):
        self.n = int(n)
        self.pars = TileVector(n, PB_NCH, 32, device)
        self.cap = int(bin_capacity)
        self.bin_start = torch.zeros(self.cap + 1, dtype=torch.int32, device=device)
        self.bin_key = torch.zeros(self.cap, 3, dtype=torch.int32, device=device)
        self.num_bins = torch.zeros(1, dtype=torch.int32, device=device)
        self.logJp = None        # plastic models: one float per particle in bin order (set by the owner, permuted with every re-bin)
        self.J = None            # equation of state: likewise
        self.cell_order = self.cell_start = self.cell_order_valid = None
        if cell_order_cache:
            self.cell_order = torch.zeros(max(self.n, 1), dtype=torch.int16, device=device)
            self.cell_start = torch.zeros(self.cap * 224, dtype=torch.int16, device=device)
            self.cell_order_valid = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)   # ZPC_BINS_* bits, ORed in by the binned entries

    def view(self):
        co = self.cell_order
        return zpc_bins_view(self.pars.view(), self.bin_start.data_ptr(), self.bin_key.data_ptr(),
                             self.num_bins.data_ptr(), self.cap,
                             co.data_ptr() if co is not None else None,
                             self.cell_start.data_ptr() if co is not None else None,
                             self.cell_order_valid.data_ptr() if co is not None else None,
                             self.status.data_ptr())

    def check_status(self, what="binned path"):
        """D2H read of the status word; raises on any ZPC_BINS_* bit (and clears it)"""
        st = int(self.status.item())
        if st:
            self.status.zero_()
            raise RuntimeError("%s: %s" % (what, bins_status_text(st)))

    def attr(self, name):
        chn, w = {"m": (PB_M, 1), "x": (PB_X, 3), "v": (PB_V, 3), "C": (PB_C, 9), "F": (PB_F, 9)}[name]
        t = self.pars.channel(chn, w)
        return t[:, 0].contiguous() if w == 1 else t


def bins_status_text(st):
    """the ZPC_BINS_* bits of a zpc_bins_view.status word (include/zpcb200.h) in words"""
    names = [n for b, n in ((BINS_HOME_BLOCK_MISSING, "a particle's home block is not in the partition"),
                            (BINS_BIN_CAPACITY, "more bins than binCapacity"), (BINS_BLOCK_CAPACITY, "more blocks than binCapacity"),
                            (BINS_STENCIL_BLOCK_MISSING, "a stencil block is absent from the partition (particle drifted past the extra "
                                                         "ring: re-bin more often)"),
                            (BINS_TMA_TIMEOUT, "a TMA transaction of the staged G2P did not complete")) if st & b]
    return "; ".join(names) or "status %d" % st


def set_tuning(p2g_sweep=-1, g2p_staged=-1):
    """zpcb200_set_tuning: pick the kernel variant of the binned P2G sweep (4 | 3 | 6 = atomic-free plane sweep) and of the binned G2P (1 = TMA-staged
    particles | 0 = plain loads); -1 keeps a setting."""
    rc = lib().zpcb200_set_tuning(int(p2g_sweep), int(g2p_staged))
    if rc:
        raise RuntimeError("zpcb200_set_tuning(%d, %d) -> %d" % (p2g_sweep, g2p_staged, rc))


def get_tuning():
    a, b = C.c_int(0), C.c_int(0)
    lib().zpcb200_get_tuning(C.byref(a), C.byref(b))
    return dict(p2g_sweep=a.value, g2p_staged=b.value)


def model_fcr(volume, E=5.0e4, nu=0.4, rho=1000.0):
    return zpc_fixed_corotated(rho, volume, 3, E, nu)


class _Scratch:
    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self.buf


_scratch = _Scratch()


def _two_phase(fn, args_before, args_after, stream, device="cuda"):
    nbytes = C.c_size_t(0)
    st = _stream_ptr(stream)
    _check(fn(None, C.byref(nbytes), *args_before, *args_after, st), fn.__name__ + "(size)")
    tmp = _scratch.get(nbytes.value, device)
    cap = C.c_size_t(tmp.numel())
    _check(fn(C.c_void_p(tmp.data_ptr()), C.byref(cap), *args_before, *args_after, st), fn.__name__)


class IndexBuckets:
    """IndexBuckets<3, i32, int> (container/IndexBuckets.hpp:12-59): _table of occupied cells, _counts / _offsets / _indices."""

    def __init__(self, n, dx, device="cuda", expected_cells=None):
        self.n, self.dx, self.device = int(n), float(dx), device
        # Query.tpp:24 sizes the table for pars.size() entries (x16 slots each); pass expected_cells when far fewer are occupied
        self.table = HashTable(max(int(expected_cells) if expected_cells else self.n, 1), device)
        self.counts = torch.zeros(self.n + 1, dtype=torch.int32, device=device)
        self.offsets = torch.zeros(self.n + 1, dtype=torch.int32, device=device)
        self.indices = torch.full((max(self.n, 1),), -1, dtype=torch.int32, device=device)

    def num_buckets(self):
        return self.table.size()


def index_buckets_for_particles(x_port, n, dx, displacement=0.5, stream=None, device="cuda", expected_cells=None, wide=False):
    """index_buckets_for_particles(policy, particles, dx, displacement) (simulation/particle/Query.tpp:9-58)"""
    ib = IndexBuckets(n, dx, device, expected_cells)
    _two_phase(lib().zpcb200_index_buckets_build_wide if wide else lib().zpcb200_index_buckets_build, (x_port, C.c_size_t(n), C.c_float(dx), C.c_float(displacement), ib.table.view(),
                                                   C.c_void_p(ib.counts.data_ptr()), C.c_void_p(ib.offsets.data_ptr()),
                                                   C.c_void_p(ib.indices.data_ptr()), C.c_void_p(ib.table.overflow.data_ptr())), (),
               stream, device)
    return ib


class LBvh:
    """LBvh<3, int, f32> (container/Bvh.hpp:82-174): orderedBvs / auxIndices / parents / levels in DFS pre-order, leafInds.
    build(bvs) / refit(bvs) take a float32 [n, 6] device tensor of boxes {min, max}."""

    def __init__(self, device="cuda"):
        self.device = device
        self.n = 0
        self.orderedBvs = self.auxIndices = self.parents = self.levels = self.leafInds = None

    def num_nodes(self):
        return 2 * self.n - 1 if self.n > 2 else self.n

    def view(self):
        return zpc_lbvh_view(*[t.data_ptr() if t is not None else None
                               for t in (self.orderedBvs, self.auxIndices, self.parents, self.levels, self.leafInds)])

    def build(self, bvs, refit=True, stream=None):
        assert bvs.dtype == torch.float32 and bvs.is_contiguous() and bvs.shape[-1] == 6
        self.n = int(bvs.shape[0])
        nn = max(self.num_nodes(), 1)
        self.orderedBvs = torch.zeros(nn, 6, dtype=torch.float32, device=self.device)
        self.auxIndices, self.parents, self.levels = (torch.full((nn,), -7, dtype=torch.int32, device=self.device) for _ in range(3))
        self.leafInds = torch.full((max(self.n, 1),), -7, dtype=torch.int32, device=self.device)
        _two_phase(lib().zpcb200_lbvh_build, (C.c_void_p(bvs.data_ptr()), C.c_size_t(self.n), self.view(), C.c_int(int(refit))), (),
                   stream, self.device)
        return self

    def query(self, query_bvs, stream=None):
        """batched iter_neighbors: returns (offsets[nq + 1], ids) — ids[offsets[q]:offsets[q+1]] overlap query box q"""
        nq = int(query_bvs.shape[0])
        counts = torch.zeros(nq + 1, dtype=torch.int32, device=self.device)
        offsets = torch.zeros(nq + 1, dtype=torch.int32, device=self.device)
        st = _stream_ptr(stream)
        q = C.c_void_p(query_bvs.data_ptr())
        _check(lib().zpcb200_lbvh_query(self.view(), C.c_size_t(self.n), q, C.c_size_t(nq), C.c_void_p(counts.data_ptr()), None, None, st),
               "lbvh_query(count)")
        _two_phase(lib().zpcb200_exclusive_scan_sum_i32, (port(counts), port(offsets), C.c_size_t(nq + 1)), (), stream, self.device)
        total = int(offsets[nq].item())
        ids = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        _check(lib().zpcb200_lbvh_query(self.view(), C.c_size_t(self.n), q, C.c_size_t(nq), None, C.c_void_p(offsets.data_ptr()),
                                        C.c_void_p(ids.data_ptr()), st), "lbvh_query(fill)")
        return offsets, ids[:total]

    def refit(self, bvs, stream=None):
        if int(bvs.shape[0]) != self.n:
            raise RuntimeError("bvh topology changes, require rebuild!")          # Bvh.hpp:1239-1240
        _two_phase(lib().zpcb200_lbvh_refit, (C.c_void_p(bvs.data_ptr()), C.c_size_t(self.n), self.view()), (), stream, self.device)


# ---- functors as free functions (policy = stream holder; all calls asynchronous) -------------------
def partition_for_particles(x_port, n, dx, table, stream=None, enlarge=(0, 2), wide=False):
    """SparsityCompute.tpp:6-24 / SparsityOp.hpp:41-112 (CleanSparsity, ComputeSparsity, EnlargeSparsity{lo,hi}).
    wide: 64-bit block codes (block coordinates up to +-2^20 instead of +-512)."""
    _two_phase(lib().zpcb200_partition_build_wide if wide else lib().zpcb200_partition_build, (x_port, C.c_size_t(n), C.c_float(dx), table.view(),
                                               C.c_int(enlarge[0]), C.c_int(enlarge[1]),
                                               C.c_void_p(table.overflow.data_ptr())), (), stream)


def clean_grid_blocks(grids, table, stream=None):
    _check(lib().zpcb200_clean_grid(grids.view(), C.c_void_p(table.cnt.data_ptr()), _stream_ptr(stream)),
           "clean_grid")


def model_eos(volume, bulk=4.0e4, gamma=7.15, viscosity=0.0, rho=1000.0):
    return zpc_equation_of_state(rho, volume, 3, bulk, gamma, viscosity)


def model_vonmises(volume, E=5.0e4, nu=0.4, yield_stress=240e6, rho=1000.0):
    return zpc_vonmises_fixed_corotated(rho, volume, 3, E, nu, yield_stress)


DRUCKER_PRAGER_YIELD_SURFACE = 0.816496580927726 * 2.0 * 0.5 / (3.0 - 0.5)   # ConstitutiveModel.hpp:756


def model_drucker_prager(volume, E=5.0e4, nu=0.4, cohesion=0.0, beta=1.0, volume_correction=True,
                         yield_surface=DRUCKER_PRAGER_YIELD_SURFACE, logJp0=0.0, fa=30.0, rho=1000.0):
    """DruckerPragerConfig (physics/ConstitutiveModel.hpp:748-757)"""
    return zpc_drucker_prager(rho, volume, 3, E, nu, logJp0, fa, cohesion, beta, int(bool(volume_correction)), yield_surface)


def model_nacc(volume, E=5.0e4, nu=0.4, fa=45.0, xi=0.8, beta=0.5, hardening_on=True, logJp0=-0.01, rho=1000.0):
    """NACCConfig (physics/ConstitutiveModel.hpp:758-776)"""
    return zpc_nacc(rho, volume, 3, E, nu, logJp0, fa, xi, beta, int(bool(hardening_on)))


def p2g_transfer(pars, table, grids, dt, model, stream=None):
    if isinstance(model, (zpc_drucker_prager, zpc_nacc)):
        if getattr(pars, "logJp", None) is None:
            raise ValueError("the plastic models need the per-particle logJp attribute (P2G.hpp:93)")
        dp = isinstance(model, zpc_drucker_prager)
        if isinstance(pars, ParticleBins):          # logJp: one float per particle in bin order, kept next to the bins
            fn = lib().zpcb200_p2g_apic_drucker_prager_binned if dp else lib().zpcb200_p2g_apic_nacc_binned
            _check(fn(pars.view(), C.c_void_p(pars.logJp.data_ptr()), table.view(), grids.view(), C.c_float(dt), model, _stream_ptr(stream)),
                   "p2g(plastic, binned)")
            return
        fn = lib().zpcb200_p2g_apic_drucker_prager if dp else lib().zpcb200_p2g_apic_nacc
        _check(fn(pars.view(), table.view(), grids.view(), C.c_float(dt), model, _stream_ptr(stream)), "p2g(plastic)")
        return
    if isinstance(model, zpc_vonmises_fixed_corotated) and isinstance(pars, ParticleBins):
        _check(lib().zpcb200_p2g_apic_vonmises_binned(pars.view(), table.view(), grids.view(), C.c_float(dt), model,
                                                      _stream_ptr(stream)), "p2g(vonmises, binned)")
        return
    if isinstance(model, zpc_vonmises_fixed_corotated):
        _check(lib().zpcb200_p2g_apic_vonmises(pars.view(), table.view(), grids.view(), C.c_float(dt), model,
                                               _stream_ptr(stream)), "p2g(vonmises)")
        return
    if isinstance(model, zpc_equation_of_state) and isinstance(pars, ParticleBins):
        if pars.J is None:
            raise ValueError("the equation of state needs the per-particle J attribute (P2G.hpp:67)")
        _check(lib().zpcb200_p2g_apic_eos_binned(pars.view(), C.c_void_p(pars.J.data_ptr()), table.view(), grids.view(), C.c_float(dt), model,
                                                 _stream_ptr(stream)), "p2g(eos, binned)")
        return
    if isinstance(model, zpc_equation_of_state):
        _check(lib().zpcb200_p2g_apic_eos(pars.view(), table.view(), grids.view(), C.c_float(dt), model,
                                          _stream_ptr(stream)), "p2g(eos)")
        return
    if isinstance(pars, ParticleBins):
        rc = lib().zpcb200_p2g_apic_fcr_binned(pars.view(), table.view(), grids.view(), C.c_float(dt), model,
                                               _stream_ptr(stream))
    else:
        rc = lib().zpcb200_p2g_apic_fcr(pars.view(), table.view(), grids.view(), C.c_float(dt), model,
                                        _stream_ptr(stream))
    _check(rc, "p2g")


_MODEL_KIND = {zpc_fixed_corotated: 0, zpc_vonmises_fixed_corotated: 1, zpc_drucker_prager: 2, zpc_nacc: 3, zpc_equation_of_state: 4}


def g2p2g_transfer(pars, table, dx, dt, model, gridv, gridr, stream=None):
    """G2P2GTransfer{cuda_c, wrapv<apic>{}, dt, model, grid, x, r, table, particles} (simulation/transfer/G2P2G.hpp:49-141): gridv, gridr =
    float32 [numBlocks * 64, 3] DOF vectors; the force terms are ADDED to gridr"""
    assert gridv.dtype == torch.float32 and gridr.dtype == torch.float32 and gridv.is_contiguous() and gridr.is_contiguous()
    _check(lib().zpcb200_g2p2g_apic(pars.view(), table.view(), C.c_float(dx), C.c_float(dt), C.c_int(_MODEL_KIND[type(model)]), C.byref(model),
                                    C.c_void_p(gridv.data_ptr()), C.c_void_p(gridr.data_ptr()), _stream_ptr(stream)), "g2p2g")


def compute_grid_block_velocity(grids, table, dt, extf, mode, max_vel_sqr, stream=None):
    e = (C.c_float * 3)(*[float(v) for v in extf])
    _check(lib().zpcb200_grid_update(grids.view(), C.c_void_p(table.cnt.data_ptr()), C.c_float(dt), e,
                                     C.c_int(mode), C.c_void_p(max_vel_sqr.data_ptr()), _stream_ptr(stream)),
           "grid_update")


def grid_momentum_to_velocity(grids, table, max_vel_sqr, m_chn=0, mv_chn=1, stream=None):
    """GridMomentumToVelocity{cuda_c, grid, mChn, mvChn, maxVel} (GridOp.hpp:184-214): v = mv / m, no gravity"""
    _check(lib().zpcb200_grid_momentum_to_velocity(grids.view(), C.c_void_p(table.cnt.data_ptr()), C.c_int(m_chn), C.c_int(mv_chn),
                                                   C.c_void_p(max_vel_sqr.data_ptr()), _stream_ptr(stream)), "grid_momentum_to_velocity")


def grid_angular_momentum(grids, table, sum6, m_chn=0, mv_chn=1, stream=None):
    """GridAngularMomentum{cuda_c, table, grid, mChn, mvChn, sum} (GridOp.hpp:216-262); sum6: six float64 on the device, added to"""
    assert sum6.dtype == torch.float64 and sum6.numel() >= 6 and sum6.is_cuda
    _check(lib().zpcb200_grid_angular_momentum(grids.view(), table.view(), C.c_int(m_chn), C.c_int(mv_chn),
                                               C.c_void_p(sum6.data_ptr()), _stream_ptr(stream)), "grid_angular_momentum")


GEOM_PLANE, GEOM_SPHERE, GEOM_CUBOID = 0, 1, 2
MAX_COLLIDERS = 4   # ZPCB200_MAX_COLLIDERS
COLLIDER_STICKY, COLLIDER_SLIP, COLLIDER_SEPARATE = 0, 1, 2


def _collider(geom, ctype, origin, normal, translation=(0, 0, 0), velocity=(0, 0, 0), rotation=None, omega=(0, 0, 0), scale=1.0,
              dscale_dt=0.0):
    """Collider{levelset, type, s, dsdt, R, omega, b, dbdt} (geometry/Collider.h:10-24,136-143); rotation = 3x3 row-major"""
    R = [1, 0, 0, 0, 1, 0, 0, 0, 1] if rotation is None else [float(v) for row in rotation for v in row]
    return zpc_collider(geom, ctype, (C.c_float * 3)(*origin), (C.c_float * 3)(*normal), (C.c_float * 3)(*translation),
                        (C.c_float * 3)(*velocity), (C.c_float * 9)(*R), (C.c_float * 3)(*omega), float(scale), float(dscale_dt))


def plane_collider(origin, normal, ctype=COLLIDER_STICKY, **motion):
    return _collider(GEOM_PLANE, ctype, origin, normal, **motion)


def sphere_collider(center, radius, ctype=COLLIDER_STICKY, **motion):
    return _collider(GEOM_SPHERE, ctype, center, (radius, 0.0, 0.0), **motion)


def cuboid_collider(box_min, box_max, ctype=COLLIDER_STICKY, **motion):
    """Collider over AnalyticLevelSet<Cuboid>{min, max} (geometry/AnalyticLevelSet.h:55-126)"""
    return _collider(GEOM_CUBOID, ctype, box_min, box_max, **motion)


def apply_boundary_condition(collider, table, grids, stream=None):
    """ApplyBoundaryConditionOnGridBlocks{cuda_c, collider, table, grids} (GridOp.hpp:112-164)."""
    _check(lib().zpcb200_apply_boundary(grids.view(), table.view(), collider, _stream_ptr(stream)), "apply_boundary")


def compute_grid_block_velocity_with_boundaries(grids, table, dt, extf, mode, colliders, max_vel_sqr, stream=None):
    """ComputeGridBlockVelocity + ApplyBoundaryConditionOnGridBlocks for each collider, fused into one grid pass"""
    e = (C.c_float * 3)(*[float(v) for v in extf])
    arr = (zpc_collider * max(len(colliders), 1))(*colliders)
    _check(lib().zpcb200_grid_update_bc(grids.view(), table.view(), C.c_float(dt), e, C.c_int(mode), arr, C.c_int(len(colliders)),
                                        C.c_void_p(max_vel_sqr.data_ptr()), _stream_ptr(stream)), "grid_update_bc")


def p2g_transfer_halo(bins, table, grids, dt, model, halo, stream=None):
    """binned P2G (fixed-corotated) with the halo send fused into the write-back (zpcb200_p2g_apic_fcr_binned_halo)"""
    _check(lib().zpcb200_p2g_apic_fcr_binned_halo(bins.view(), table.view(), grids.view(), C.c_float(dt), model, halo, _stream_ptr(stream)),
           "p2g(binned, halo)")


def grid_update_halo(grids, table, dt, extf, mode, colliders, max_vel_sqr, halo, stream=None):
    """ComputeGridBlockVelocity (+ colliders) with the halo receive fused in (zpcb200_grid_update_halo)"""
    e = (C.c_float * 3)(*[float(v) for v in extf])
    arr = (zpc_collider * max(len(colliders), 1))(*colliders)
    _check(lib().zpcb200_grid_update_halo(grids.view(), table.view(), C.c_float(dt), e, C.c_int(mode), arr, C.c_int(len(colliders)),
                                          C.c_void_p(max_vel_sqr.data_ptr()), halo, _stream_ptr(stream)), "grid_update_halo")


def halo_codes(table, capacity, codes, stream=None):
    _check(lib().zpcb200_halo_codes(table.view(), C.c_int(capacity), C.c_void_p(codes.data_ptr()), _stream_ptr(stream)), "halo_codes")


def halo_build(all_codes, world, rank, capacity, seg, peer_out, pos_out, status, stream=None):
    _two_phase(lib().zpcb200_halo_build, (C.c_void_p(all_codes.data_ptr()), C.c_int(world), C.c_int(rank), C.c_int(capacity), C.c_int(seg),
                                          C.c_void_p(peer_out.data_ptr()), C.c_void_p(pos_out.data_ptr()), C.c_void_p(status.data_ptr())), (), stream)


def g2p_transfer(pars, table, grids, dt, stream=None, model=None):
    if isinstance(model, zpc_equation_of_state) and isinstance(pars, ParticleBins):
        _check(lib().zpcb200_g2p_apic_eos_binned(pars.view(), C.c_void_p(pars.J.data_ptr()), table.view(), grids.view(), C.c_float(dt),
                                                 _stream_ptr(stream)), "g2p(eos, binned)")
        return
    if isinstance(model, zpc_equation_of_state):
        _check(lib().zpcb200_g2p_apic_eos(pars.view(), table.view(), grids.view(), C.c_float(dt), _stream_ptr(stream)),
               "g2p(eos)")
        return
    if isinstance(pars, ParticleBins):
        rc = lib().zpcb200_g2p_apic_binned(pars.view(), table.view(), grids.view(), C.c_float(dt),
                                           _stream_ptr(stream))
    else:
        rc = lib().zpcb200_g2p_apic(pars.view(), table.view(), grids.view(), C.c_float(dt), _stream_ptr(stream))
    _check(rc, "g2p")


def bin_particles(pars, table, dx, bins, order_out=None, stream=None):
    _two_phase(lib().zpcb200_bin_particles, (pars.view(), table.view(), C.c_float(dx), bins.view(),
                                             C.c_void_p(order_out.data_ptr() if order_out is not None else None)),
               (), stream)


def rebin_particles(src, table, dx, dst, stream=None, order_out=None):
    """order_out (int32 [n], optional): the permutation applied, dst slot i <- src slot order_out[i] (side arrays follow it)"""
    if order_out is None:
        _two_phase(lib().zpcb200_rebin_particles, (src.view(), table.view(), C.c_float(dx), dst.view()), (), stream)
    else:
        _two_phase(lib().zpcb200_rebin_particles_ordered, (src.view(), table.view(), C.c_float(dx), dst.view(),
                                                           C.c_void_p(order_out.data_ptr())), (), stream)


def gather_f32(src, idx, dst, stream=None):
    """dst[i] = src[idx[i]] (zpcb200_gather_f32): permutes a per-particle side array with the order a re-bin returned"""
    _check(lib().zpcb200_gather_f32(C.c_void_p(src.data_ptr()), C.c_void_p(idx.data_ptr()), C.c_void_p(dst.data_ptr()),
                                    C.c_size_t(int(dst.numel())), _stream_ptr(stream)), "gather_f32")


def unbin_particles(bins, pars, stream=None):
    _check(lib().zpcb200_unbin_particles(bins.view(), pars.view(), _stream_ptr(stream)), "unbin")


def kernel_launch_count():
    return int(lib().zpcb200_kernel_launch_count())
