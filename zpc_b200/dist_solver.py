"""Multi-GPU substep: one process per GPU, particles sharded by spatial slab, one grid-block halo exchange per
substep over NCCL (the reference has no counterpart — Specification.md lists it as future work).

Each rank runs the single-GPU solver on its own particles.  A grid block that more than one rank's particles touch
("shared") holds only PARTIAL sums after the local P2G, so after P2G every pair of ranks exchanges the 7 channels of
the blocks they share and adds what it receives; every rank then runs the grid update redundantly on its (now
complete) copies — one message per neighbour per substep instead of two — and G2P stays local.  The CFL scalar
max|v|^2 is an all_reduce(max).  Ownership of particles need not change for correctness: a particle that wanders into
another rank's slab is still right (its blocks simply become shared).  For load balance, migrate() hands every particle
to the rank that owns its current home block (BlockOwnership over the cuts of shard_by_blocks; one all_to_all of counts,
one of the particle records).

The exchange topology (which blocks are shared with which rank) is rebuilt only when the partition is, i.e.
with the re-bin every `rebin_every` substeps (partition="with_rebin").
"""
import torch
import torch.distributed as dist

from . import api
from .solver import MpmSolver

_BIAS = 1 << 20


def pack_keys(keys):
    """int32 [n,3] block keys -> int64 codes whose order is the lexicographic (x,y,z) order"""
    k = keys.to(torch.int64) + _BIAS
    return (k[:, 0] << 42) | (k[:, 1] << 21) | k[:, 2]


def shard_by_blocks(x, dx, world, order="xmajor"):
    """SURVEY §8(e) partitioning of an arbitrary particle cloud (host side, numpy): home block of every particle (the block
    ComputeSparsity assigns: floor_div(floor(x/dx + 0.5) - 2, 4), SparsityOp.hpp:68-79), active blocks sorted x-major (or along
    the Morton curve), exclusive prefix sum of the particles per block, cut into `world` contiguous ranges of (nearly) equal
    PARTICLE counts.  Returns (owner[n] int32 rank of every particle, cuts[world + 1] block-range boundaries, keys[nb, 3] the
    sorted block keys).  A block is never split: every particle of a block has the same owner, so the halo of a rank is the
    one-block ring its stencils reach into."""
    import numpy as np
    x = np.asarray(x, np.float32)
    cell = np.floor(x / np.float32(dx) + np.float32(0.5)).astype(np.int64) - 2
    blk = cell >> 2                                               # arithmetic shift = floor division by 4
    lo = blk.min(0) if len(blk) else np.zeros(3, np.int64)
    rel = blk - lo
    if order == "morton":
        def spread(v):                                            # 21 bits -> every third bit
            v = v.astype(np.uint64) & np.uint64(0x1fffff)
            v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
            v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
            v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
            v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
            v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
            return v
        code = (spread(rel[:, 0]) << np.uint64(2)) | (spread(rel[:, 1]) << np.uint64(1)) | spread(rel[:, 2])
    else:
        ext = rel.max(0) + 1 if len(rel) else np.ones(3, np.int64)
        code = ((rel[:, 0] * ext[1] + rel[:, 1]) * ext[2] + rel[:, 2]).astype(np.uint64)
    ucodes, inverse, counts = np.unique(code, return_inverse=True, return_counts=True)
    first = np.zeros(len(ucodes), np.int64)
    first[inverse[::-1]] = np.arange(len(code))[::-1]             # any representative particle of each block
    keys = blk[first].astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(counts)])            # exclusive prefix sum (+ total)
    n = len(code)
    # block b goes to the rank whose target range [r n / world, (r + 1) n / world) contains the midpoint of its particles
    mid = offsets[:-1] + counts / 2.0
    block_owner = np.minimum((mid * world / max(n, 1)).astype(np.int64), world - 1)
    cuts = np.searchsorted(block_owner, np.arange(world + 1), side="left")
    return block_owner[inverse].astype(np.int32), cuts.astype(np.int64), keys


class BlockOwnership:
    """Which rank owns which block, as a function of the block key alone (so that blocks that become active later have an
    owner too): the x-major code of the key against the boundary codes of the cuts shard_by_blocks made."""

    def __init__(self, keys, cuts, lo=None, ext=None):
        import numpy as np
        keys = np.asarray(keys, np.int64)
        self.lo = keys.min(0) - 64 if lo is None else np.asarray(lo, np.int64)           # room for the cloud to move
        self.ext = (keys.max(0) - self.lo + 65) if ext is None else np.asarray(ext, np.int64)
        codes = self.code(keys)
        assert (np.diff(codes) > 0).all(), "keys must be the x-major sorted block list of shard_by_blocks"
        world = len(cuts) - 1
        # rank r owns the codes in [bounds[r], bounds[r + 1]); first / last rank are open-ended
        self.bounds = np.array([codes[cuts[r]] if 0 < cuts[r] < len(codes) else (0 if r == 0 else np.iinfo(np.int64).max)
                                for r in range(1, world)], np.int64)
        self.world = world

    def code(self, blk):
        rel = blk - self.lo
        return (rel[..., 0] * self.ext[1] + rel[..., 1]) * self.ext[2] + rel[..., 2]

    def owner_of_positions(self, x, dx):
        """x: float32 [n, 3] torch tensor (any device) -> int64 [n] owning rank of every particle's home block"""
        cell = torch.floor(x / dx + 0.5).to(torch.int64) - 2
        blk = cell >> 2
        lo = torch.as_tensor(self.lo, device=x.device)
        ext = torch.as_tensor(self.ext, device=x.device)
        rel = blk - lo
        code = (rel[:, 0] * ext[1] + rel[:, 1]) * ext[2] + rel[:, 2]
        bounds = torch.as_tensor(self.bounds, device=x.device)
        return torch.searchsorted(bounds, code, right=True)


def migrate_particles(attrs, dest, group=None):
    """Moves every particle to the rank dest[i] names.  attrs: dict name -> [n, w] (or [n]) float32 tensors on one device;
    dest: int64 [n].  One all_to_all of the counts, one of the packed records (100 B per particle for x, v, m, C, F: the
    exchange of SURVEY §8(e)); particles that stay are part of the same exchange (rank -> itself), so the result is simply
    what arrives, ordered by source rank and, within a source, in the source's order.  Returns the new attrs dict."""
    world = dist.get_world_size(group)
    names = list(attrs)
    widths = [attrs[k].shape[1] if attrs[k].dim() == 2 else 1 for k in names]
    n = dest.numel()
    packed = torch.cat([attrs[k].reshape(n, -1) for k in names], dim=1) if n else torch.zeros(0, sum(widths), dtype=torch.float32, device=dest.device)
    if dest.is_cuda and n:
        # the library's own stable radix sort on the few bits a rank number has (one pass), not torch.argsort
        from . import api
        keys = dest.to(torch.int32)
        ko, order32 = torch.empty_like(keys), torch.empty_like(keys)
        api.cuda_exec().radix_sort_pair(keys, torch.arange(n, dtype=torch.int32, device=dest.device), ko, order32, kind="i32",
                                        sbit=0, ebit=max(1, (world - 1).bit_length()))
        order = order32.long()
    else:
        order = torch.argsort(dest, stable=True)   # the gloo / CPU tests of the host logic
    send_counts = torch.bincount(dest, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    send = packed[order].contiguous()
    recv = torch.empty(sum(rc), packed.shape[1], dtype=packed.dtype, device=packed.device)
    dist.all_to_all_single(recv, send, rc, sc, group=group)
    out, col = {}, 0
    for k, w in zip(names, widths):
        t = recv[:, col:col + w].contiguous()
        out[k] = t if attrs[k].dim() == 2 else t[:, 0].contiguous()
        col += w
    return out


class HaloExchange:
    """Host-side plumbing of the one-ring grid-block exchange.  `pack(ids, buf)` / `unpack_add(ids, buf)` move
    the tiles listed in ids between the grid and a contiguous buffer; the CUDA versions call the C ABI
    (zpcb200_halo_pack / zpcb200_halo_unpack_add)."""

    def __init__(self, group=None, nch=7, device="cuda", pack=None, unpack_add=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nch = nch
        self.device = device
        self._pack = pack
        self._unpack_add = unpack_add
        self.peers = []       # [(rank, ids int32 tensor, send buf, recv buf)]

    def build(self, active_keys):
        """active_keys: int32 [nb,3] of this rank (rows in ascending key order).  Collective.
        Two all_gathers and one device->host read: every rank's keys are gathered (padded, rows stay sorted), the
        intersections with all peers come from one batched searchsorted, one nonzero splits them per peer."""
        mine = pack_keys(active_keys)
        dev = mine.device
        cnt = torch.tensor([mine.numel()], dtype=torch.int64, device=dev)
        cnts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cnts, cnt, group=self.group)
        cnts_h = cnts.tolist()
        mx = max(max(cnts_h), 1)
        pad = torch.full((mx,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
        pad[: mine.numel()] = mine
        allk = torch.empty(self.world * mx, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allk, pad, group=self.group)
        allk = allk.view(self.world, mx)
        self.peers = []
        if mine.numel() == 0:
            return self.peers
        q = mine.unsqueeze(0).expand(self.world, -1).contiguous()
        pos = torch.searchsorted(allk, q).clamp_(max=mx - 1)
        hit = torch.gather(allk, 1, pos) == q                      # [world, nb]: my block b is also active on rank r
        hit[self.rank] = False
        rows, cols = torch.nonzero(hit, as_tuple=True)             # row-major: per peer, ascending key order
        per = torch.bincount(rows, minlength=self.world).tolist()
        cols = cols.to(torch.int32)
        off = 0
        for r in range(self.world):
            n = per[r]
            if n:
                ids = cols[off:off + n].contiguous()
                self.peers.append((r, ids, torch.empty(n, self.nch, 64, dtype=torch.float32, device=dev),
                                   torch.empty(n, self.nch, 64, dtype=torch.float32, device=dev)))
            off += n
        return self.peers

    def shared_blocks(self):
        return sum(p[1].numel() for p in self.peers)

    def exchange_add(self, grids):
        """after the local P2G: send my partial sums of every shared block, add what the peers send"""
        if not self.peers:
            return
        ops = []
        for q, ids, sbuf, rbuf in self.peers:
            self._pack(grids, ids, sbuf)
            ops.append(dist.P2POp(dist.isend, sbuf, q, group=self.group))
            ops.append(dist.P2POp(dist.irecv, rbuf, q, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for q, ids, sbuf, rbuf in self.peers:      # ascending rank order: the same summation order every step
            self._unpack_add(grids, ids, rbuf)


class HaloExchangeP2P(HaloExchange):
    """Same exchange without NCCL point-to-point: every rank owns a symmetric receive buffer (torch symmetric memory,
    peer-mapped over NVLink); the pack kernel stores the shared tiles STRAIGHT INTO THE PEER'S buffer, one
    device-side barrier makes them visible, the unpack kernel adds them locally.  The two halves of the buffer
    alternate by step so that one barrier per substep suffices."""

    def __init__(self, group=None, nch=7, device="cuda", pack_ptr=None, unpack_add=None, capacity_blocks=4096):
        super().__init__(group, nch, device, None, unpack_add)
        self._pack_ptr = pack_ptr
        self.tile = nch * 64
        self.cap = 0
        self.buf = self.hdl = self.ptrs = None
        self.step = 0
        self.plan = []   # [(peer, ids, n, offset in the peer's buffer, offset in my buffer)]  offsets in tiles
        self._allocate(int(capacity_blocks))

    def _allocate(self, capacity_blocks):
        """(re)allocate the symmetric receive buffer; collective (every rank calls it with the same capacity)"""
        import torch.distributed._symmetric_memory as symm_mem
        self.cap = capacity_blocks
        self.buf = symm_mem.empty(2 * self.cap * self.tile, dtype=torch.float32, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD if self.group is None else self.group)
        self.ptrs = list(self.hdl.buffer_ptrs)

    def build(self, active_keys):
        peers = super().build(active_keys)            # reuses the key exchange; drops the NCCL staging buffers below
        dev = active_keys.device
        row = torch.zeros(self.world, dtype=torch.int64, device=dev)
        for q, ids, _, _ in peers:
            row[q] = ids.numel()
        mat = torch.zeros(self.world * self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(mat, row, group=self.group)
        M = mat.view(self.world, self.world).tolist()  # M[a][b] = number of blocks ranks a and b share
        need = max(sum(r) for r in M)                  # the same number on every rank
        if need > self.cap:
            torch.cuda.synchronize()
            self._allocate(int(need * 1.5) + 64)
        self.plan = []
        for q, ids, _, _ in peers:
            off_in_peer = sum(M[q][r] for r in range(self.rank))   # receiver q lays senders out in ascending rank order
            off_in_mine = sum(M[self.rank][r] for r in range(q))
            self.plan.append((q, ids, ids.numel(), off_in_peer, off_in_mine))
        self.peers = [(q, ids, None, None) for q, ids, _, _ in peers]
        return self.peers

    def exchange_add(self, grids):
        half = (self.step & 1) * self.cap
        self.step += 1
        for q, ids, n, off_peer, _ in self.plan:       # pack + transfer in one kernel: stores land in rank q's HBM
            self._pack_ptr(grids, ids, self.ptrs[q] + 4 * (half + off_peer) * self.tile)
        self.hdl.barrier(channel=0)                    # device-side, on the current stream
        for q, ids, n, _, off_mine in self.plan:
            self._unpack_add(grids, ids, self.buf[(half + off_mine) * self.tile:(half + off_mine + n) * self.tile])


class HaloFused:
    """The exchange with nothing of its own on the stream but one barrier: the binned P2G's write-back reduce-adds every shared tile
    straight into the peers' receive buffers (TMA bulk reduce over NVLink, zpcb200_p2g_apic_fcr_binned_halo), the grid update adds
    what arrived and zeroes the slots (zpcb200_grid_update_halo).  The maps are built on the device from one fixed-size all_gather of
    block codes (zpcb200_halo_codes / zpcb200_halo_build): no host read anywhere on the re-bin path.  include/zpcb200.h: zpc_halo_view."""

    def __init__(self, group, device, capacity_blocks, seg=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > api.HALO_MAX_PEERS:
            raise ValueError("HaloFused handles up to %d ranks" % api.HALO_MAX_PEERS)
        self.cap = int(capacity_blocks)
        self.seg = int(seg) if seg else max(8192, self.cap // 2)   # a neighbour slab can share almost half of the active set (thin slabs + rings)
        self.tile = 7 * 64
        self.buf = symm_mem.empty(2 * self.world * self.seg * self.tile, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD if group is None else group)
        self.buf.zero_()
        self.peer = torch.full((self.cap * api.HALO_K,), -1, dtype=torch.int32, device=device)
        self.pos = torch.zeros(self.cap * api.HALO_K, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.codes = torch.empty(self.cap, dtype=torch.int64, device=device)
        self.all_codes = torch.empty(self.world * self.cap, dtype=torch.int64, device=device)
        self.step = 0
        self._ptrs = (C.c_void_p * api.HALO_MAX_PEERS)(*([int(p) for p in self.hdl.buffer_ptrs] + [None] * (api.HALO_MAX_PEERS - self.world)))
        self.hdl.barrier(channel=0)      # every buffer is zero before anybody sends

    def build_from_table(self, table):
        """collective, device-only: codes -> all_gather -> maps"""
        api.halo_codes(table, self.cap, self.codes)
        dist.all_gather_into_tensor(self.all_codes, self.codes, group=self.group)
        api.halo_build(self.all_codes, self.world, self.rank, self.cap, self.seg, self.peer, self.pos, self.status)

    def view(self):
        import ctypes as C
        return api.zpc_halo_view(self.peer.data_ptr(), self.pos.data_ptr(), self.world, self.rank, self.seg, self.step & 1,
                                 self.buf.data_ptr(), self._ptrs, self.status.data_ptr())

    def barrier(self):
        self.hdl.barrier(channel=0)      # device-side, on the current stream

    def shared_blocks(self):
        return int((self.peer.view(-1, api.HALO_K)[:, 0] >= 0).sum().item())


def _cuda_pack_ptr(grids, ids, dst_ptr):
    import ctypes as C
    api._check(api.lib().zpcb200_halo_pack(grids.view(), C.c_void_p(ids.data_ptr()), C.c_int(ids.numel()), C.c_int(0),
                                           C.c_int(grids.nch), C.c_void_p(dst_ptr), api._stream_ptr()), "halo_pack(p2p)")


def _cuda_pack(grids, ids, buf):
    import ctypes as C
    api._check(api.lib().zpcb200_halo_pack(grids.view(), C.c_void_p(ids.data_ptr()), C.c_int(ids.numel()), C.c_int(0),
                                           C.c_int(grids.nch), C.c_void_p(buf.data_ptr()), api._stream_ptr()), "halo_pack")


def _cuda_unpack_add(grids, ids, buf):
    import ctypes as C
    api._check(api.lib().zpcb200_halo_unpack_add(grids.view(), C.c_void_p(ids.data_ptr()), C.c_int(ids.numel()), C.c_int(0),
                                                 C.c_int(grids.nch), C.c_void_p(buf.data_ptr()), api._stream_ptr()),
               "halo_unpack_add")


class DistMpmSolver:
    def __init__(self, P_local, dx, volume, dt, gravity=-9.8, mode=1, rebin_every=8, group=None, device="cuda",
                 transport="auto", layout="binned", halo=None, **kw):
        """layout="binned": the fast path (substep).  layout="aos": the reference's particle layout and order, partition rebuilt
        every step — what substep_host (host buffers in / out) runs on.  halo: a ready HaloExchange (tests)."""
        self.local = MpmSolver(P_local, dx, volume, dt, gravity, mode, layout=layout, rebin_every=rebin_every,
                               device=device, partition="with_rebin" if layout == "binned" else "every_step", **kw)
        self.n = self.local.n
        self.table = self.local.table
        self.halo = halo
        fused_ok = layout == "binned" and isinstance(self.local.model, api.zpc_fixed_corotated)
        if halo is None and transport in ("auto", "fused") and fused_ok:
            try:   # send inside the P2G write-back, receive inside the grid update, maps built on the device
                self.halo = HaloFused(group, device, self.local.block_cap)
            except Exception as ex:
                if transport == "fused":
                    raise
                self.halo_fallback_reason = repr(ex)
        if self.halo is None and transport in ("auto", "p2p", "fused"):
            try:
                self.halo = HaloExchangeP2P(group, 7, device, _cuda_pack_ptr, _cuda_unpack_add)
            except Exception as ex:  # no symmetric memory on this system: same exchange over NCCL send/recv
                if transport == "p2p":
                    raise
                self.halo_fallback_reason = repr(ex)
        if self.halo is None:
            self.halo = HaloExchange(group, 7, device, _cuda_pack, _cuda_unpack_add)
        self.transport = "fused" if isinstance(self.halo, HaloFused) else "p2p" if isinstance(self.halo, HaloExchangeP2P) else "nccl"
        if isinstance(self.halo, HaloFused):
            self.local.extra_status.append((self.halo.status, "halo maps: a block shared with more than %d ranks, or a segment of the receive "
                                                              "buffer too small" % api.HALO_K))
        self.group = group
        self._cfl_work = None
        self._cfl_reduced = False
        self._rebuild_topology()

    @property
    def stage_events(self):
        return self.local.stage_events

    @stage_events.setter
    def stage_events(self, v):
        self.local.stage_events = v

    def stage_times_ms(self):
        return self.local.stage_times_ms()

    def _rebuild_topology(self):
        if isinstance(self.halo, HaloFused):
            self.halo.build_from_table(self.local.table)      # nothing read back to the host
            return
        nb = self.local.table.size()
        self.halo.build(self.local.table.active_keys[:nb])

    def substep_host(self, hin, hout):
        """Reference-facing call with HOST buffers on every rank (pinned torch tensors x, v, m, C, F of this rank's particles
        in; x, v, C, F out), layout="aos": upload -> partition -> exchange topology -> clean -> P2G -> halo exchange ->
        grid update -> G2P -> download.  Collective.  Returns the global max |v|^2."""
        L = self.local
        a = L.aos
        if self._cfl_work is not None:
            self._cfl_work.wait()
            self._cfl_work = None
        for k in ("x", "v", "m", "C", "F"):
            getattr(a, k).copy_(hin[k], non_blocking=True)
        api.partition_for_particles(api.vec3_port(a.x), L.n, L.dx, L.table)
        self._rebuild_topology()                       # the partition is new every step on this path
        api.clean_grid_blocks(L.grids, L.table)
        api.p2g_transfer(a, L.table, L.grids, L.dt, L.model)
        self.halo.exchange_add(L.grids)
        L.max_vel_sqr.zero_()
        L._grid_update()
        dist.all_reduce(L.max_vel_sqr, op=dist.ReduceOp.MAX, group=self.group)
        self._cfl_reduced = True
        api.g2p_transfer(a, L.table, L.grids, L.dt, model=L.model)
        for k in ("x", "v", "C", "F"):
            hout[k].copy_(getattr(a, k), non_blocking=True)
        return float(L.max_vel_sqr.item())             # D2H read = sync point

    def migrate(self, ownership):
        """Hands every particle to the rank that owns its current home block (BlockOwnership), then rebuilds the local solver
        on what arrived: unbin -> all_to_all of 100-byte records -> partition + bin.  Collective; call it at a re-bin
        boundary, every few hundred substeps — ownership only matters for load balance, never for correctness."""
        L = self.local
        if self._cfl_work is not None:
            self._cfl_work.wait()
            self._cfl_work = None
        dev = L.device
        aos = {k: L.bins.attr(k).clone() for k in ("x", "v", "m", "C", "F")}
        if L._side:                                    # logJp / J travel with their particle (one more float per record)
            aos[L._side] = getattr(L.bins, L._side).clone()
        dest = ownership.owner_of_positions(aos["x"], L.dx)
        new = migrate_particles(aos, dest, self.group)
        moved = int((dest != dist.get_rank(self.group)).sum().item())
        P = {k: v for k, v in new.items()}
        # same block capacity as before (the halo maps and receive segments are sized by it; the default, n / 256, is too small
        # for thin slabs with their partition ring)
        kw = dict(gravity=L.extf[1], mode=L.mode, layout="binned", rebin_every=L.rebin_every, device=dev, partition="with_rebin",
                  model=L.model, colliders=L.colliders, expected_blocks=L.block_cap)
        step_no = L.step_no
        L.flush_status()                               # deferred reads of the solver that retires
        self.local = MpmSolver(P, L.dx, L.model.volume, L.dt, **kw)
        self.local.step_no, self.local.status_mode, self.local.check_status = step_no, L.status_mode, L.check_status
        self.local.extra_status = L.extra_status       # the halo maps' status word stays registered
        self.n, self.table = self.local.n, self.local.table
        self._rebuild_topology()
        return moved

    def max_vel_sqr(self):
        """global max |v|^2 of the last substep (CFL input).  Collective: the all_reduce(max) of the per-rank scalar happens HERE, when
        somebody asks — one NCCL enqueue per substep that nobody reads cost more host time than the whole 8-GPU step could hide
        (the substep is ~1.5 ms at C4; round 2 measured a 0.3 ms gap per step on every rank from host-side enqueue work alone)."""
        if not self._cfl_reduced:
            dist.all_reduce(self.local.max_vel_sqr, op=dist.ReduceOp.MAX, group=self.group)
            self._cfl_reduced = True
        return self.local.max_vel_sqr

    # ---- CUDA graph of a whole re-bin cycle: at 8 GPUs a substep is ~1.5 ms of kernels and ~14 launches + collectives, i.e. the
    # host cannot enqueue it faster than the GPU runs it; one graph launch per 2 * rebin_every substeps removes the host from the loop
    def capture_cycle(self):
        """Captures the next 2 * rebin_every substeps — two re-bins (the ping-pong particle buffers return to their roles), two
        topology rebuilds (one fixed-size all_gather each), 2 * rebin_every device barriers — into one CUDA graph.  Needs the fused
        halo (its maps are built on the device: nothing in the cycle reads back to the host).  Collective."""
        L = self.local
        if not isinstance(self.halo, HaloFused) or L.rebin_every <= 0:
            raise ValueError("capture_cycle needs the fused halo transport and rebin_every > 0")
        k = 2 * L.rebin_every
        while L.step_no == 0 or L.step_no % k != 0 or (self.halo.step & 1):
            self.substep()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        step0, bins0, hstep0 = L.step_no, L.bins, self.halo.step
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            for _ in range(k):
                self.substep()
        assert L.bins is bins0 and (self.halo.step - hstep0) % 2 == 0
        L.step_no, self.halo.step, self._graph_len = step0, hstep0, k      # capturing executed nothing
        return k

    def replay_cycle(self):
        """one graph launch = 2 * rebin_every substeps; the status words are read once per replay (outside the graph)"""
        self._graph.replay()
        self.local.step_no += self._graph_len
        self._cfl_reduced = False
        if self.local.check_status:
            self.local.check_status_words()

    def substep(self):
        L = self.local
        self._cfl_reduced = False
        if L.prepare():
            self._rebuild_topology()
        L._mark("begin")
        api.clean_grid_blocks(L.grids, L.table)
        L._mark("clean")
        if isinstance(self.halo, HaloFused):
            hv = self.halo.view()
            api.p2g_transfer_halo(L.bins, L.table, L.grids, L.dt, L.model, hv)      # shared tiles go to the peers from the write-back
            L._mark("p2g")
            self.halo.barrier()
            L._mark("halo")
            L.max_vel_sqr.zero_()
            api.grid_update_halo(L.grids, L.table, L.dt, L.extf, L.mode, L.colliders, L.max_vel_sqr, hv)   # ... and are added here
            self.halo.step += 1
        else:
            api.p2g_transfer(L.bins, L.table, L.grids, L.dt, L.model)
            L._mark("p2g")
            self.halo.exchange_add(L.grids)
            L._mark("halo")
            L.max_vel_sqr.zero_()
            L._grid_update()                # the local solver's own update: colliders included (ComputeGridBlockVelocity + boundaries)
        L._mark("grid_update")             # the CFL scalar stays per-rank until max_vel_sqr() asks for it
        api.g2p_transfer(L.bins, L.table, L.grids, L.dt, model=L.model)   # the J variant for an equation of state
        L._mark("g2p")
        L.step_no += 1
