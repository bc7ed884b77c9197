"""Composed explicit APIC substep on one GPU, through the public functor API.

    partition_for_particles -> CleanGridBlocks -> P2GTransfer -> ComputeGridBlockVelocity -> G2PTransfer

`layout="aos"` keeps the reference's Particles layout and order (drop-in path); `layout="binned"` keeps
particles in block-binned AoSoA TileVectors and re-bins every `rebin_every` substeps.

`partition="every_step"` rebuilds the hash-grid partition each substep exactly like the reference's composed step
(EnlargeSparsity{0,2}).  `partition="with_rebin"` (binned layout only) rebuilds it only together with the re-bin,
enlarged by one more ring (EnlargeSparsity{-1,3}): a particle that has drifted by less than a block since then
still finds every block of its stencil, the extra blocks stay empty (mass 0) and results are unchanged.

`model=` selects the constitutive model (api.model_fcr | model_vonmises | model_eos | model_drucker_prager | model_nacc;
default fixed-corotated from E, nu): the binned layout carries the 25 channels m, x, v, C, F and, for Drucker-Prager / NACC
(logJp) and the equation of state (J), a side array in bin order that is permuted with every re-bin.  `colliders=` (up to
four api.plane_collider / sphere_collider / cuboid_collider) are applied inside the grid-update pass
(ComputeGridBlockVelocity + ApplyBoundaryConditionOnGridBlocks fused, GridOp.hpp:71-164).
"""
import torch

from . import api


def default_expected_blocks(n):
    return max(n // 256, 1024)


class MpmSolver:
    def __init__(self, P, dx, volume, dt, gravity=-9.8, mode=1, layout="binned", expected_blocks=None,
                 rebin_every=8, E=5.0e4, nu=0.4, device="cuda", shuffle_free=True, partition="every_step", model=None,
                 colliders=()):
        if partition not in ("every_step", "with_rebin") or (partition == "with_rebin" and layout != "binned"):
            raise ValueError(partition)
        self.partition_mode = partition
        self.enlarge = (0, 2) if partition == "every_step" else (-1, 3)
        self.device = device
        self.dx, self.dt, self.mode = float(dx), float(dt), int(mode)
        self.extf = (0.0, float(gravity), 0.0)
        self.model = model if model is not None else api.model_fcr(volume, E, nu)
        # per-particle scalar the model carries besides F: logJp (plastic models), J (equation of state)
        self._side = ("logJp" if isinstance(self.model, (api.zpc_drucker_prager, api.zpc_nacc))
                      else "J" if isinstance(self.model, api.zpc_equation_of_state) else None)
        self.colliders = list(colliders)
        if len(self.colliders) > api.MAX_COLLIDERS:
            raise ValueError("at most %d colliders per grid pass" % api.MAX_COLLIDERS)
        self.layout = layout
        self.n = int(P["x"].shape[0])
        eb = expected_blocks or default_expected_blocks(self.n)
        self.table = api.HashTable(eb, device)
        self.block_cap = self.table.table_size // 16
        self.grids = api.Grids(dx, self.block_cap, 7, device)
        self.max_vel_sqr = torch.zeros(1, dtype=torch.float32, device=device)
        self.rebin_every = int(rebin_every)
        self.check_status = True      # read the device status words at every re-bin (one small D2H; off inside graph capture)
        self.status_mode = "sync"     # | "deferred": asynchronous read-back, looked at one re-bin later (check_status_words)
        self._status_pending = []
        self.extra_status = []        # [(device int tensor, message)] read together with them
        self.step_no = 0
        self.stage_events = None
        self.aos = api.Particles(P, device)
        if layout == "binned":
            self.bins = api.ParticleBins(self.n, self.block_cap, device)
            self.bins_alt = api.ParticleBins(self.n, self.block_cap, device)
            self.order = torch.arange(self.n, dtype=torch.int32, device=device)      # overwritten by bin_particles
            api.partition_for_particles(api.vec3_port(self.aos.x), self.n, self.dx, self.table, enlarge=self.enlarge)
            api.bin_particles(self.aos, self.table, self.dx, self.bins, self.order)
            if self._side:
                src = getattr(self.aos, self._side)
                if src is None:
                    raise ValueError("this model needs the per-particle %s attribute (P2G.hpp:67,93)" % self._side)
                # two persistent side buffers that ping-pong with bins / bins_alt (a captured graph ends on the buffer it
                # started from), permuted by the library's own gather
                setattr(self.bins, self._side, torch.empty_like(src))
                setattr(self.bins_alt, self._side, torch.empty_like(src))
                api.gather_f32(src, self.order, getattr(self.bins, self._side))
                self._rebin_order = torch.arange(self.n, dtype=torch.int32, device=device)
            self.bins.check_status("bin_particles")
            self.aos = None if shuffle_free else self.aos
        elif layout != "aos":
            raise ValueError(layout)

    # positions as an iterator port for the partition build
    def _x_port(self):
        if self.layout == "binned":
            return self.bins.pars.port(api.PB_X)
        return api.vec3_port(self.aos.x)

    def _pars(self):
        return self.bins if self.layout == "binned" else self.aos

    # ---- optional per-stage CUDA-event timing (bench.py) ----
    def _mark(self, name):
        if self.stage_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.stage_events.append((name, e))

    def stage_times_ms(self):
        """sum of elapsed ms per stage over everything recorded since stage_events was reset"""
        out = {}
        ev = self.stage_events or []
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            # "begin" closes a gap nobody claimed (between two marked sections: host-side work the GPU waited for, collectives on
            # other streams, the topology rebuild of the multi-GPU solver): reported as "gap"
            k = n1 if n1 != "begin" else "gap"
            out[k] = out.get(k, 0.0) + e0.elapsed_time(e1)
        return out

    def partition(self, stream=None):
        self._mark("begin")
        api.partition_for_particles(self._x_port(), self.n, self.dx, self.table, stream, enlarge=self.enlarge)
        self._mark("partition")

    def transfer(self, stream=None):
        """the fused P2G + grid + G2P part of the substep (the roofline-quoted part)"""
        self._mark("begin")
        api.clean_grid_blocks(self.grids, self.table, stream)
        self._mark("clean")
        api.p2g_transfer(self._pars(), self.table, self.grids, self.dt, self.model, stream)
        self._mark("p2g")
        self.max_vel_sqr.zero_()
        self._grid_update(stream)
        self._mark("grid_update")
        api.g2p_transfer(self._pars(), self.table, self.grids, self.dt, stream, model=self.model)
        self._mark("g2p")

    def _grid_update(self, stream=None):
        if self.colliders:
            api.compute_grid_block_velocity_with_boundaries(self.grids, self.table, self.dt, self.extf, self.mode, self.colliders,
                                                            self.max_vel_sqr, stream)
        else:
            api.compute_grid_block_velocity(self.grids, self.table, self.dt, self.extf, self.mode, self.max_vel_sqr, stream)

    def rebin(self, stream=None):
        self.partition(stream)
        self._mark("begin")
        if self._side:   # the side array (logJp / J) follows the permutation of the re-bin
            api.rebin_particles(self.bins, self.table, self.dx, self.bins_alt, stream, order_out=self._rebin_order)
            api.gather_f32(getattr(self.bins, self._side), self._rebin_order, getattr(self.bins_alt, self._side), stream)
        else:
            api.rebin_particles(self.bins, self.table, self.dx, self.bins_alt, stream)
        self.bins, self.bins_alt = self.bins_alt, self.bins
        self._mark("rebin")
        if self.check_status and not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
            self.check_status_words()

    def _status_words(self):
        # the bins that retired at the last re-bin (a stencil block missing from the partition = a particle out-ran the extra ring),
        # the current bins (capacity; strays since the re-bin), the table's overflow flag, whatever the owner registered (multi-GPU:
        # the halo maps)
        return [self.bins_alt.status, self.bins.status, self.table.overflow] + [t for t, _ in self.extra_status]

    def _raise_on_status(self, vals, owners):
        """vals: the words as host ints; owners: the objects they were read from (the bins may have swapped roles since)"""
        if not any(vals):
            return
        retired, current, table = owners[:3]
        if vals[0]:
            retired.status.zero_()
            raise RuntimeError("%s: %s" % ("substeps up to the last re-bin", api.bins_status_text(vals[0])))
        if vals[1]:
            current.status.zero_()
            raise RuntimeError("%s: %s" % ("re-bin / substeps since the last re-bin", api.bins_status_text(vals[1])))
        if vals[2]:
            raise RuntimeError("hash-grid partition overflow: raise expected_blocks")
        for v, (t, what) in zip(vals[3:], owners[3:]):
            if v:
                t.zero_()
                raise RuntimeError("%s (status %d)" % (what, v))

    def check_status_words(self):
        """Reads every status word of the path and raises on the first that is set.  Called at every re-bin and after every graph
        replay (the replayed re-bins cannot read back to the host).

        status_mode "sync" (default): ONE blocking D2H read, the error surfaces at the re-bin that follows the fault.
        status_mode "deferred": the words are copied to pinned host memory asynchronously and looked at once the copy has landed —
        at a later re-bin / replay or in flush_status() — so the host never waits for the GPU on this path (at 8 GPUs a substep is
        1.4 ms: a host that stops at every re-bin cannot enqueue fast enough).  particles_host() flushes."""
        words = self._status_words()
        owners = [self.bins_alt, self.bins, self.table] + list(self.extra_status)
        dev = torch.cat([w.reshape(1).to(torch.int32) for w in words])
        if self.status_mode != "deferred":
            self._raise_on_status(dev.tolist(), owners)
            return
        if dev.is_cuda:
            host = torch.empty(dev.numel(), dtype=torch.int32, pin_memory=True)
            host.copy_(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:                                             # host-logic tests: tensors on the CPU, nothing to wait for
            host, ev = dev.clone(), None
        self._status_pending.append((ev, host, owners))
        self.poll_status(block=False)

    def poll_status(self, block=False):
        """looks at the deferred reads that have landed (all of them when block=True); raises like check_status_words"""
        while self._status_pending:
            ev, host, owners = self._status_pending[0]
            if ev is not None:
                if block:
                    ev.synchronize()
                elif not ev.query():
                    return
            self._status_pending.pop(0)
            self._raise_on_status(host.tolist(), owners)

    def flush_status(self):
        """deferred mode: one more read of the words as they are now, then wait for every outstanding read"""
        if self.check_status and self.status_mode == "deferred" and self.layout == "binned":
            self.check_status_words()
        self.poll_status(block=True)

    def rebin_due(self):
        return self.layout == "binned" and self.step_no > 0 and self.rebin_every > 0 and self.step_no % self.rebin_every == 0

    def prepare(self, stream=None):
        """partition (+ re-bin when due) for the coming substep; returns True when the partition was rebuilt"""
        if self.rebin_due():
            self.rebin(stream)  # leaves a partition built from the current positions
            return True
        if self.partition_mode == "every_step":
            self.partition(stream)
            return True
        return False

    def substep(self, stream=None):
        self.prepare(stream)
        self.transfer(stream)
        self.step_no += 1

    # ---- CUDA graph of a whole re-bin cycle (launch-bound sizes: C1 / C2) ----
    def capture_cycle(self):
        """Captures the next 2 * rebin_every substeps (two re-bins, so that the ping-pong particle buffers end up in their
        original roles) into one CUDA graph: ~14 launches per substep become one graph launch per cycle.  Runs eagerly up
        to a cycle boundary first, so every scratch buffer already has its final size.  Binned layout, partition rebuilt
        with the re-bin (nothing in that sequence reads back to the host)."""
        if self.layout != "binned" or self.rebin_every <= 0 or self.partition_mode != "with_rebin":
            raise ValueError("capture_cycle needs layout='binned', rebin_every > 0 and partition='with_rebin'")
        k = 2 * self.rebin_every
        while self.step_no == 0 or self.step_no % k != 0:
            self.substep()
        torch.cuda.synchronize()
        step0, bins0 = self.step_no, self.bins
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            for _ in range(k):
                self.substep()
        assert self.bins is bins0                      # two swaps: the graph starts and ends on the same buffer
        self.step_no, self._graph_len = step0, k       # capturing executed nothing
        return k

    def replay_cycle(self):
        """one graph launch = 2 * rebin_every substeps; the status words are read once per replay (one small D2H after it)"""
        self._graph.replay()
        self.step_no += self._graph_len
        if self.check_status:
            self.check_status_words()

    def substep_host(self, hin, hout, stream=None):
        """Reference-facing call with HOST buffers (pinned torch tensors x,v,m,C,F in; x,v,C,F out), any particle
        order: upload -> partition -> clean -> P2G -> grid update -> G2P on the reference's AoS layout -> download.
        Returns max |v|^2 (the CFL scalar the reference reads back every substep, simulation/mpm/Simulator.hpp:19-26)."""
        a = self.aos
        for k in ("x", "v", "m", "C", "F"):
            getattr(a, k).copy_(hin[k], non_blocking=True)
        api.partition_for_particles(api.vec3_port(a.x), self.n, self.dx, self.table, stream)
        api.clean_grid_blocks(self.grids, self.table, stream)
        api.p2g_transfer(a, self.table, self.grids, self.dt, self.model, stream)
        self.max_vel_sqr.zero_()
        self._grid_update(stream)
        api.g2p_transfer(a, self.table, self.grids, self.dt, stream, model=self.model)
        for k in ("x", "v", "C", "F"):
            hout[k].copy_(getattr(a, k), non_blocking=True)
        return float(self.max_vel_sqr.item())   # D2H read = sync point

    def substep_host_pipelined(self, hin, hout, chunks=8):
        """substep_host with the PCIe transfers overlapped with the kernels (same results: the AoS kernels are per-particle).
        Positions go up first (the partition needs all of them); v, m, C, F follow in `chunks` pieces on a copy stream while
        the compute stream scatters each piece as it lands; after the grid update every piece is gathered and its x, v, C, F
        start downloading on a second copy stream while the next piece is gathered.  The floor is one PCIe direction at a
        time (the download depends on the whole upload through the grid)."""
        a, n = self.aos, self.n
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(), torch.cuda.Stream()
        s_in, s_out = self._s_in, self._s_out
        bounds = [n * c // chunks for c in range(chunks + 1)]
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        ev_in = []
        with torch.cuda.stream(s_in):
            a.x.copy_(hin["x"], non_blocking=True)
            ev_x = torch.cuda.Event()
            ev_x.record(s_in)
            for c in range(chunks):
                lo, hi = bounds[c], bounds[c + 1]
                for k in ("v", "m", "C", "F"):
                    getattr(a, k)[lo:hi].copy_(hin[k][lo:hi], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_in)
                ev_in.append(e)
        cur.wait_event(ev_x)
        api.partition_for_particles(api.vec3_port(a.x), n, self.dx, self.table)
        api.clean_grid_blocks(self.grids, self.table)
        for c in range(chunks):
            cur.wait_event(ev_in[c])
            if bounds[c + 1] > bounds[c]:
                api.p2g_transfer(a.range(bounds[c], bounds[c + 1]), self.table, self.grids, self.dt, self.model)
        self.max_vel_sqr.zero_()
        self._grid_update()
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            if hi > lo:
                api.g2p_transfer(a.range(lo, hi), self.table, self.grids, self.dt, model=self.model)
            e = torch.cuda.Event()
            e.record(cur)
            s_out.wait_event(e)
            with torch.cuda.stream(s_out):
                for k in ("x", "v", "C", "F"):
                    hout[k][lo:hi].copy_(getattr(a, k)[lo:hi], non_blocking=True)
        cur.wait_stream(s_out)
        return float(self.max_vel_sqr.item())   # D2H read on the compute stream = sync point (after the downloads)

    def particles_host(self):
        """AoS dict on the host, in the solver's CURRENT particle order."""
        if self.layout == "binned":
            self.flush_status()
            out = {k: self.bins.attr(k).cpu().numpy() for k in ("x", "v", "m", "C", "F")}
            if self._side:
                out[self._side] = getattr(self.bins, self._side).cpu().numpy()
            return out
        return self.aos.to_host()
