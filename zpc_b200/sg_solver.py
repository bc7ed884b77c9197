"""The composed MPM substep on SparseGrid<3,f32,8> + bht (geometry/SparseGrid.hpp) with the block-binned fast path:
partition (with each re-bin) -> clean -> P2G -> grid update -> G2P on particles kept in octant bins.  The SparseGrid
twin of zpc_b200/solver.py's MpmSolver(layout="binned", partition="with_rebin"); torch supplies memory and streams only."""
import torch

from . import api


class SgMpmSolver:
    def __init__(self, P, dx, volume, dt, gravity=-9.8, mode=1, rebin_every=8, E=5.0e4, nu=0.4, expected_blocks=None, device="cuda",
                 model=None):
        self.model = model if model is not None else api.model_fcr(volume, E, nu)
        # per-particle scalar the model carries besides F: logJp (Drucker-Prager, NACC), J (equation of state) — kept next to the bins
        # in bin order, two persistent buffers that ping-pong with bins / bins_alt (like MpmSolver)
        self._side = ("logJp" if isinstance(self.model, (api.zpc_drucker_prager, api.zpc_nacc))
                      else "J" if isinstance(self.model, api.zpc_equation_of_state) else None)
        self.dx, self.dt, self.mode = float(dx), float(dt), int(mode)
        self.extf = (0.0, float(gravity), 0.0)
        self.n = int(P["x"].shape[0])
        self.rebin_every = int(rebin_every)
        # side-8 blocks hold 512 cells = 4 096 particles at 8 per cell; the partition adds one block in +x, +y, +z (EnlargeSparsity{0,2})
        nb = int(expected_blocks or max(self.n // 1024, 512))
        self.sg = api.SparseGrid(7, nb, device)
        self.sg.scale(self.dx)
        self.max_vel_sqr = torch.zeros(1, dtype=torch.float32, device=device)
        self.step_no = 0
        self.stage_events = None
        aos = api.Particles(P, device)
        api.sg_partition_for_particles(api.vec3_port(aos.x), self.n, self.sg)
        cap = 8 * nb + 64                       # bins are octants: at most 8 per block
        self.bins, self.bins_alt = api.ParticleBins(self.n, cap, device), api.ParticleBins(self.n, cap, device)
        self.order = torch.empty(self.n, dtype=torch.int32, device=device)
        api.sg_bin_particles(aos, self.sg, self.bins, self.order)
        if self._side:
            src = getattr(aos, self._side)
            if src is None:
                raise ValueError("this model needs the per-particle %s attribute (P2G.hpp:67,93)" % self._side)
            setattr(self.bins, self._side, torch.empty_like(src))
            setattr(self.bins_alt, self._side, torch.empty_like(src))
            api.gather_f32(src, self.order, getattr(self.bins, self._side))
            self._rebin_order = torch.empty(self.n, dtype=torch.int32, device=device)
        self._check("sg_bin_particles")

    def _check(self, what):
        if int(self.sg.table.overflow.item()):
            raise RuntimeError("%s: SparseGrid partition overflow (raise expected_blocks)" % what)
        self.bins.check_status(what)

    def _mark(self, name):
        if self.stage_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.stage_events.append((name, e))

    def stage_times_ms(self):
        out = {}
        ev = self.stage_events or []
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            k = n1 if n1 != "begin" else "gap"
            out[k] = out.get(k, 0.0) + e0.elapsed_time(e1)
        return out

    def rebin(self):
        self._mark("begin")
        api.sg_partition_for_particles(self.bins.pars.port(api.PB_X), self.n, self.sg)
        self._mark("partition")
        if self._side:   # the side array follows the permutation of the re-bin
            api.sg_rebin_particles(self.bins, self.sg, self.bins_alt, order_out=self._rebin_order)
            api.gather_f32(getattr(self.bins, self._side), self._rebin_order, getattr(self.bins_alt, self._side))
        else:
            api.sg_rebin_particles(self.bins, self.sg, self.bins_alt)
        self.bins, self.bins_alt = self.bins_alt, self.bins
        self._mark("rebin")
        if int(self.bins_alt.status.item()):
            self.bins_alt.check_status("substeps since the last re-bin")
        self._check("sg_rebin_particles")

    def substep(self):
        if self.step_no > 0 and self.rebin_every > 0 and self.step_no % self.rebin_every == 0:
            self.rebin()
        self._mark("begin")
        api.sg_clean(self.sg)
        self._mark("clean")
        api.sg_p2g_transfer(self.bins, self.sg, self.dt, self.model)
        self._mark("p2g")
        self.max_vel_sqr.zero_()
        api.sg_compute_grid_velocity(self.sg, self.dt, self.extf, self.mode, self.max_vel_sqr)
        self._mark("grid_update")
        api.sg_g2p_transfer(self.bins, self.sg, self.dt, model=self.model)    # the J variant for an equation of state
        self._mark("g2p")
        self.step_no += 1

    def particles_host(self):
        out = {k: self.bins.attr(k).cpu().numpy() for k in ("x", "v", "m", "C", "F")}
        if self._side:
            out[self._side] = getattr(self.bins, self._side).cpu().numpy()
        return out
