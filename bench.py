#!/usr/bin/env python
"""bench.py — MPM particle<->sparse-grid substep throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C3|C2|C1] [--impl ours|reference]

One "step" = one explicit APIC substep over the whole synthetic particle cloud:
  partition build -> clear grid -> P2G -> grid update -> G2P   (+ the re-bin every `--rebin-every` substeps).
value   = particle-substeps / s, inputs resident in HBM (whole job, all ranks; max-over-ranks time).
e2e     = the same metric through the reference-facing host-buffer call (MpmSolver.substep_host_pipelined): pinned host
          arrays -> H2D -> substep on the reference's AoS layout -> D2H of the particle state, every step, the copies
          overlapped with the chunked P2G / G2P.  PCIe-bound at N = 1: 6.4 GB up + 6.1 GB down per step, and the download
          cannot start before the last upload has been scattered (the grid couples all particles).
roofline= the dominant kernel (binned P2G): algorithmic bytes (100 B/particle read + 28 B per
          active cell written at 8 ppc = 103.5 B/particle) / its CUDA-event time, against MEASURED_PEAKS.json.
cpu_baseline = the reference's own OpenMP path (oracle/_ref, built from /root/reference) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpm_particle_substeps_per_sec"
UNIT = "particle-substeps/s"
BYTES_PER_PARTICLE = dict(clean=28 / 8, p2g=100 + 28 / 8, grid_update=(28 + 12) / 8, g2p=48 + 96 + 12 / 8)  # sum 257.5
# EquationOfStateConfig (SURVEY §8(d) "cheap variant"): P2G reads x v m C J = 68 B, G2P reads x J = 16 B and writes x v C J = 64 B -> 161.5
BYTES_PER_PARTICLE_EOS = dict(clean=28 / 8, p2g=68 + 28 / 8, grid_update=(28 + 12) / 8, g2p=16 + 64 + 12 / 8)
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def run_reference_sample(G, steps, warmup, sample_s=50):
    """Times the reference (oracle/_ref = unmodified reference headers on omp_exec, all host threads; falls back to
    the single-thread C port when the reference library was not built) on an s^3-cell sub-cube of the workload."""
    import numpy as np  # noqa: F401
    from oracle.pyoracle import Oracle, Ref
    from zpc_b200 import synth
    cores = os.cpu_count() or 1
    if Ref.available():
        P = synth.elastic_cube(sample_s, G)
        n = P["x"].shape[0]
        r = Ref()
        h = r.mpm(n, P["dx"], cores, max(n // 8, 1))
        h.set_particles(P)

        def step():
            h.partition(); h.clean_grid(); h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
            h.grid_update(synth.DT, synth.GRAVITY, 1); h.g2p(synth.DT)
        kind = "reference"
        sample = "%d-particle sub-cube (%d^3 cells, 8 ppc, dx=1/%d) of the workload, omp_exec().threads(%d), %d timed substeps after %d warm-up" % (
            n, sample_s, G, cores, steps, warmup)
    else:
        sample_s = min(sample_s, 20)
        P = synth.elastic_cube(sample_s, G)
        n = P["x"].shape[0]
        o = Oracle()
        cores = 1

        def step():
            o.substep(P, P["dx"], synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"], synth.GRAVITY, 1)
        kind = "port"
        sample = "%d-particle sub-cube (%d^3 cells, 8 ppc, dx=1/%d), scalar C port, 1 thread" % (n, sample_s, G)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=n / dt, unit=UNIT, cores=cores, kind=kind, sample=sample, ms_per_step=dt * 1e3, n=n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default=os.environ.get("ZPC_BENCH_CONFIG", "C3"))
    ap.add_argument("--rebin-every", type=int, default=8)
    ap.add_argument("--partition", default="with_rebin", choices=["with_rebin", "every_step"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--model", default="fcr", choices=["fcr", "eos"], help="fcr = the headline workload (fixed-corotated); eos = the weakly "
                    "compressible fluid of SURVEY §8(d) on the same cloud (single GPU, no e2e leg)")
    ap.add_argument("--e2e-timeout", type=int, default=240, help="seconds after which the N>1 e2e leg is abandoned (the line is still printed)")
    ap.add_argument("--e2e-pipelined", type=int, default=8, help="chunks of MpmSolver.substep_host_pipelined: PCIe copies overlapped with the chunked "
                    "P2G / G2P (default; measured 232 ms against 256 ms for the plain call at C3 — both PCIe-bound: 12.5 GB per step over one link); "
                    "0 = the plain MpmSolver.substep_host")
    ap.add_argument("--halo", default="auto", choices=["auto", "fused", "p2p", "nccl"])
    ap.add_argument("--graph", default="auto", choices=["auto", "off"], help="N > 1: replay the substeps as CUDA graphs of 2 x rebin_every substeps "
                    "(DistMpmSolver.capture_cycle: kernels, re-bins, topology all_gathers and device barriers in one launch); the per-stage "
                    "times then come from a separate eager pass.  auto = fall back to eager launches if the capture is refused")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the N-GPU vs 1-GPU self-check (zpc_b200/selfcheck.py) that precedes the timing")
    ap.add_argument("--refcuda-steps", type=int, default=2, help="N = 1: timed substeps of the reference's own CUDA path on the same GPU "
                    "(oracle/_ref/libzpcref_cuda.so, a process of its own; informational `vs_reference_cuda` block; 0 = skip)")
    ap.add_argument("--prims-log2", type=int, default=26, help="N = 1: size of the C5 primitive line (reduce / exclusive_scan / radix_sort_pair, "
                    "i32 / u32 keys) added as `prims`; 0 = skip")
    ap.add_argument("--p2g-sweep", type=int, default=-1, choices=[-1, 3, 4, 5, 6], help="binned P2G sweep variant (zpcb200_set_tuning; 5 = packed fp32)")
    ap.add_argument("--g2p-staged", type=int, default=-1, choices=[-1, 0, 1], help="binned G2P particle staging variant")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from zpc_b200 import synth
    G, s = synth.CONFIGS[args.config]
    n_total = 8 * s ** 3
    workload = "%s: %d particles (%d^3 cells x 8 ppc), %d^3 sparse grid (dx=1/%d), %s APIC fp32" % (
        args.config, n_total, s, G, G, "fixed-corotated" if args.model == "fcr" else "equation-of-state (bulk 4e4, viscosity 0.01)")
    bytes_pp = BYTES_PER_PARTICLE if args.model == "fcr" else BYTES_PER_PARTICLE_EOS
    if args.model == "eos" and (int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.impl != "ours"):
        raise SystemExit("--model eos is a single-GPU line of this implementation")

    if args.impl == "reference-cuda":
        # informational: the reference's own CUDA functors on the same GPU (oracle/_ref/libzpcref_cuda.so, `make -C oracle refcuda`),
        # in a process of its own.  Not the contract's reference arm (that is the CPU path below).
        if rank != 0:
            return 0
        r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "bench", str(G), str(s), str(max(args.steps, 1)), str(max(args.warmup, 0))],
                           cwd=ROOT, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else json.dumps(
            dict(impl="reference-cuda", unavailable=(r.stderr.strip().splitlines() or ["failed"])[-1][:200]))
        print(line)
        return 0
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference_sample(G, max(args.steps, 1), max(args.warmup, 0))
        line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                    data="synthetic", impl="reference",
                    config=dict(workload=workload, sample=r["sample"]),
                    cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    from zpc_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hbm_peak, peak_src = peaks()
    api.set_tuning(args.p2g_sweep, args.g2p_staged)
    # ---- N > 1: the sharded path is checked against the single-GPU solver before anything is timed ----------------
    mg_parity = None
    if world > 1 and not args.no_parity_check:
        from zpc_b200.selfcheck import multi_gpu_parity
        try:
            plain = multi_gpu_parity(transport=args.halo)
            moved = multi_gpu_parity(transport=args.halo, migrate=True)
            graphed = multi_gpu_parity(transport=args.halo, graph=True) if (args.graph == "auto" and plain["transport"] == "fused") else None
            mg_parity = dict(ok=bool(plain["ok"] and moved["ok"] and (graphed is None or graphed["ok"])), substeps=plain,
                             substeps_with_migration=moved, substeps_as_cuda_graph=graphed,
                             what="N-GPU sharded substeps vs the single-GPU solver on the same cloud, 24^3 .. 96^3 cells by world size (particles cross the "
                                  "slab cuts), all particle attributes within 5e-5 after 6 substeps; zpc_b200/selfcheck.py")
        except Exception as ex:   # collective: raised on every rank alike
            mg_parity = dict(ok=False, error=repr(ex))

    # ---- build the (rank's shard of the) workload ----------------------------------------------------------
    if world == 1:
        from zpc_b200.solver import MpmSolver
        P = synth.elastic_cube(s, G)
        kw_model = {}
        if args.model == "eos":
            P = {k: v for k, v in P.items() if k != "F"}
            P["J"] = np.ones(n_total, np.float32)
            kw_model = dict(model=api.model_eos(P["volume"], 4.0e4, 7.15, 0.01))
            args.e2e_steps = 0
        sol = MpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, layout="binned", rebin_every=args.rebin_every,
                        partition=args.partition, **kw_model)
        n_local = sol.n
    else:
        from zpc_b200.dist_solver import DistMpmSolver
        P = synth.elastic_cube_slab(s, G, rank, world)
        sol = DistMpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, rebin_every=args.rebin_every,
                            transport=args.halo)
        n_local = sol.n
        # the device status words (stray particles, capacities, halo maps) come back through an asynchronous copy, looked at one
        # re-bin later and flushed after the timed region: at 1.4 ms per substep a host that stops at every re-bin cannot keep the
        # GPUs fed (MpmSolver.check_status_words)
        sol.local.status_mode = "deferred"
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sol.substep()
    barrier()
    graph_len, graph_note = 0, None
    if world > 1 and args.graph == "auto" and getattr(sol, "transport", None) == "fused":
        try:
            graph_len = sol.capture_cycle()
            sol.replay_cycle()                 # one untimed replay: graph upload
        except Exception as ex:                # collective code path: refused on every rank alike
            graph_len, graph_note = 0, "capture refused: %r" % (ex,)
        barrier()
    launches0 = api.kernel_launch_count()
    if not graph_len:
        sol.stage_events = []
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph_len:
        for _ in range(args.steps // graph_len):
            sol.replay_cycle()
        for _ in range(args.steps % graph_len):
            sol.substep()
    else:
        for _ in range(args.steps):
            sol.substep()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = api.kernel_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    stage_steps = args.steps
    if graph_len:
        # kernels inside a graph replay are not counted by the library's launch counter and carry no stage events: one eager cycle
        # (same kernels, same order) supplies both — its wall time is not the headline, `ms` above is
        barrier()                              # rank 0 just spent 0.15 s closing the clock sampler: start the pass together
        launches0 = api.kernel_launch_count()
        sol.stage_events = []
        for _ in range(graph_len):
            sol.substep()
        barrier()
        launches = launches + (api.kernel_launch_count() - launches0) * (args.steps // graph_len)
        stage_steps = graph_len
    stage = sol.stage_times_ms()
    sol.stage_events = None
    status_note = None
    if world > 1:
        try:
            sol.local.flush_status()           # every outstanding read of the status words, and one of their final state
        except RuntimeError as ex:
            status_note = repr(ex)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        cnt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(cnt)
        launches = int(cnt.item())
        bad = torch.tensor([1 if status_note else 0], device="cuda", dtype=torch.int64)
        dist.all_reduce(bad)
        status_words = dict(mode="deferred (asynchronous read-back, flushed after the timed region)", ranks_with_a_status_bit=int(bad.item()),
                            rank0=status_note)
    else:
        status_words = dict(mode="read at every re-bin (blocking)", ranks_with_a_status_bit=0, rank0=None)   # a set bit raises there
    per_rank = None
    if dist is not None:
        # every rank's own stage times: the step is as slow as the slowest rank, the others wait at the barrier / the CFL reduction
        names = ("clean", "p2g", "halo", "grid_update", "g2p", "partition", "rebin", "gap")
        mine = torch.tensor([stage.get(k, 0.0) / stage_steps for k in names], device="cuda", dtype=torch.float64)
        allr = torch.zeros(world * len(names), device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = {k: [round(v, 4) for v in allr.view(world, len(names))[:, i].tolist()] for i, k in enumerate(names)}
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    nblocks = sol.table.size()
    transport = getattr(sol, "transport", None)

    # ---- per-kernel roofline from the live CUDA-event stage times ---------------------------------------------
    per_step = {k: v / stage_steps for k, v in stage.items()}
    n_rebins = (sum(1 for i in range(args.warmup, args.warmup + args.steps) if i > 0 and args.rebin_every > 0 and i % args.rebin_every == 0)
                if not graph_len else graph_len // max(args.rebin_every, 1))
    fused_ms = sum(per_step.get(k, 0.0) for k in ("clean", "p2g", "grid_update", "g2p"))
    kern = {}
    for k, bpp in bytes_pp.items():
        if per_step.get(k):
            gbs = bpp * n_local / (per_step[k] * 1e-3) / 1e9
            kern[k] = dict(ms=per_step[k], algorithmic_gbps=gbs, frac=gbs / hbm_peak)
    # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this config (per launch)
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tr.get("config") == args.config and world == 1 and args.model == "fcr":
            traffic = tr["kernels"]["p2g_binned_kernel"]["dram_bytes"]
            traffic_src = "profiles/r02_traffic.json: " + tr.get("source", "")   # a capture of this command under ncu, not this run
    except Exception:
        pass
    dom = "p2g"
    roof = dict(bound="hbm", kernel="p2g_binned_kernel", achieved=kern.get(dom, {}).get("algorithmic_gbps"), peak=hbm_peak, unit="GB/s",
                frac=kern.get(dom, {}).get("frac"), traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                algorithmic_bytes_per_launch=bytes_pp[dom] * n_local)
    fused_bpp = sum(bytes_pp.values())
    fused_gbps = fused_bpp * n_local / (fused_ms * 1e-3) / 1e9 if fused_ms else None
    fused = dict(ms=fused_ms, bytes_per_particle=fused_bpp, achieved=fused_gbps, frac=(fused_gbps / hbm_peak) if fused_gbps else None,
                 kernels=kern, partition_ms=per_step.get("partition"), halo_ms=per_step.get("halo"), gap_ms=per_step.get("gap"), rebin_ms_each=(stage.get("rebin", 0.0) / n_rebins) if n_rebins else None,
                 rebins_in_timed_region=n_rebins)

    tuning_now = api.get_tuning()
    vs_refcuda, prims = None, None

    def make_line(e2e, cpu):
        return dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=workload, layout="block-binned AoSoA TileVector<f32,32>, re-bin every %d substeps" % args.rebin_every,
                                partition=("hash-grid partition rebuilt every substep (EnlargeSparsity{0,2})" if (args.partition == "every_step" and world == 1)
                                           else "hash-grid partition rebuilt with each re-bin, one extra ring (EnlargeSparsity{-1,3})"),
                                kernel_variants=tuning_now, active_blocks=nblocks, l2="inputs (%.1f GB particle state) exceed the 126 MB L2; no explicit flush" % (n_local * 100 / 1e9),
                                parallelism="1 GPU" if world == 1 else "x-slab shards over %d GPUs, halo exchange of shared grid blocks via %s" % (
                                    world, {"fused": "TMA bulk reduce-adds from the P2G write-back straight into the peers' symmetric-memory buffers over NVLink, "
                                                     "one device barrier, receive fused into the grid update",
                                            "p2p": "peer stores into symmetric memory over NVLink + device barrier"}.get(transport, "NCCL send/recv"))),
                    substeps_per_sec=1e3 / ms_per_step, roofline=roof, fused_step=fused, cpu_baseline=cpu, e2e=e2e, gpu_launches=launches,
                    clocks=clocks, cuda_graph=dict(substeps_per_graph=graph_len, note=graph_note, stage_times="from one eager cycle after the timed region") if world > 1 else None,
                    per_rank_stage_ms=per_rank, status_words=status_words, multi_gpu_parity=mg_parity, vs_reference_cuda=vs_refcuda, prims=prims)

    # the e2e leg at N > 1 is collective: if it has not finished after --e2e-timeout seconds (a rank stuck in a collective), rank 0
    # still prints the line — the device-resident numbers above are complete — and every rank leaves
    import threading
    finished = threading.Event()

    def _bail():
        if finished.is_set():
            return
        if rank == 0:
            print(json.dumps(make_line(dict(value=None, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0,
                                            note="e2e at N>1 did not finish within %d s" % args.e2e_timeout), None)), flush=True)
        os._exit(0)
    watchdog = None
    if world > 1 and args.e2e_steps > 0:
        watchdog = threading.Timer(args.e2e_timeout, _bail)
        watchdog.daemon = True
        watchdog.start()

    # ---- e2e: host buffers in, host buffers out, through the reference-facing call (N = 1) -----------------------
    e2e = None
    if world == 1 and args.e2e_steps > 0:
        del sol
        torch.cuda.empty_cache()
        from zpc_b200.solver import MpmSolver
        sol2 = MpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, layout="aos")
        hin = {k: torch.from_numpy(P[k]).pin_memory() for k in ("x", "v", "m", "C", "F")}
        hout = {k: torch.empty_like(hin[k]).pin_memory() for k in ("x", "v", "C", "F")}
        bi = sum(hin[k].numel() * 4 for k in hin)
        bo = sum(hout[k].numel() * 4 for k in hout) + 4
        host_call = (lambda: sol2.substep_host_pipelined(hin, hout, args.e2e_pipelined)) if args.e2e_pipelined > 0 else (lambda: sol2.substep_host(hin, hout))
        host_call()                            # warm-up (allocations, first-touch)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_call()
            torch.cuda.synchronize()
            for k in ("x", "v", "C", "F"):     # feed the result back like a host-side caller would
                hin[k], hout[k] = hout[k], hin[k]
        dt = (time.perf_counter() - t0) / args.e2e_steps
        e2e = dict(value=n_total / dt, unit=UNIT, h2d_bytes_per_step=bi, d2h_bytes_per_step=bo, ms_per_step=dt * 1e3,
                   steps=args.e2e_steps, path=("MpmSolver.substep_host_pipelined, %d chunks" % args.e2e_pipelined if args.e2e_pipelined > 0
                                                 else "MpmSolver.substep_host") + " (AoS drop-in kernels)")
        del sol2
    elif world > 1 and args.e2e_steps > 0:
        # every rank feeds its own shard from pinned host memory and reads it back: N PCIe links in parallel.  All ranks run the
        # same code on same-shaped data, so a failure is raised on every rank at the same point (no rank is left in a collective).
        try:
            from zpc_b200.dist_solver import DistMpmSolver, HaloFused
            # the AoS path exchanges with pack -> peer stores -> unpack-add; the fused halo (send inside the BINNED P2G's write-back) has
            # no such entry, so the host-buffer solver builds its own peer-to-peer exchange when the fast path ran fused
            halo_obj = None if isinstance(sol.halo, HaloFused) else sol.halo
            del sol
            torch.cuda.empty_cache()
            sol2 = DistMpmSolver(P, P["dx"], P["volume"], synth.DT, synth.GRAVITY, mode=1, layout="aos", halo=halo_obj,
                                 transport="auto" if args.halo in ("auto", "fused") else args.halo)
            hin = {k: torch.from_numpy(P[k]).pin_memory() for k in ("x", "v", "m", "C", "F")}
            hout = {k: torch.empty_like(hin[k]).pin_memory() for k in ("x", "v", "C", "F")}
            bi = sum(hin[k].numel() * 4 for k in hin)
            bo = sum(hout[k].numel() * 4 for k in hout) + 4
            sol2.substep_host(hin, hout)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                sol2.substep_host(hin, hout)
                torch.cuda.synchronize()
                for k in ("x", "v", "C", "F"):
                    hin[k], hout[k] = hout[k], hin[k]
            barrier()
            dt = (time.perf_counter() - t0) / args.e2e_steps
            tt = torch.tensor([dt, float(bi), float(bo)], device="cuda", dtype=torch.float64)
            tmax = tt.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt)
            e2e = dict(value=n_total / float(tmax[0].item()), unit=UNIT, h2d_bytes_per_step=int(tt[1].item()), d2h_bytes_per_step=int(tt[2].item()),
                       ms_per_step=float(tmax[0].item()) * 1e3, steps=args.e2e_steps,
                       path="DistMpmSolver.substep_host on every rank (AoS drop-in kernels, halo exchange, host buffers per rank)")
            transport = sol2.transport
            del sol2
        except Exception as ex:  # raised on every rank alike; the device-resident number above stands
            e2e = dict(value=None, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0, note="e2e at N>1 failed: %r" % (ex,))
    elif world > 1:
        e2e = dict(value=None, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0, note="--e2e-steps 0")

    if watchdog is not None:
        watchdog.cancel()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference_sample(G, 8, 2)   # ~10-30 core-seconds on the box's host cores: a bounded sample of the same workload
            cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"], ms_per_step=r["ms_per_step"])
        except Exception as ex:  # the checker is optional; never let it break the GPU number
            cpu = dict(value=None, unit=UNIT, cores=0, kind="unavailable", sample=str(ex))

    # ---- informational: the reference's own CUDA path on the same GPU (SURVEY §8(d) secondary baseline), per stage -------------
    if rank == 0 and world == 1 and args.refcuda_steps > 0 and args.model == "fcr":
        try:
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, "-m", "oracle.refcuda_runner", "bench", str(G), str(s), str(args.refcuda_steps), "1"],
                               cwd=ROOT, capture_output=True, text=True, timeout=600)
            if r.returncode == 0 and r.stdout.strip():
                ref = json.loads(r.stdout.strip().splitlines()[-1][r.stdout.strip().splitlines()[-1].index("{"):])
                ours_stage = dict(partition=(per_step.get("partition") or 0.0), clean=per_step.get("clean"), p2g=per_step.get("p2g"),
                                  grid_update=per_step.get("grid_update"), g2p=per_step.get("g2p"))
                vs_refcuda = dict(reference_ms_per_step=ref["ms_per_step"], reference_stage_ms=ref["stage_ms"], ours_ms_per_step=ms_per_step,
                                  ours_stage_ms=ours_stage,
                                  speedup=dict(step=ref["ms_per_step"] / ms_per_step,
                                               **{k: ref["stage_ms"][k] / ours_stage[k] for k in ("clean", "p2g", "grid_update", "g2p") if ours_stage.get(k)}),
                                  note="the reference's generic functors on cuda_exec() (range_launch + CUB + HashTable lock-CAS insert), unmodified, "
                                       "compiled for sm_100 (oracle/_ref/libzpcref_cuda.so), same cloud, particles resident; it rebuilds its partition "
                                       "every substep (ours: with the re-bin, amortised in ours_ms_per_step)")
            else:
                vs_refcuda = dict(unavailable=(r.stderr.strip().splitlines() or ["failed"])[-1][:200])
        except Exception as ex:
            vs_refcuda = dict(unavailable=repr(ex)[:200])
    # ---- C5 in one line: the primitives at 2^k keys, algorithmic GB/s against the same peak ------------------------------------
    if rank == 0 and world == 1 and args.prims_log2 > 0:
        try:
            from benchmarks.prims_sweep import time_prims
            prims = time_prims(args.prims_log2, hbm_peak)
        except Exception as ex:
            prims = dict(unavailable=repr(ex)[:200])

    finished.set()
    if rank == 0:
        print(json.dumps(make_line(e2e, cpu)), flush=True)
    if dist is not None:
        # the line is out.  Leave without tearing NCCL / symmetric memory / captured graphs down object by object: a process group
        # destroyed while a CUDA graph still references its communicator has been seen to hang (round 2, N = 2), and nothing is
        # left to flush — every rank waits for the others, then exits
        try:
            torch.cuda.synchronize()
            dist.barrier()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
