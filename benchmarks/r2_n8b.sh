#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-8}
for halo in fused p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 32 --warmup 8 --e2e-steps 0 --no-parity-check --halo $halo > gpurun_out/${TAG:-r2}_bench_n${N}_$halo.log 2> gpurun_out/${TAG:-r2}_bench_n${N}_$halo.err
echo "bench $halo rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG:-r2}_bench_n${N}_$halo.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "halo", d["fused_step"]["halo_ms"], "clean", d["fused_step"]["kernels"]["clean"]["ms"], "p2g", d["fused_step"]["kernels"]["p2g"]["ms"], "g2p", d["fused_step"]["kernels"]["g2p"]["ms"], "upd", d["fused_step"]["kernels"]["grid_update"]["ms"], "rebin", d["fused_step"]["rebin_ms_each"], "part", d["fused_step"]["partition_ms"])
PY
done
