#!/bin/bash
# session 2, 8 GPUs: the driver-style scaling line on the final kernels
set -u
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 32 --warmup 8 --e2e-steps 0 > gpurun_out/s2n8_bench.log 2> gpurun_out/s2n8_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/s2n8_bench.log").read().strip().splitlines()[-1])
f=d["fused_step"]
print("ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], d["cuda_graph"], "parity", (d["multi_gpu_parity"] or {}).get("ok"), "halo", f["halo_ms"], "gap", f["gap_ms"], {k:round(v["ms"],3) for k,v in f["kernels"].items()})
PY
tail -2 gpurun_out/s2n8_bench.err
