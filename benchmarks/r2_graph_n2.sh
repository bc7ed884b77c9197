#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > gpurun_out/g_dist.log 2>&1; echo "dist_check rc=$?"; grep -E "dist_check" gpurun_out/g_dist.log | cut -c1-300
for g in auto off; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 32 --warmup 8 --e2e-steps 0 --graph $g > gpurun_out/g_$g.log 2> gpurun_out/g_$g.err
echo "graph=$g rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/g_$g.log").read().strip().splitlines()[-1])
f=d["fused_step"]
print("ms/step %.3f"%d["ms_per_step"], d["cuda_graph"], "parity", d["multi_gpu_parity"]["ok"] if d["multi_gpu_parity"] else None, "gap", f["gap_ms"], "launches", d["gpu_launches"])
PY
grep -E "Error|error" gpurun_out/g_$g.err | head -3
done
