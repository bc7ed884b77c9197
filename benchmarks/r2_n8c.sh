#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-8}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 32 --warmup 8 --e2e-steps 0 --no-parity-check ${BENCH_ARGS:-} > gpurun_out/${TAG:-r2}_n$N.log 2> gpurun_out/${TAG:-r2}_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG:-r2}_n$N.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"])
for k,v in d["per_rank_stage_ms"].items(): print(k, v)
PY
