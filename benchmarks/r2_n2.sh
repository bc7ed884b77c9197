#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-2}
for mode in plain migrate; do
  if [ $mode = migrate ]; then export ZPC_MIGRATE=1; else unset ZPC_MIGRATE; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > gpurun_out/${TAG:-r2}_dist_$mode.log 2>&1
  echo "dist_check $mode rc=$?"; grep -E "dist_check|Error|error" gpurun_out/${TAG:-r2}_dist_$mode.log | tail -3 | cut -c1-700
done
unset ZPC_MIGRATE
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 32 --warmup 8 --e2e-steps 0 --no-parity-check ${BENCH_ARGS:-} > gpurun_out/${TAG:-r2}_bench_n$N.log 2> gpurun_out/${TAG:-r2}_bench_n$N.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG:-r2}_bench_n$N.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "halo", d["fused_step"]["halo_ms"], "p2g", d["fused_step"]["kernels"]["p2g"]["ms"], "g2p", d["fused_step"]["kernels"]["g2p"]["ms"], "upd", d["fused_step"]["kernels"]["grid_update"]["ms"], "rebin", d["fused_step"]["rebin_ms_each"], "part", d["fused_step"]["partition_ms"])
print("parity", d["multi_gpu_parity"], d["config"]["parallelism"][:90])
PY
tail -3 gpurun_out/${TAG:-r2}_bench_n$N.err
