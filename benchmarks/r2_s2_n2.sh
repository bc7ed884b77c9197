#!/bin/bash
# session 2, 2 GPUs: multi-GPU parity on the final kernels (plain, migration, graph replay), the 2-GPU pytest, the driver-style bench line; wide block codes
set -u
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x -p no:cacheprovider -k "wide_block_codes or multi_gpu" > gpurun_out/s2n2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s2n2_tests.log
for mode in plain migrate graph; do
  unset ZPC_MIGRATE ZPC_GRAPH
  if [ $mode = migrate ]; then export ZPC_MIGRATE=1; fi
  if [ $mode = graph ]; then export ZPC_GRAPH=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > gpurun_out/s2n2_dist_$mode.log 2>&1
  echo "dist_check $mode rc=$?"; grep -E "dist_check|Error|error" gpurun_out/s2n2_dist_$mode.log | tail -2 | cut -c1-400
done
unset ZPC_MIGRATE ZPC_GRAPH
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 32 --warmup 8 > gpurun_out/s2n2_bench.log 2> gpurun_out/s2n2_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/s2n2_bench.log").read().strip().splitlines()[-1])
f=d["fused_step"]
print("ms/step %.3f"%d["ms_per_step"], "value %.3g"%d["value"], d["cuda_graph"], "parity", d["multi_gpu_parity"], "halo", f["halo_ms"], "gap", f["gap_ms"], "e2e", d["e2e"])
PY
tail -2 gpurun_out/s2n2_bench.err
