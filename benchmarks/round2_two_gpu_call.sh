#!/bin/bash
# Second GPU call of the next round, two GPUs:
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash benchmarks/round2_two_gpu_call.sh'
# sharded substeps vs the single-GPU solver (with and without particle migration), then the 2-GPU bench line.
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tests/dist_check.py > gpurun_out/r2_dist_check.log 2>&1; echo "dist_check rc=$?"
ZPC_MIGRATE=1 timeout 300 $T tests/dist_check.py > gpurun_out/r2_dist_check_migrate.log 2>&1; echo "dist_check (migrate) rc=$?"
ZPC_E2E=1 timeout 300 $T tests/dist_check.py > gpurun_out/r2_dist_check_e2e.log 2>&1; echo "dist_check (host buffers) rc=$?"
timeout 600 $T bench.py --gpus 2 --steps 16 --warmup 4 > gpurun_out/r2_bench_2gpu.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/r2_dist_check.log gpurun_out/r2_dist_check_migrate.log; cut -c1-300 gpurun_out/r2_bench_2gpu.log | tail -2
