#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_prims.py -m gpu -q -x --timeout 900 -k "billion or block_key" > gpurun_out/r2c10_prims_tests.log 2>&1; echo "prims tests rc=$?"; tail -3 gpurun_out/r2c10_prims_tests.log
timeout 900 python bench.py --steps 16 --warmup 4 > gpurun_out/r2c10_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2c10_bench.log | cut -c1-1500
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --refcuda-steps 0 --prims-log2 0 --e2e-pipelined 8 > gpurun_out/r2c10_bench_e2e_pipe.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/r2c10_bench_e2e_pipe.log').read().strip().splitlines()[-1]);print('e2e pipelined',d['e2e'])"
for lg in 20 24 28; do timeout 200 python -m oracle.refcuda_runner prims-bench $lg 5; done > gpurun_out/r2c10_prims_vs_refcuda.jsonl 2> gpurun_out/r2c10_prims_vs_refcuda.err
grep -o '{"log2n".*' gpurun_out/r2c10_prims_vs_refcuda.jsonl | cut -c1-600
timeout 900 python benchmarks/prims_sweep.py --min-log2 20 --max-log2 30 > gpurun_out/r2c10_prims_sweep.jsonl 2> gpurun_out/r2c10_prims_sweep.err; tail -2 gpurun_out/r2c10_prims_sweep.jsonl | cut -c1-700
