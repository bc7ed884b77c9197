#!/bin/bash
# Round 2, GPU call 1: pending tests with real tracebacks, default bench, sweep-5 A/B, the reference's own CUDA path.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2c1_smi.log 2>&1
timeout 1200 python -m pytest tests/test_zz_gpu_pending.py -m gpu -q --runxfail -rA --tb=short --timeout 600 > gpurun_out/r2c1_pending.log 2>&1
echo "pending rc=$?"; tail -15 gpurun_out/r2c1_pending.log
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2c1_bench.log 2>&1
cut -c1-600 gpurun_out/r2c1_bench.log | tail -2
timeout 300 python benchmarks/variants.py --combos 4:1,5:1,4:1,5:1 --tag r2_sweep5 > gpurun_out/r2c1_variants_sweep5.jsonl 2> gpurun_out/r2c1_variants_sweep5.err
tail -4 gpurun_out/r2c1_variants_sweep5.jsonl | cut -c1-300
timeout 400 python bench.py --impl reference-cuda --config C2 --steps 5 --warmup 2 > gpurun_out/r2c1_refcuda_c2.log 2>&1
cut -c1-800 gpurun_out/r2c1_refcuda_c2.log | tail -2
timeout 600 python bench.py --impl reference-cuda --config C3 --steps 3 --warmup 1 > gpurun_out/r2c1_refcuda_c3.log 2>&1
cut -c1-800 gpurun_out/r2c1_refcuda_c3.log | tail -2
for lg in 20 24 28; do timeout 200 python -m oracle.refcuda_runner prims-bench $lg 5; done > gpurun_out/r2c1_prims_vs_refcuda.jsonl 2> gpurun_out/r2c1_prims_vs_refcuda.err
tail -3 gpurun_out/r2c1_prims_vs_refcuda.jsonl | cut -c1-500
