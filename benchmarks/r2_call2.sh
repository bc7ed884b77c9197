#!/bin/bash
# Round 2, GPU call 2: parity of the plane sweep (tuning 6) + A/B timing against sweep 4.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x --timeout 600 -k "plane or not sweep" > gpurun_out/r2c2_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2c2_tests.log
timeout 400 python benchmarks/variants.py --combos 4:1,6:1,4:1,6:1 --tag r2_plane > gpurun_out/r2c2_variants.jsonl 2> gpurun_out/r2c2_variants.err
cut -c1-330 gpurun_out/r2c2_variants.jsonl; tail -3 gpurun_out/r2c2_variants.err
