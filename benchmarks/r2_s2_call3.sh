#!/bin/bash
# session 2, call 3: fixed-point arena (sweep 9), G2P without IEEE-division slow paths / local-memory stores: parity + A/B at C3
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x -p no:cacheprovider > gpurun_out/s2c3_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s2c3_tests.log
timeout 600 python benchmarks/variants.py --config C3 --steps 8 --warmup 3 --combos 4:1,8:1,9:1,4:1,8:1,9:1 > gpurun_out/s2c3_ab.log 2> gpurun_out/s2c3_ab.err; echo "ab rc=$?"; cut -c1-330 gpurun_out/s2c3_ab.log
