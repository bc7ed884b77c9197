#!/bin/bash
# session 2, call 4: pipelined records + Jacobi early-out as the default sweep 4: parity (whole mpm + models files), A/B against builds without either
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mpm.py tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider > gpurun_out/s2c4_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s2c4_tests.log
for v in "" nopipe noearly base0; do
  if [ -n "$v" ]; then export ZPCB200_LIB=$PWD/zpc_b200/build/exp/$v.so; fi
  timeout 300 python benchmarks/variants.py --config C3 --steps 8 --warmup 3 --combos 4:1,4:1 --tag "${v:-default}" >> gpurun_out/s2c4_ab.log 2>> gpurun_out/s2c4_ab.err; echo "$v rc=$?"
done
cut -c1-300 gpurun_out/s2c4_ab.log
