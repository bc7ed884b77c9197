#!/bin/bash
# session 2, call 2: flush microbenchmark (float CAS vs int atomics vs RMW, tile vs padded layout); 512-thread CTAs / 512-particle chunks
set -u
mkdir -p gpurun_out
(cd benchmarks/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/flush flush.cu && /tmp/flush) > gpurun_out/s2c2_flush.log 2>&1; cat gpurun_out/s2c2_flush.log
(cd benchmarks/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 ffma2.cu && /tmp/ffma2) > gpurun_out/s2c2_ffma2.log 2>&1; cat gpurun_out/s2c2_ffma2.log
ZPCB200_LIB=$PWD/zpc_b200/build/exp/nt512.so timeout 600 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x -p no:cacheprovider -k "sweep4 or sweep8" > gpurun_out/s2c2_nt512_tests.log 2>&1; echo "nt512 tests rc=$?"; tail -3 gpurun_out/s2c2_nt512_tests.log
ZPCB200_LIB=$PWD/zpc_b200/build/exp/nt512.so timeout 600 python benchmarks/variants.py --config C3 --steps 8 --warmup 3 --combos 4:1,8:1,4:1 --tag nt512 > gpurun_out/s2c2_nt512.log 2> gpurun_out/s2c2_nt512.err; echo "nt512 rc=$?"; cut -c1-330 gpurun_out/s2c2_nt512.log
