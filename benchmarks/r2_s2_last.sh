#!/bin/bash
# last call of round 2: the whole -m gpu suite on the final tree, one short bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02_gputests_last.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_gputests_last.log
timeout 200 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --refcuda-steps 0 --prims-log2 0 --e2e-steps 0 > gpurun_out/r02_bench_last.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r02_bench_last.log | cut -c1-200
