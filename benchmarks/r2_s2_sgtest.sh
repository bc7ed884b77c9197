#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_sparsegrid.py -m gpu -q -x -p no:cacheprovider -k "solver_matches" > gpurun_out/s2_sgtest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/s2_sgtest.log
