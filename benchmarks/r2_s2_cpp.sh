#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_cpp_host.py -m gpu -q -x -p no:cacheprovider -s > gpurun_out/s2_cpp.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/s2_cpp.log | cut -c1-300
