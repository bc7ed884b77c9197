#!/bin/bash
# session 2, call 8 (1 GPU): SparseGrid fast path with every model, smoke()
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_sparsegrid.py -m gpu -q -x -p no:cacheprovider -k "sparsegrid" > gpurun_out/s2c8_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s2c8_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2c8_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s2c8_smoke.log
