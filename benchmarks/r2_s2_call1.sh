#!/bin/bash
# session 2, call 1: the atomic-free flush (sweep 7) and the Jacobi early-out (sweep 8): parity, racecheck, A/B timing at C3, G2P2G pin
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x -p no:cacheprovider -k "sweep7 or sweep8" > gpurun_out/s2c1_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s2c1_tests.log
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider -k "g2p2g" -s > gpurun_out/s2c1_g2p2g.log 2>&1; echo "g2p2g rc=$?"; tail -5 gpurun_out/s2c1_g2p2g.log
timeout 600 python benchmarks/variants.py --config C3 --steps 8 --warmup 3 --combos 4:1,7:1,8:1,4:1,7:1,8:1 > gpurun_out/s2c1_ab.log 2> gpurun_out/s2c1_ab.err; echo "ab rc=$?"; cut -c1-330 gpurun_out/s2c1_ab.log
ZPCB200_P2G_SWEEP=7 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x -p no:cacheprovider \
  -k "binned_path_matches_oracle and cube8 and sweep7" > gpurun_out/s2c1_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/s2c1_racecheck.log
