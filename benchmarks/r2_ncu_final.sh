#!/bin/bash
# launch list of the default bench command + one full capture of the dominant kernels (numbers printed under ncu are never bench values)
set -u
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 10 --warmup 2 --no-cpu-baseline --e2e-steps 0 --refcuda-steps 0 --prims-log2 0 > gpurun_out/r02_launch_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_launches.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'p2g_binned_kernel|g2p_binned_staged_kernel' -s 6 -c 2 \
  -o gpurun_out/r02_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --refcuda-steps 0 --prims-log2 0 > gpurun_out/r02_ncu.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/r02_full.ncu-rep
# sanitizers on the shared-memory / TMA / mbarrier kernels, small cases
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x -p no:cacheprovider \
  -k "binned_path_matches_oracle and cube8" > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mpm.py tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider \
  -k "(binned_path_matches_oracle and cube8) or status_word or single_particle or dense_cluster" > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_memcheck.log
