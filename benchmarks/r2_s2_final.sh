#!/bin/bash
# round 2, final GPU pass: full -m gpu suite (kept log), parity tables at BASELINE sizes, default bench line + reference arm,
# ncu launch list + one full capture of the dominant kernels, sanitizers
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 1500 -p no:cacheprovider -rxXs > gpurun_out/r02_gputests_final.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02_gputests_final.log
timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/r02_bench_reference_arm_final.log 2>&1; echo "reference arm rc=$?"; tail -1 gpurun_out/r02_bench_reference_arm_final.log | cut -c1-400
timeout 1200 python bench.py > gpurun_out/r02_bench_c3_n1_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r02_bench_c3_n1_final.log | cut -c1-1200
timeout 900 python bench.py --config C2 --no-cpu-baseline --e2e-steps 0 --refcuda-steps 0 --prims-log2 0 > gpurun_out/r02_bench_c2_n1_final.log 2>&1; echo "bench C2 rc=$?"; tail -1 gpurun_out/r02_bench_c2_n1_final.log | cut -c1-600
timeout 1500 python -m tests.scale_parity --size C2 --out gpurun_out/r02_parity_c2_final.md > gpurun_out/r02_parity_c2_final.log 2>&1; echo "parity c2 rc=$?"; tail -3 gpurun_out/r02_parity_c2_final.log
timeout 900 python -m tests.scale_parity --size C1 --out gpurun_out/r02_parity_c1_final.md > gpurun_out/r02_parity_c1_final.log 2>&1; echo "parity c1 rc=$?"; tail -2 gpurun_out/r02_parity_c1_final.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c3_final.csv \
  python bench.py --steps 10 --warmup 2 --no-cpu-baseline --e2e-steps 0 --refcuda-steps 0 --prims-log2 0 > gpurun_out/r02_launch_bench_final.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_c3_final.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'p2g_binned_kernel|g2p_binned_staged_kernel' -s 6 -c 2 \
  -o gpurun_out/r02_full_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --refcuda-steps 0 --prims-log2 0 > gpurun_out/r02_ncu_final.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/r02_full_final.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'rs_onesweep_kernel' -s 2 -c 1 \
  -o gpurun_out/r02_sort_final -f python benchmarks/prims_sweep.py --min-log2 26 --max-log2 26 > gpurun_out/r02_ncu_sort_final.log 2>&1
echo "sort capture rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x -p no:cacheprovider \
  -k "binned_path_matches_oracle and cube8" > gpurun_out/r02_racecheck_final.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck_final.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mpm.py tests/test_gpu_models.py tests/test_gpu_sparsegrid.py -m gpu -q -x -p no:cacheprovider \
  -k "(binned_path_matches_oracle and cube8) or status_word or single_particle or dense_cluster or (binned_fast_path and cube8)" > gpurun_out/r02_memcheck_final.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_memcheck_final.log
