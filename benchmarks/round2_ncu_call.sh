#!/bin/bash
# Profiles of the next round (one GPU; numbers printed under ncu are never bench values):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash benchmarks/round2_ncu_call.sh'
# 1. launch list of the default bench command (per-launch times: the kernels' SHARES of the step must agree with the live
#    CUDA-event stage times of bench.py);  2. one `--set full` capture of the dominant kernels at C3;  afterwards, here:
#      ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > profiles/r02_ncu_full_c3_raw.csv
#      python benchmarks/ncu_summary.py profiles/r02_ncu_full_c3_raw.csv C3 > profiles/r02_traffic.json   (bench.py reads r01_traffic.json: repoint it)
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 10 --warmup 2 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r02_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'p2g_binned_kernel|g2p_binned_staged_kernel' -s 6 -c 2 \
  -o gpurun_out/r02_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r02_ncu.log 2>&1
# 3. the packed-fp32 sweep: same capture with sweep 5, to compare issue-slot and shared-pipe utilisation with sweep 4
ncu --set full --clock-control none --import-source on -k regex:'p2g_binned_kernel' -s 3 -c 1 \
  -o gpurun_out/r02_full_sweep5 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --p2g-sweep 5 > gpurun_out/r02_ncu_sweep5.log 2>&1
ls -la gpurun_out/r02_*
