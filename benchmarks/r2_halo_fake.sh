#!/bin/bash
mkdir -p gpurun_out
python benchmarks/halo_fake.py --config C2 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_update_bc_kernel -s 8 -c 1 -o gpurun_out/r2_updhalo -f python benchmarks/halo_fake.py --config C2 > gpurun_out/r2_updhalo.log 2>&1
tail -2 gpurun_out/r2_updhalo.log
