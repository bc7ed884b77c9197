#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'p2g_plane_kernel' -s 2 -c 1 \
  -o gpurun_out/r2_plane -f python benchmarks/variants.py --combos 6:1 --steps 1 --warmup 2 > gpurun_out/r2_ncu_plane.log 2>&1
tail -3 gpurun_out/r2_ncu_plane.log; ls -la gpurun_out/r2_plane*
