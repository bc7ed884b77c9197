#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 100 python benchmarks/sg_solver_bench.py --config C3 --steps 16 --warmup 4 > gpurun_out/s2_sg_solver.log 2> gpurun_out/s2_sg_solver.err; echo "rc=$?"; cut -c1-1500 gpurun_out/s2_sg_solver.log; tail -3 gpurun_out/s2_sg_solver.err
