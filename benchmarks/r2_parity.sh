#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m tests.scale_parity --size C2 --out gpurun_out/r02_parity_c2.md > gpurun_out/r02_parity_c2.log 2>&1; echo "c2 rc=$?"; tail -5 gpurun_out/r02_parity_c2.log
timeout 900 python -m tests.scale_parity --size C1 --out gpurun_out/r02_parity_c1.md > gpurun_out/r02_parity_c1.log 2>&1; echo "c1 rc=$?"; tail -3 gpurun_out/r02_parity_c1.log
timeout 900 python -m tests.scale_parity --size C2 --no-jitter --refs cuda --out gpurun_out/r02_parity_c2_rest.md > gpurun_out/r02_parity_c2_rest.log 2>&1; echo "c2 rest rc=$?"
