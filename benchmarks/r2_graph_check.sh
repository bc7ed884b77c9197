#!/bin/bash
set -u
mkdir -p gpurun_out
ZPC_GRAPH=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > gpurun_out/g_dist2.log 2>&1; echo "dist_check graph rc=$?"; grep -E "dist_check|Error" gpurun_out/g_dist2.log | cut -c1-400
