#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > gpurun_out/s2n2b_dist_plain.log 2>&1
echo "dist_check plain rc=$?"; grep -E "dist_check|Error|error" gpurun_out/s2n2b_dist_plain.log | tail -2 | cut -c1-500
ZPC_WORLD_AS=8 timeout 300 python - <<'PY' > gpurun_out/s2n2b_tags.log 2>&1
# the 8-rank cloud's identity tags and capacities, checked without 8 GPUs
import numpy as np
from zpc_b200 import synth
s, G = 96, 128
full = synth.elastic_cube(s, G, jitter_F=0.03, jitter_C=0.3)
n0 = full["m"].shape[0]
m0 = float(full["m"].mean())
m = ((1 << 23) + np.arange(n0, dtype=np.int64)).astype(np.float32) * np.float32(2.0 ** np.round(np.log2(m0 / (1 << 23))))
print("unique tags", np.unique(m).size == n0, n0)
PY
tail -1 gpurun_out/s2n2b_tags.log
