#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-2}
for halo in fused p2p fused p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 32 --warmup 8 --e2e-steps 0 --no-parity-check --halo $halo ${BENCH_ARGS:-} > gpurun_out/ab.log 2> gpurun_out/ab.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
f=d["fused_step"]; k=f["kernels"]
print("$halo ms/step %.3f gap %.3f halo %.3f clean %.3f p2g %.3f upd %.3f g2p %.3f rebin %.3f part %.3f"%(d["ms_per_step"], f["gap_ms"], f["halo_ms"], k["clean"]["ms"], k["p2g"]["ms"], k["grid_update"]["ms"], k["g2p"]["ms"], f["rebin_ms_each"], f["partition_ms"]))
PY
done
