#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 32 --warmup 8 ${BENCH_ARGS:-} > gpurun_out/${TAG:-r2}_bench_n$N.log 2> gpurun_out/${TAG:-r2}_bench_n$N.err
echo "rc=$?"; tail -1 gpurun_out/${TAG:-r2}_bench_n$N.log | cut -c1-3000; tail -5 gpurun_out/${TAG:-r2}_bench_n$N.err
