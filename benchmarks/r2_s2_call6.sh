#!/bin/bash
# session 2, call 6: SparseGrid block-binned fast path (parity + timing vs the any-order kernels), sort tile size, reduce_prod
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparsegrid.py -m gpu -q -x -p no:cacheprovider > gpurun_out/s2c6_sg_tests.log 2>&1; echo "sg tests rc=$?"; tail -4 gpurun_out/s2c6_sg_tests.log
timeout 900 python -m pytest tests/test_gpu_prims.py -m gpu -q -x --timeout 900 -k "not billion" > gpurun_out/s2c6_prims_tests.log 2>&1; echo "prims tests rc=$?"; tail -2 gpurun_out/s2c6_prims_tests.log
timeout 600 python benchmarks/sg_fast.py --config C2 > gpurun_out/s2c6_sg_fast.log 2> gpurun_out/s2c6_sg_fast.err; echo "sg_fast rc=$?"; cut -c1-900 gpurun_out/s2c6_sg_fast.log; tail -3 gpurun_out/s2c6_sg_fast.err
for v in "" rsitems12 rsitems20; do
  if [ -n "$v" ]; then export ZPCB200_LIB=$PWD/zpc_b200/build/exp/$v.so; fi
  echo "== ${v:-default}" >> gpurun_out/s2c6_sweep.log
  timeout 600 python benchmarks/prims_sweep.py --min-log2 24 --max-log2 28 >> gpurun_out/s2c6_sweep.log 2>> gpurun_out/s2c6_sweep.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/s2c6_sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('log2n','sort_pair_ms','sort_pair_24bit_ms','sort_frac','torch_sort_ms')})
PY
