#!/bin/bash
# session 2, call 7: sort tile size / occupancy variants; reduce_prod and the new merge sort key types
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prims.py -m gpu -q -x --timeout 900 -k "not billion" > gpurun_out/s2c7_prims_tests.log 2>&1; echo "prims tests rc=$?"; tail -3 gpurun_out/s2c7_prims_tests.log
for v in "" rsitems20 rsitems24 rsminb5; do
  if [ -n "$v" ]; then export ZPCB200_LIB=$PWD/zpc_b200/build/exp/$v.so; fi
  echo "== ${v:-default}" >> gpurun_out/s2c7_sweep.log
  timeout 600 python benchmarks/prims_sweep.py --min-log2 24 --max-log2 28 >> gpurun_out/s2c7_sweep.log 2>> gpurun_out/s2c7_sweep.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/s2c7_sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('log2n','sort_pair_ms','sort_pair_24bit_ms','sort_frac','torch_sort_ms')})
PY
