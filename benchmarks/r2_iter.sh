#!/bin/bash
# quick iteration call: parity of the binned path on the plane sweep, then A/B timing; NCU=1 adds a full capture of the plane kernel
set -u
mkdir -p gpurun_out
TAG=${TAG:-iter}
timeout 900 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x --timeout 600 -k "plane or not sweep" > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 400 python benchmarks/variants.py --combos ${COMBOS:-4:1,6:1,6:1} --tag $TAG > gpurun_out/${TAG}_variants.jsonl 2> gpurun_out/${TAG}_variants.err
cut -c1-330 gpurun_out/${TAG}_variants.jsonl; tail -3 gpurun_out/${TAG}_variants.err
if [ "${NCU:-0}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-p2g_plane_kernel}" -s 2 -c 1 \
    -o gpurun_out/${TAG} -f python benchmarks/variants.py --combos 6:1 --steps 1 --warmup 2 > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
fi
