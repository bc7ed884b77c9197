#!/bin/bash
set -u
mkdir -p gpurun_out
for pf in 0 300 448 600 900; do
  ZPCB200_PLANE_PREFETCH=$pf timeout 300 python benchmarks/variants.py --combos 6:1 --tag pf$pf 2>/dev/null | cut -c1-200
done
