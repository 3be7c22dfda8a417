#!/bin/bash
# time the Lorenz-96 step kernel variants at n = 1e8 (MB_L96_VARIANT: see l96_dispatch in csrc/pf_l96.cu)
for v in ${@:-20}; do
  echo "== MB_L96_VARIANT=$v"
  MB_L96_VARIANT=$v timeout 300 python scratch/c3_bench.py 1e8 2>&1 | grep -E "pf_l96 step|PF full step|rs_"
done
