#!/bin/bash
# Round 2: A/B of the SELL kernel (gather before / after the value test), 3 repetitions, alternating.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for rep in 1 2 3; do
  python profiles/ab_spmv.py . head
  VCL_B200_LIB_OVERRIDE=$PWD/build/ab_SU/libvcl_b200.so python profiles/ab_spmv.py . sell_uncond
done 2>&1 | grep -v "^+" | tee gpurun_out/ab_sell_r2j.log
VCL_B200_LIB_OVERRIDE=$PWD/build/ab_SU/libvcl_b200.so timeout 600 python -m pytest tests -m gpu -q -k "sell" 2>&1 | tail -3
