#!/bin/bash
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for rep in 1 2 3; do
  python profiles/ab_spmv.py . head
  VCL_B200_LIB_OVERRIDE=$PWD/build/ab_pre/libvcl_b200.so python profiles/ab_spmv.py . pre_cols
done 2>&1 | grep -v "^+" | tee gpurun_out/ab_sell_r2k.log
