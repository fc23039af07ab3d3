#!/bin/bash
# Round 2: L2 prefetch (cp.async.bulk.prefetch.L2) of the block after next, by stride extrapolation -- A/B against head.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( for rep in 1 2; do
  python profiles/ab_stages.py head
  for v in PF1 PF2; do VCL_B200_LIB_OVERRIDE=$PWD/build/ab_$v/libvcl_b200.so python profiles/ab_stages.py $v; done
done ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_l2pf_r2m.log
