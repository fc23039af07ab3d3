#!/bin/bash
# Round 2: float build with 5 / 6 resident CTAs per SM (8-byte entries: the ring is 32.8 KB per CTA) -- A/B against head (4 CTAs).
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( for rep in 1 2; do
  echo "== head"; python profiles/float_time.py
  for v in F5 F6; do echo "== $v"; VCL_B200_LIB_OVERRIDE=$PWD/build/ab_$v/libvcl_b200.so python profiles/float_time.py; done
done ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_float_ctas_r2n.log
VCL_B200_LIB_OVERRIDE=$PWD/build/ab_F6/libvcl_b200.so timeout 900 python -m pytest tests/test_gpu_float.py -m gpu -q 2>&1 | tail -3
