#!/bin/bash
# Round 2, sixth pass: stage-depth variants of the CSR ring (in-flight bytes vs latency), foreign-plan test.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( python profiles/ab_stages.py head
  for v in S3 S3c S4c; do VCL_B200_LIB_OVERRIDE=$PWD/build/ab_$v/libvcl_b200.so python profiles/ab_stages.py $v; done ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_stages_r2f.log
timeout 600 python -m pytest tests -m gpu -q -k "foreign or small" 2>&1 | tail -5
