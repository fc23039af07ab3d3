#!/bin/bash
# A/B on one 2-GPU box: HEAD (interior blocks gather with the plain addressing) vs the previous commit, 512^3 CG and kernel durations.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for rep in 1 2; do
for v in head pre; do
  if [ $v = pre ]; then export VCL_B200_LIB_OVERRIDE=$PWD/build/ab_pre/libvcl_b200.so; else unset VCL_B200_LIB_OVERRIDE; fi
  $TR --master-port 29513 bench.py --gpus 2 --workload cg512 --steps 100 --warmup 10 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v cg512 N=2', d['value'], d['ms_per_step'])"
  $TR --master-port 29531 profiles/dist_trace.py 322 64 2>/dev/null | grep "rank 0" -A3 | grep "csr_stream\|cg_update" | sed "s/^/$v /"
done
done 2>&1 | grep -v "^+" | tee gpurun_out/ab_interior_r2r.log
