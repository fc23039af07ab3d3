#!/bin/bash
# 512^3 CG on 2 GPUs: halo push inside cg_update (fused_push) vs from the head of the product kernel.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for rep in 1 2; do
for v in fused head_push; do
  if [ $v = head_push ]; then export VCL_B200_NO_FUSED_PUSH=1; else unset VCL_B200_NO_FUSED_PUSH; fi
  $TR --master-port 29513 bench.py --gpus 8 --workload cg512 --steps 100 --warmup 10 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v cg512 N=8', d['value'], d['ms_per_step'])"
done
done 2>&1 | grep -v "^+"
