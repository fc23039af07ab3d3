#!/bin/bash
# Round 2, fifth pass (2 GPUs): row-partitioned BiCGStab / Jacobi-CG checks, plan variants, full GPU suite on the current code.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py --big > gpurun_out/dist_check_w2_r2e.log 2>&1; echo "dist_check w2 rc=$?"
grep -v "^W\|^\*\|OMP_NUM" gpurun_out/dist_check_w2_r2e.log | tail -40
VCL_B200_DIST_TRANSPORT=nccl timeout 600 $TR --master-port 29514 tests/dist_check.py > gpurun_out/dist_check_w2_nccl_r2e.log 2>&1; echo "dist_check w2 nccl rc=$?"
grep -v "^W\|^\*\|OMP_NUM" gpurun_out/dist_check_w2_nccl_r2e.log | tail -24
timeout 300 python tests/dist_check.py > gpurun_out/dist_check_w1_r2e.log 2>&1; echo "dist_check w1 rc=$?"
tail -12 gpurun_out/dist_check_w1_r2e.log
timeout 300 python profiles/ab_plans.py 2>&1 | tee gpurun_out/ab_plans_r2.log
rm -f gpurun_out/parity_deltas.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_r2e.log
