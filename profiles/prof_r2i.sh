#!/bin/bash
# Round 2 (2 GPUs): row-partitioned GMRES checks (both transports, world 1 and 2).
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_w2_r2i.log 2>&1; echo "dist_check w2 rc=$?"
grep "gmres\|DIST_CHECK\|FAIL\|Error\|error" gpurun_out/dist_check_w2_r2i.log | head -30
VCL_B200_DIST_TRANSPORT=nccl timeout 600 $TR --master-port 29514 tests/dist_check.py > gpurun_out/dist_check_w2_nccl_r2i.log 2>&1; echo "dist_check w2 nccl rc=$?"
grep "gmres\|DIST_CHECK\|FAIL" gpurun_out/dist_check_w2_nccl_r2i.log | head -20
timeout 300 python tests/dist_check.py > gpurun_out/dist_check_w1_r2i.log 2>&1; echo "dist_check w1 rc=$?"
grep "gmres\|DIST_CHECK\|FAIL" gpurun_out/dist_check_w1_r2i.log | head
timeout 900 python -m pytest tests -m gpu -q -x -k "gmres or dist" 2>&1 | tail -4
