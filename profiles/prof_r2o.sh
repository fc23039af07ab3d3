#!/bin/bash
# Round 2 (2 GPUs): SELL slabs on the row-partitioned path (both transports, world 1 and 2), N = 2 bench line after the clocks refactor.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_w2_r2o.log 2>&1; echo "dist_check w2 rc=$?"
grep "SELL\|DIST_CHECK\|FAIL\|rror" gpurun_out/dist_check_w2_r2o.log | head -30
VCL_B200_DIST_TRANSPORT=nccl timeout 600 $TR --master-port 29514 tests/dist_check.py > gpurun_out/dist_check_w2_nccl_r2o.log 2>&1; echo "dist_check w2 nccl rc=$?"
grep "SELL\|DIST_CHECK\|FAIL" gpurun_out/dist_check_w2_nccl_r2o.log | head -20
timeout 300 python tests/dist_check.py > gpurun_out/dist_check_w1_r2o.log 2>&1; echo "dist_check w1 rc=$?"
grep "SELL\|DIST_CHECK\|FAIL" gpurun_out/dist_check_w1_r2o.log | head
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --no-extras > gpurun_out/bench_n2_r2o.json 2> gpurun_out/bench_n2_r2o.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/bench_n2_r2o.err; grep "^{" gpurun_out/bench_n2_r2o.json | head -c 400
