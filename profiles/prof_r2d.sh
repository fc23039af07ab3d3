#!/bin/bash
# Round 2, fourth pass, on a 2-GPU box: row-partitioned checks (incl. 128^3 against the reference golden), both bench arms at N=2 and N=1,
# the new foreign-plan test.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py --big > gpurun_out/dist_check_w2_r2d.log 2>&1; echo "dist_check w2 rc=$?"
grep -v "^W\|^\*" gpurun_out/dist_check_w2_r2d.log | tail -14
timeout 300 python tests/dist_check.py --big > gpurun_out/dist_check_w1_r2d.log 2>&1; echo "dist_check w1 rc=$?"
tail -3 gpurun_out/dist_check_w1_r2d.log
timeout 600 python -m pytest tests -m gpu -q -k "foreign or small or dist" 2>&1 | tail -5
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_n2_r2d.json 2> gpurun_out/bench_ref_n2_r2d.err; echo "ref n2 rc=$?"
cat gpurun_out/bench_ref_n2_r2d.json | head -c 1500
timeout 900 $TR --master-port 29513 bench.py --gpus 2 > gpurun_out/bench_n2_r2d.json 2> gpurun_out/bench_n2_r2d.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2_r2d.err
timeout 900 python bench.py > gpurun_out/bench_n1_r2d.json 2> gpurun_out/bench_n1_r2d.err; echo "bench n1 rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_n1_r2d.json 2> gpurun_out/bench_ref_n1_r2d.err; echo "ref n1 rc=$?"
