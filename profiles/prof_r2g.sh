#!/bin/bash
# Round 2, seventh pass (8 GPUs): DIST_CHECK at world 8 (incl. 128^3 against the reference golden, BiCGStab, Jacobi-CG), bench lines at N = 8 and 4.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29511 tests/dist_check.py --big > gpurun_out/dist_check_w8_r2g.log 2>&1; echo "dist_check w8 rc=$?"
grep "DIST_CHECK\|FAIL\|128^3" gpurun_out/dist_check_w8_r2g.log | head -20
timeout 900 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 > gpurun_out/bench_n8_r2g.json 2> gpurun_out/bench_n8_r2g.err; echo "bench n8 rc=$?"
tail -c 600 gpurun_out/bench_n8_r2g.err
timeout 900 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 > gpurun_out/bench_n4_r2g.json 2> gpurun_out/bench_n4_r2g.err; echo "bench n4 rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --workload cg512 --steps 100 --warmup 10 > gpurun_out/bench_cg512_n8_r2g.json 2> gpurun_out/bench_cg512_n8_r2g.err; echo "cg512 n8 rc=$?"
