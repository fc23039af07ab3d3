#!/bin/bash
# Gather-vector CG at 8 GPUs: DIST_CHECK and one cg512 line.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py --big > gpurun_out/dist_check_w8_r2w.log 2>&1; echo "rc=$?"
grep "DIST_CHECK\|FAIL\|rror" gpurun_out/dist_check_w8_r2w.log | head -6
$TR --master-port 29513 bench.py --gpus 8 --workload cg512 --steps 100 --warmup 10 2>/dev/null | grep "^{" > gpurun_out/bench_cg512_n8_r2w.json
python -c "import json; d=json.loads(open('gpurun_out/bench_cg512_n8_r2w.json').read()); print('gvec cg512 N=8', d['value'], d['ms_per_step'], d['gpu_launches'])"
