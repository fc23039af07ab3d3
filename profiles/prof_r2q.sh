#!/bin/bash
# Round 2 (2 GPUs): SELL pass rotation + fence-only-for-pushers: dist_check, kernel durations at ~16.7M rows per GPU (322^3 / 2), bench N = 2.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py --big > gpurun_out/dist_check_w2_r2q.log 2>&1; echo "dist_check w2 rc=$?"
grep "DIST_CHECK\|FAIL\|rror" gpurun_out/dist_check_w2_r2q.log | head
timeout 600 $TR --master-port 29531 profiles/dist_trace.py 322 64 > gpurun_out/dist_trace_w2_r2q.log 2>&1; echo "trace rc=$?"
grep "rank 0\|csr_stream\|cg_update" gpurun_out/dist_trace_w2_r2q.log | head -8
timeout 900 $TR --master-port 29513 bench.py --gpus 2 > gpurun_out/bench_n2_r2q.json 2> gpurun_out/bench_n2_r2q.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2_r2q.json') if l.startswith('{')][-1])
print('csr', d['ms_per_step'], 'sell', d['sell_partitioned'], 'cg512', d['cg']['lap3d_512']['iterations_per_sec'])
PY
