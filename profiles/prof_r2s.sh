#!/bin/bash
# Round 2, final 8-GPU record with the final code: bench lines at N = 8 / 4 / 2 (default workload incl. cg512), DIST_CHECK world 8.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  timeout 900 $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N > gpurun_out/bench_n${N}_r2s.json 2> gpurun_out/bench_n${N}_r2s.err; echo "bench n$N rc=$?"
done
timeout 900 $TR --nproc-per-node 8 --master-port 29521 tests/dist_check.py --big > gpurun_out/dist_check_w8_r2s.log 2>&1; echo "dist_check w8 rc=$?"
grep "DIST_CHECK\|FAIL" gpurun_out/dist_check_w8_r2s.log | head
python - <<'PY'
import json
for N in (8, 4, 2):
    d=json.loads([l for l in open('gpurun_out/bench_n%d_r2s.json' % N) if l.startswith('{')][-1])
    print(N, 'value %.0f ms %.4f launches %d bitexact %s sell %.4f cg512 %.1f' % (d['value'], d['ms_per_step'], d['gpu_launches'], d['parity']['spmv_bitexact'], d['sell_partitioned']['ms_per_step'], d['cg']['lap3d_512']['iterations_per_sec']))
PY
