#!/bin/bash
# What the driver runs at round end, in one go: GPU tests, smoke(), both bench arms (N = 1).
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
rm -f gpurun_out/parity_deltas.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('gpurun_out/bench_ref_final.json') if l.startswith('{')][-1])
print('ours value %.1f e2e %.1f ms %.4f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['clocks']))
print('ref  value %.1f e2e %.1f cores %d same_workload %s' % (r['value'], r['e2e']['value'], r['cpu_baseline']['cores'], r['config']['workload']==d['config']['workload']))
print('ratio value %.1f e2e %.2f' % (d['value']/r['value'], d['e2e']['value']/r['e2e']['value']))
print('sell', d['sell']['value'], 'cg1024', d['cg']['lap2d_1024']['iterations_per_sec'], 'cg512', d['cg']['lap3d_512']['iterations_per_sec'])
print('legacy', {k:v for k,v in d['legacy_cuda_baseline'].items() if k.startswith('ratio')})
PY
