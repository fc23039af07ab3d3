#!/bin/bash
# Round 2, second pass: GPU test-suite (all failures listed), CG driver forms on config 1, bench line under faulthandler.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
rm -f gpurun_out/parity_deltas.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -15 gpurun_out/pytest_r2b.log
timeout 600 python profiles/cg_forms.py 1024 512 2048 > gpurun_out/cg_forms_r2b.log 2>&1
cat gpurun_out/cg_forms_r2b.log
timeout 900 python -X faulthandler bench.py > gpurun_out/bench_n1_r2b.json 2> gpurun_out/bench_n1_r2b.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_n1_r2b.err
head -c 600 gpurun_out/bench_n1_r2b.json
