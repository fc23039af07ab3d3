#!/bin/bash
# Round 2: compute-sanitizer over the round-2 kernels (memcheck, racecheck), one GPU.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_r2.py > gpurun_out/sanitize_r2_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -12 gpurun_out/sanitize_r2_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python profiles/sanitize_r2.py > gpurun_out/sanitize_r2_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/sanitize_r2_racecheck.log
timeout 900 compute-sanitizer --tool memcheck viennacl-dev_b200/ref_binding/_build/ref_tree_test > gpurun_out/sanitize_r2_refbind.log 2>&1; echo "refbind memcheck rc=$?"
tail -4 gpurun_out/sanitize_r2_refbind.log
