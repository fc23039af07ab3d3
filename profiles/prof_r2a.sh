#!/bin/bash
# Round 2, first pass: GPU test-suite (incl. the at-size parity tests), L2-policy experiment on config 1, default bench line,
# full ncu captures of the CSR / SELL SpMV kernels and of the persistent CG kernel.  Run under gpurun on one GPU.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
rm -f gpurun_out/parity_deltas.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
tail -5 gpurun_out/pytest_r2a.log
timeout 600 python profiles/cg1024_l2.py 1024 512 > gpurun_out/cg1024_l2_r2a.log 2>&1
cat gpurun_out/cg1024_l2_r2a.log
timeout 900 python bench.py > gpurun_out/bench_n1_r2a.json 2> gpurun_out/bench_n1_r2a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_n1_r2a.err
for what in spmv cg1024; do
  case $what in
    spmv) rx="csr_stream|sell_kernel"; skip=2; cnt=2; extra="";;
    cg1024) rx="cg_persistent"; skip=0; cnt=1; extra="--cache-control none";;
  esac
  timeout 600 ncu --set full --clock-control none $extra --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/prof_r2a_$what -f \
      python profiles/run_kernels.py $what > gpurun_out/prof_r2a_$what.log 2>&1
done
ls -la gpurun_out | tail -12
