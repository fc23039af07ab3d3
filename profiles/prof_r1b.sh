#!/bin/bash
# Round-1 (second pass, after the TMA-pipeline rewrite) profiling recipe (run under gpurun, one GPU).  Outputs land in gpurun_out/, summaries are copied to profiles/.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
# 1. launch list of the default bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r1b.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
# 2. full captures of the hot kernels
for what in spmv cg bicgstab bicgstab_jacobi gmres cg_sell; do
  case $what in
    spmv) rx="csr_stream|sell_kernel"; skip=2; cnt=4;;
    cg) rx="cg_update|csr_stream"; skip=3; cnt=2;;
    cg_sell) rx="sell_kernel"; skip=2; cnt=1;;
    bicgstab) rx="bicgstab|csr_stream"; skip=1; cnt=4;;
    bicgstab_jacobi) rx="pbicg|csr_stream"; skip=3; cnt=5;;
    gmres) rx="gmres|csr_stream"; skip=60; cnt=8;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/prof_r1b_$what -f \
      python profiles/run_kernels.py $what > gpurun_out/prof_r1b_$what.log 2>&1
done
ls -la gpurun_out
