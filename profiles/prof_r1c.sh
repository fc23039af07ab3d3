#!/bin/bash
# Round-1, third pass (float build, persistent CG, fused diagonal preconditioners): profiling recipe, run under gpurun on one GPU.
# Outputs land in gpurun_out/, summaries are copied to profiles/ (summarize_ncu.py, ncu_key.py).
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
# 1. launch list of the default bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r1c.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r1c.log 2>&1
# 2. full captures of the hot kernels
for what in spmv cg cg1024 spmv_f32 bicgstab_jacobi gmres_jacobi; do
  extra=""
  case $what in
    spmv) rx="csr_stream|sell_kernel"; skip=2; cnt=2;;
    cg) rx="cg_update|csr_stream"; skip=3; cnt=2;;
    cg1024) rx="cg_persistent"; skip=0; cnt=1; extra="--cache-control none";;
    spmv_f32) rx="csr_stream|sell_kernel|ell_kernel"; skip=3; cnt=3;;
    bicgstab_jacobi) rx="pbicg|csr_stream"; skip=3; cnt=5;;
    gmres_jacobi) rx="gmres|csr_stream"; skip=60; cnt=4;;
  esac
  ncu --set full --clock-control none $extra --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/prof_r1c_$what -f \
      python profiles/run_kernels.py $what > gpurun_out/prof_r1c_$what.log 2>&1
done
ls -la gpurun_out | tail -20
