#!/bin/bash
# Round 2, final profiling pass (1 GPU): launch list of the default bench command, full ncu captures of the hot kernels with the final
# code (CSR, SELL, SPLIT instantiation, fused CG pair, one-pass / two-phase persistent CG), full GPU test-suite, default bench line.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
rm -f gpurun_out/parity_deltas.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2h.log
tail -6 gpurun_out/pytest_r2h.log
timeout 900 python bench.py > gpurun_out/bench_n1_r2h.json 2> gpurun_out/bench_n1_r2h.err; echo "bench rc=$?"
for L in 1 2 3 4; do VCL_BENCH_LANES=$L timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $L e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'value', d['value'])"; done 2>&1 | tee gpurun_out/e2e_lanes_r2h.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r2h.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2h.log 2>&1
for what in spmv spmv_split cg cg1024_1 cg1024_2; do
  extra=""; arg=""
  case $what in
    spmv) rx="csr_stream|sell_kernel"; skip=2; cnt=2;;
    spmv_split) rx="csr_stream"; skip=1; cnt=1;;
    cg) rx="cg_update|csr_stream"; skip=3; cnt=2;;
    cg1024_1) rx="cg_onepass"; skip=0; cnt=1; extra="--cache-control none"; arg="1";;
    cg1024_2) rx="cg_persistent"; skip=0; cnt=1; extra="--cache-control none"; arg="2";;
  esac
  timeout 600 ncu --set full --clock-control none $extra --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/prof_r2h_$what -f \
      python profiles/run_kernels.py ${what%%_[12]} $arg > gpurun_out/prof_r2h_$what.log 2>&1
done
ls -la gpurun_out | tail -12
