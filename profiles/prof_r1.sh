set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"csr_stream|sell_kernel" -s 2 -c 4 -o gpurun_out/prof_spmv python profiles/run_kernels.py spmv > gpurun_out/prof_spmv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cg_update|csr_stream" -s 2 -c 4 -o gpurun_out/prof_cg python profiles/run_kernels.py cg > gpurun_out/prof_cg.log 2>&1
ls -la gpurun_out
