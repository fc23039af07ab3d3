#!/bin/bash
# Round 2, third pass: A/B of the CSR kernel variants (the 256^3 product went from 0.257 to 0.272 ms between r1 and r2a),
# ncu of the one-pass persistent CG kernel, the small-golden tests.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for rep in 1 2; do
  python profiles/ab_spmv.py build/r1tree r1
  python profiles/ab_spmv.py . head
  for v in B C D; do VCL_B200_LIB_OVERRIDE=$PWD/build/ab_$v/libvcl_b200.so python profiles/ab_spmv.py . ab_$v; done
done 2>&1 | grep -v "^+" | tee gpurun_out/ab_spmv_r2c.log
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -k small 2>&1 | tail -5
for form in 1 2; do
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"cg_onepass|cg_persistent" -c 1 -o gpurun_out/prof_r2c_cg1024_form$form -f \
      python profiles/run_kernels.py cg1024 $form > gpurun_out/prof_r2c_cg1024_form$form.log 2>&1
done
ls -la gpurun_out | tail -5
