#!/bin/bash
# Gather-vector form of the row-partitioned CG (p and its halo contiguous inside the window): correctness and A/B against VCL_B200_NO_GVEC=1.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
NP=${NP:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py --big 2>&1 | grep "DIST_CHECK\|FAIL\|128^3\|rror" | head -12
for v in gvec nogvec gvec nogvec; do
  if [ $v = nogvec ]; then export VCL_B200_NO_GVEC=1; else unset VCL_B200_NO_GVEC; fi
  $TR --master-port 29513 bench.py --gpus $NP --workload cg512 --steps 100 --warmup 10 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v cg512 N=$NP', d['value'], d['ms_per_step'])"
done
unset VCL_B200_NO_GVEC
$TR --master-port 29514 bench.py --gpus $NP 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default bench N=$NP cg512', d['cg']['lap3d_512']['iterations_per_sec'], d['cg']['lap3d_512']['parity'])"
