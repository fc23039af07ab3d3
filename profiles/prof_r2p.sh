#!/bin/bash
# Round 2 (8 GPUs): where do the ~55 us per iteration of the 512^3 CG go at 8 GPUs?  CUPTI kernel durations (dist_trace.py) and the
# %globaltimer stamps of the VCL_PEER_DEBUG build; final DIST_CHECK at world 8 with all solvers and SELL slabs.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 profiles/dist_trace.py 512 64 > gpurun_out/dist_trace_w8_r2p.log 2>&1; echo "trace rc=$?"
grep -v "^W\|^\*\|OMP_NUM\|^$" gpurun_out/dist_trace_w8_r2p.log | tail -30
VCL_B200_LIB_OVERRIDE=$PWD/build/ab_dbg/libvcl_b200.so timeout 600 $TR --master-port 29532 bench.py --gpus 8 --workload cg512 --steps 100 --warmup 10 > gpurun_out/peer_debug_w8_r2p.json 2> gpurun_out/peer_debug_w8_r2p.err; echo "debug rc=$?"
grep "peer debug" gpurun_out/peer_debug_w8_r2p.err | head -24
timeout 900 $TR --master-port 29511 tests/dist_check.py --big > gpurun_out/dist_check_w8_r2p.log 2>&1; echo "dist_check w8 rc=$?"
grep "DIST_CHECK\|FAIL" gpurun_out/dist_check_w8_r2p.log | head
timeout 900 $TR --master-port 29513 bench.py --gpus 8 > gpurun_out/bench_n8_r2p.json 2> gpurun_out/bench_n8_r2p.err; echo "bench n8 rc=$?"
