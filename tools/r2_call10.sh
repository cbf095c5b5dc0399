#!/bin/bash
# Round 2, tenth GPU call (2 GPUs): the row-owner layer boundary - parity at world 2, latency against the previous kernel and NCCL.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c10_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c10_${name}.log" | cut -c1-300)"
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step tests_p2p_tp 600 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_tp.py -q -rs -m gpu -x
step boundary 300 $TR --nproc-per-node 2 --master-port 29541 tools/bench_boundary.py
step bench_n2 400 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 24 --warmup 4
grep -h "^boundary" gpurun_out/r2c10_boundary.log
grep -h '^{' gpurun_out/r2c10_bench_n2.log | cut -c1-1200
tail -5 gpurun_out/r2c10_tests_p2p_tp.log
