#!/bin/bash
# Round 2, thirteenth GPU call (8 GPUs): config[4] session and the default bench at N=8 with the 2048-row fused boundary.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c13_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c13_${name}.log" | cut -c1-300)"
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step mixed70b_n8 500 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --workload llama3-70b-gptq-mixed
step bench_n8 400 $TR --nproc-per-node 8 --master-port 29572 bench.py --gpus 8 --steps 24 --warmup 4
step boundary 120 $TR --nproc-per-node 8 --master-port 29573 tools/bench_boundary.py
grep -h "^boundary" gpurun_out/r2c13_boundary.log
for f in mixed70b_n8 bench_n8; do grep -h '^{' gpurun_out/r2c13_$f.log | cut -c1-2400; done
