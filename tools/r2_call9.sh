#!/bin/bash
# Round 2, ninth GPU call (8 GPUs): the driver's N=8 / N=4 bench lines, the config[4] session at TP8, the TP8 launch list.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c9_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c9_${name}.log" | cut -c1-300)"
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step bench_n8 500 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 24 --warmup 4
step mixed70b_n8 600 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --workload llama3-70b-gptq-mixed
step bench_n4 400 $TR --nproc-per-node 4 --master-port 29533 bench.py --gpus 4 --steps 24 --warmup 4 --no-extra
step trace_n8 300 $TR --nproc-per-node 8 --master-port 29534 tools/step_trace.py --workload llama3-8b-gptq --out gpurun_out/r2c9_step_trace_llama3_8b_tp8.txt
step trace70b_n8 400 $TR --nproc-per-node 8 --master-port 29535 tools/step_trace.py --workload llama3-70b-gptq --out gpurun_out/r2c9_step_trace_llama3_70b_tp8.txt
for f in bench_n8 mixed70b_n8 bench_n4; do grep -h '^{' gpurun_out/r2c9_$f.log | cut -c1-1500; done
head -14 gpurun_out/r2c9_step_trace_llama3_8b_tp8.txt
head -14 gpurun_out/r2c9_step_trace_llama3_70b_tp8.txt
