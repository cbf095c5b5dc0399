#!/bin/bash
# Round 2, eleventh GPU call (8 GPUs): the row-owner layer boundary at world 4 and 8 - parity, latency, the N=8 bench lines.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c11_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c11_${name}.log" | cut -c1-300)"
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step tests_p2p 400 python -m pytest tests/test_gpu_p2p.py -q -rs -m gpu -k "4 or 8"
step boundary 200 $TR --nproc-per-node 8 --master-port 29551 tools/bench_boundary.py
step bench_n8 500 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 24 --warmup 4
step mixed70b_n8 600 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --workload llama3-70b-gptq-mixed
step trace_n8 300 $TR --nproc-per-node 8 --master-port 29554 tools/step_trace.py --workload llama3-8b-gptq --out gpurun_out/r2c11_step_trace_llama3_8b_tp8.txt
step trace70b_n8 400 $TR --nproc-per-node 8 --master-port 29555 tools/step_trace.py --workload llama3-70b-gptq --out gpurun_out/r2c11_step_trace_llama3_70b_tp8.txt
tail -4 gpurun_out/r2c11_tests_p2p.log
grep -h "^boundary" gpurun_out/r2c11_boundary.log
for f in bench_n8 mixed70b_n8; do grep -h '^{' gpurun_out/r2c11_$f.log | cut -c1-1500; done
head -12 gpurun_out/r2c11_step_trace_llama3_8b_tp8.txt
head -12 gpurun_out/r2c11_step_trace_llama3_70b_tp8.txt
