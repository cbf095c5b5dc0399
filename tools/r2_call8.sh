#!/bin/bash
# Round 2, eighth GPU call (2 GPUs): chooser in-graph test, persistent prefill GEMM, vectorised RoPE, 70B-mixed dry run at TP2.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c8_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c8_${name}.log" | cut -c1-300)"
}
step tests_gemm_ops 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_chooser.py -q -rs -m gpu
step tests_all 1500 python -m pytest tests -q -rs -m gpu
step bench 400 python bench.py --steps 24 --warmup 4 --no-cpu-baseline
step bench_gemm_w4 300 python tools/bench_gemm.py --w4
step mixed70b_tp2 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus 2 --workload llama3-70b-gptq-mixed --layers 4 --requests 64
step mixed8b 400 python bench.py --workload llama3-8b-gptq-mixed --requests 192
for f in bench mixed70b_tp2 mixed8b; do grep -h '^{' gpurun_out/r2c8_$f.log | cut -c1-900; done
grep "w4a16" gpurun_out/r2c8_bench_gemm_w4.log | tail -8
