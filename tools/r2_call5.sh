#!/bin/bash
# Round 2, fifth GPU call (1 GPU): fused split-KV merge, graph-captured batch bookkeeping; attention micro-bench; e2e host profile.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c5_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c5_${name}.log" | cut -c1-300)"
}
step tests_all 1500 python -m pytest tests -q -rs -m gpu
step bench_attn 300 python tools/bench_attn.py --out gpurun_out/r2c5_bench_attn.txt
step trace_l3 300 python tools/step_trace.py --workload llama3-8b-gptq --out gpurun_out/r2c5_step_trace_l3.txt
step profile_e2e 300 python tools/profile_e2e.py --steps 200
step bench 300 python bench.py --steps 24 --warmup 4 --no-extra --no-cpu-baseline
cat gpurun_out/r2c5_bench_attn.txt
head -14 gpurun_out/r2c5_step_trace_l3.txt
head -40 gpurun_out/r2c5_profile_e2e.log
grep -h '^{' gpurun_out/r2c5_bench.log | cut -c1-400
