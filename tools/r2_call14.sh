#!/bin/bash
# Round 2, fourteenth GPU call (1 GPU): final full GPU suite, the churn-aware graph policy on the 8B session, the default bench line.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c14_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c14_${name}.log" | cut -c1-300)"
}
step tests_all 400 python -m pytest tests -q -rs -m gpu
step mixed8b 200 python bench.py --workload llama3-8b-gptq-mixed --requests 192
step bench 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline
for f in mixed8b bench; do grep -h '^{' gpurun_out/r2c14_$f.log | cut -c1-1800; done
tail -6 gpurun_out/r2c14_tests_all.log
