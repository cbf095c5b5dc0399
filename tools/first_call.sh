#!/bin/bash
# Everything the first GPU call of a round should answer, in one box acquisition (1 GPU, ~45 min):
#   gpurun --timeout 3000 -- 'bash tools/first_call.sh'
# 1. the GPU parity suite, 2. the default bench line, 3. parity + A/B of the switched variants, 4. memcheck of the op tests.
# Every step has its own timeout and log under gpurun_out/first_*; a failing step does not stop the next one.
set -u
mkdir -p gpurun_out
step() {  # name, timeout seconds, command...
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/first_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/first_${name}.log" | cut -c1-200)"
}
step tests 1500 python -m pytest tests -m gpu -x -q
step bench 600 python bench.py --steps 24 --warmup 4
step bench_l3 600 python bench.py --workload llama3-8b-gptq --steps 24 --warmup 4 --no-cpu-baseline
bash tools/ab_switches.sh 2>&1 | tail -n 40
step memcheck 800 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider
grep -h '"metric"' gpurun_out/first_bench.log gpurun_out/first_bench_l3.log | cut -c1-400
