#!/bin/bash
# compute-sanitizer over the kernels' parity tests on one B200 (the reference has no sanitizer coverage at all).
#   gpurun --timeout 1700 -- 'bash tools/sanitize.sh'
# memcheck on every op test, racecheck (shared-memory hazards) and synccheck (mbarrier / bar.sync misuse) on the three
# kernels with hand-rolled pipelines.  Output: gpurun_out/sanitize_*.log, one summary line each on stdout.
set -u
mkdir -p gpurun_out
run() {  # tool, log suffix, pytest selection
  local tool=$1 name=$2; shift 2
  timeout 800 compute-sanitizer --tool "$tool" --error-exitcode 3 --launch-timeout 120 \
    python -m pytest "$@" -x -q -m gpu -p no:cacheprovider > "gpurun_out/sanitize_${name}.log" 2>&1
  echo "$name: exit $? $(grep -c 'ERROR SUMMARY' "gpurun_out/sanitize_${name}.log") summaries, $(grep -h 'ERROR SUMMARY' "gpurun_out/sanitize_${name}.log" | tail -1)"
}
run memcheck  mem_ops   tests/test_gpu_ops.py
run memcheck  mem_gemm  tests/test_gpu_gemm.py -k "not large"
run racecheck race_attn tests/test_gpu_ops.py -k "attn_decode or attn_prefill"
run racecheck race_gemm tests/test_gpu_gemm.py -k "w4 or gptq"
run synccheck sync_all  tests/test_gpu_ops.py tests/test_gpu_gemm.py
