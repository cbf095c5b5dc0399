#!/bin/bash
# Round 2, first GPU call (1 GPU): everything written in round 1 that never ran, plus the two bench lines.
#   gpurun --timeout 2700 -- 'bash tools/r2_call1.sh'
set -u
mkdir -p gpurun_out
step() {  # name, timeout seconds, command...
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c1_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c1_${name}.log" | cut -c1-300)"
}
nvidia-smi -L
B200_EXPERIMENTAL=1 step tests 1500 python -m pytest tests -m gpu -q -rs
step bench 400 python bench.py --steps 24 --warmup 4
step bench_l3 400 python bench.py --workload llama3-8b-gptq --steps 24 --warmup 4 --no-cpu-baseline
for sw in none B200_W4_CLUSTER B200_F16_ALIGNED; do
  echo "== GEMM times, $sw" >> gpurun_out/r2c1_ab_gemm.log
  env $( [ $sw = none ] || echo $sw=1 ) timeout 400 python tools/bench_gemm.py --w4 >> gpurun_out/r2c1_ab_gemm.log 2>&1
done
for sw in B200_ATTN_PERSISTENT B200_W4_CLUSTER; do
  for model in llama2-7b-gptq llama3-8b-gptq; do
    echo "== bench $model, $sw" >> gpurun_out/r2c1_ab_bench.log
    env $sw=1 timeout 400 python bench.py --workload $model --steps 20 --warmup 5 --no-cpu-baseline >> gpurun_out/r2c1_ab_bench.log 2>&1
  done
done
grep -h '"metric"\|^==' gpurun_out/r2c1_bench.log gpurun_out/r2c1_bench_l3.log gpurun_out/r2c1_ab_bench.log | cut -c1-330
tail -n 60 gpurun_out/r2c1_ab_gemm.log
