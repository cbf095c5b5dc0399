#!/bin/bash
# A/B of the environment-switched kernel variants on one B200 (run under gpurun; writes gpurun_out/ab_*.log).
#   gpurun --timeout 1500 -- 'bash tools/ab_switches.sh'
# 1. parity of every switch (tests/test_gpu_experimental.py) and of the Santacoder / Falcon families (tests/test_gpu_santacoder.py, test_gpu_falcon.py), 2. per-GEMM times, 3. the bench line with and without.
set -u
mkdir -p gpurun_out
B200_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_experimental.py tests/test_gpu_santacoder.py tests/test_gpu_falcon.py tests/test_gpu_generate.py::test_prompt_prefix_equals_the_same_tokens_typed_in tests/test_gpu_gemm.py::test_gemm_w4a16_model_shapes -q -m gpu > gpurun_out/ab_parity.log 2>&1
tail -5 gpurun_out/ab_parity.log
for sw in none B200_W4_CLUSTER B200_F16_ALIGNED; do
  echo "== GEMM times, $sw" >> gpurun_out/ab_gemm.log
  env $( [ $sw = none ] || echo $sw=1 ) timeout 600 python tools/bench_gemm.py --w4 >> gpurun_out/ab_gemm.log 2>&1
done
for sw in none B200_ATTN_PERSISTENT B200_W4_CLUSTER; do
  for model in llama2-7b-gptq llama3-8b-gptq tinyllama-fp16; do
    echo "== bench $model, $sw" >> gpurun_out/ab_bench.log
    env $( [ $sw = none ] || echo $sw=1 ) timeout 600 python bench.py --workload $model --steps 20 --warmup 5 --no-cpu-baseline >> gpurun_out/ab_bench.log 2>&1
  done
done
grep -h '"metric"\|^==' gpurun_out/ab_bench.log | cut -c1-220
