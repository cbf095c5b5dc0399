#!/bin/bash
# Round 2, seventh GPU call (1 GPU): fused chooser tests, traces, ncu captures summarised on the box (reports are too large to bring back).
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c7_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c7_${name}.log" | cut -c1-300)"
}
step tests_chooser 600 python -m pytest tests/test_gpu_chooser.py -q -rs -m gpu
step tests_all 1500 python -m pytest tests -q -rs -m gpu
step trace_l3 300 python tools/step_trace.py --workload llama3-8b-gptq --out gpurun_out/r2c7_step_trace_l3.txt
step trace_7b 300 python tools/step_trace.py --workload llama2-7b-gptq --out gpurun_out/r2c7_step_trace_7b.txt
step bench_attn 300 python tools/bench_attn.py --out gpurun_out/r2c7_bench_attn.txt
step ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c7_l3_launches.csv \
     python tools/profile_decode.py --workload llama3-8b-gptq --layers 2 --ctx 1536 --steps 3
# decode-step kernels of the second eager step: skip the prefill + first decode step's launches of these kernels
step ncu_full 900 ncu --set full --clock-control none \
     -k regex:"gemm_w4a16|attn_decode_paged_kernel|attn_prefill_paged|gemm_f16|rmsnorm_residual|rope_kv_write|splitk_silu" -c 30 \
     -o /tmp/r2c7_l3_full python tools/profile_decode.py --workload llama3-8b-gptq --layers 2 --ctx 1536 --steps 2
ncu -i /tmp/r2c7_l3_full.ncu-rep --page raw --csv > gpurun_out/r2c7_l3_full_raw.csv 2> gpurun_out/r2c7_ncu_export.log
ls -la gpurun_out/ | head -30
tail -5 gpurun_out/r2c7_tests_chooser.log | cut -c1-400
