#!/bin/bash
# Round 2, sixth GPU call (1 GPU): paged-convention / speculation tests, long-trajectory parity report, cuBLAS A/B, ncu captures.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c6_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c6_${name}.log" | cut -c1-300)"
}
step tests_all 1500 python -m pytest tests -q -rs -m gpu
step parity 900 python tools/parity_report.py --out gpurun_out/r2c6_parity_report.json
step bench_gemm 400 python tools/bench_gemm.py
step trace_l3 300 python tools/step_trace.py --workload llama3-8b-gptq --out gpurun_out/r2c6_step_trace_l3.txt
step trace_7b 300 python tools/step_trace.py --workload llama2-7b-gptq --out gpurun_out/r2c6_step_trace_7b.txt
step ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c6_l3_launches.csv \
     python tools/profile_decode.py --workload llama3-8b-gptq --layers 2 --ctx 1536 --steps 3
step ncu_full 900 ncu --set full --clock-control none --import-source on \
     -k regex:"gemm_w4a16|attn_decode_paged_kernel|attn_prefill_paged|gemm_f16|rmsnorm_residual|rope_kv_write|splitk_silu" -c 40 \
     -o gpurun_out/r2c6_l3_full python tools/profile_decode.py --workload llama3-8b-gptq --layers 2 --ctx 1536 --steps 2
ls -la gpurun_out/r2c6_* | head -20
tail -3 gpurun_out/r2c6_parity.log | cut -c1-600
grep "f16  " gpurun_out/r2c6_bench_gemm.log | head -20
