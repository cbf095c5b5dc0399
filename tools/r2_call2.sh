#!/bin/bash
# Round 2, second GPU call (2 GPUs): TP / p2p parity that needs two devices, p2p A/B, step traces, fp16 GEMM A/B.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/r2_call2.sh'
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c2_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c2_${name}.log" | cut -c1-300)"
}
nvidia-smi -L
# the step traces run on GPU 1 while the 2-GPU tests wait for nothing: keep it simple, run sequentially
B200_EXPERIMENTAL=1 step tp_tests 600 python -m pytest tests/test_gpu_tp.py tests/test_gpu_p2p.py -q -rs -m gpu
B200_EXPERIMENTAL=1 B200_P2P_ALLREDUCE=1 step tp_tests_p2p 600 python -m pytest tests/test_gpu_tp.py -q -rs -m gpu
for sw in 0 1; do
  echo "== bench llama2-7b-fp16 tp2, B200_P2P_ALLREDUCE=$sw" >> gpurun_out/r2c2_p2p_bench.log
  B200_P2P_ALLREDUCE=$sw timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 >> gpurun_out/r2c2_p2p_bench.log 2>&1
done
grep -h '"metric"\|^==' gpurun_out/r2c2_p2p_bench.log | cut -c1-260
step trace_l3 300 python tools/step_trace.py --workload llama3-8b-gptq --out gpurun_out/r2c2_step_trace_l3.txt
step trace_7b 300 python tools/step_trace.py --workload llama2-7b-gptq --out gpurun_out/r2c2_step_trace_7b.txt
for sw in none B200_F16_ALIGNED; do
  echo "== GEMM times, $sw" >> gpurun_out/r2c2_ab_gemm_f16.log
  env $( [ $sw = none ] || echo $sw=1 ) timeout 400 python tools/bench_gemm.py >> gpurun_out/r2c2_ab_gemm_f16.log 2>&1
done
grep "^==\|f16" gpurun_out/r2c2_ab_gemm_f16.log
head -30 gpurun_out/r2c2_step_trace_l3.txt
