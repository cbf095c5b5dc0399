#!/bin/bash
# Round 2, third GPU call (2 GPUs): deferred split-K + fused NVLink boundary: parity, then A/B benches and a step trace.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c3_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c3_${name}.log" | cut -c1-300)"
}
nvidia-smi -L
step tests_new 900 python -m pytest tests/test_gpu_splitk.py tests/test_gpu_p2p.py tests/test_gpu_tp.py -q -rs -m gpu -x
step tests_all 1500 python -m pytest tests -q -rs -m gpu
B200_P2P_ALLREDUCE=0 step tests_tp_nccl 600 python -m pytest tests/test_gpu_tp.py -q -rs -m gpu
B200_DEFER_SPLITK=0 step tests_nodefer 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_generate.py -q -rs -m gpu
for df in 1 0; do
  for model in llama3-8b-gptq llama2-7b-gptq; do
    echo "== bench $model, B200_DEFER_SPLITK=$df" >> gpurun_out/r2c3_bench.log
    B200_DEFER_SPLITK=$df timeout 400 python bench.py --workload $model --steps 24 --warmup 4 --no-cpu-baseline >> gpurun_out/r2c3_bench.log 2>&1
  done
done
for sw in 1 0; do
  echo "== bench llama2-7b-fp16 tp2, B200_P2P_ALLREDUCE=$sw" >> gpurun_out/r2c3_bench.log
  B200_P2P_ALLREDUCE=$sw timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 >> gpurun_out/r2c3_bench.log 2>&1
done
grep -h '"metric"\|^==' gpurun_out/r2c3_bench.log | cut -c1-230
step trace_l3 300 python tools/step_trace.py --workload llama3-8b-gptq --out gpurun_out/r2c3_step_trace_l3.txt
head -24 gpurun_out/r2c3_step_trace_l3.txt
