#!/bin/bash
# Round 2, fourth GPU call (2 GPUs): tcgen05 paged prefill attention, continuous-batching session, new bench.py.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c4_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c4_${name}.log" | cut -c1-300)"
}
nvidia-smi -L
step pf_paged 300 python -m pytest tests/test_gpu_ops.py -q -rs -m gpu -k "prefill_paged"
step tests_all 1500 python -m pytest tests -q -rs -m gpu
B200_PREFILL_PAGED=0 step tests_varlen_prefill 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_generate.py tests/test_gpu_server.py -q -rs -m gpu
step bench_default 600 python bench.py --steps 24 --warmup 4
B200_PREFILL_PAGED=0 step bench_varlen 300 python bench.py --steps 24 --warmup 4 --no-extra --no-cpu-baseline
step bench_mixed 400 python bench.py --workload llama3-8b-gptq-mixed --requests 192
step bench_tp2 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5
step bench_ref 300 python bench.py --impl reference --steps 4 --warmup 3
for f in bench_default bench_varlen bench_mixed bench_tp2 bench_ref; do grep -h '^{' gpurun_out/r2c4_$f.log | cut -c1-600; done
