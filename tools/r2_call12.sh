#!/bin/bash
# Round 2, twelfth GPU call (2 GPUs): fused boundary for up to 2048 rows (TP prefill chunks in one C call) + full GPU suite.
set -u
mkdir -p gpurun_out
step() {
  local name=$1 limit=$2; shift 2
  local t0=$SECONDS
  timeout "$limit" "$@" > "gpurun_out/r2c12_${name}.log" 2>&1
  echo "[$name] exit $? in $((SECONDS - t0)) s: $(tail -n 1 "gpurun_out/r2c12_${name}.log" | cut -c1-300)"
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step tests_p2p_tp 600 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_tp.py -q -rs -m gpu
step boundary 200 $TR --nproc-per-node 2 --master-port 29561 tools/bench_boundary.py
step mixed70b_tp2 600 $TR --nproc-per-node 2 --master-port 29562 bench.py --gpus 2 --workload llama3-70b-gptq-mixed --layers 4 --requests 64
step tests_all 1500 python -m pytest tests -q -rs -m gpu --deselect tests/test_gpu_p2p.py --deselect tests/test_gpu_tp.py
step bench_n2 400 $TR --nproc-per-node 2 --master-port 29563 bench.py --gpus 2 --steps 24 --warmup 4 --no-extra
tail -4 gpurun_out/r2c12_tests_p2p_tp.log
grep -h "^boundary" gpurun_out/r2c12_boundary.log
for f in mixed70b_tp2 bench_n2; do grep -h '^{' gpurun_out/r2c12_$f.log | cut -c1-300; grep -h '^{' gpurun_out/r2c12_$f.log | grep -o '"prefill": {[^}]*}'; done
tail -3 gpurun_out/r2c12_tests_all.log
