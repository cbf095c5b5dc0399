#!/bin/bash
# Round 2, last GPU call (2 GPUs): full GPU suite and the TP2 bench line with the split-KV threshold of one CTA per SM.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -rs -m gpu > gpurun_out/r2c19_tests_all.log 2>&1; echo "[tests] exit $?: $(tail -n 1 gpurun_out/r2c19_tests_all.log)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 24 --warmup 4 --no-extra > gpurun_out/r2c19_bench_n2.log 2>&1
echo "[bench_n2] exit $?"; grep -h '^{' gpurun_out/r2c19_bench_n2.log | cut -c1-400
