#!/bin/bash
# Validation + A/B of the peer-memory all-reduce on a 2-GPU box (writes gpurun_out/p2p_*.log):
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/ab_p2p.sh'
set -u
mkdir -p gpurun_out
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_p2p.py "tests/test_gpu_experimental.py::test_switch_keeps_parity[B200_P2P_ALLREDUCE]" \
  -q -m gpu > gpurun_out/p2p_parity.log 2>&1
tail -5 gpurun_out/p2p_parity.log
for sw in 0 1; do
  echo "== bench llama2-7b-fp16 tp2, B200_P2P_ALLREDUCE=$sw" >> gpurun_out/p2p_bench.log
  B200_P2P_ALLREDUCE=$sw timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 >> gpurun_out/p2p_bench.log 2>&1
done
grep -h '"metric"\|^==' gpurun_out/p2p_bench.log | cut -c1-220
