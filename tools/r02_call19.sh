#!/bin/bash
# two GPUs: the static-table broadcast through NCCL inside the library, one process per GPU
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nccl_broadcast_check.py > gpurun_out/nccl_check.log 2>&1; echo "direct rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/nccl_check.log | tail -12
timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q --timeout 600 2>&1 | tail -3
