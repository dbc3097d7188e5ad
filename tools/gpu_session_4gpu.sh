#!/bin/bash
# 4-GPU verification session (gpurun --gpus 4): multi-GPU parity tests on 4 ranks, sharded bench at 32 q (fused swaps) with its parity_check
mkdir -p gpurun_out
O=gpurun_out
QCB_DIST_NPROC=4 timeout -k 10 1200 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > $O/r02an_4gpu_pytest_dist.log 2>&1; echo "dist rc=$?"; tail -n 6 $O/r02an_4gpu_pytest_dist.log | cut -c1-220
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655"
timeout -k 10 900 $TR bench.py --gpus 4 --nqubits 32 --steps 2 --warmup 1 > $O/r02an_4gpu_bench_32q.json 2> $O/r02an_4gpu_bench_32q.err; echo "bench rc=$?"; cut -c1-300 $O/r02an_4gpu_bench_32q.json; grep -v "^W\|^\*\|OMP_NUM" $O/r02an_4gpu_bench_32q.err | tail -n 5
