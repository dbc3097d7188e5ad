#!/bin/bash
# 2-GPU verification session (gpurun --gpus 2): multi-GPU parity tests (incl. shards without a second buffer: in-place peer swaps), sharded bench at
# 31 q, the 34 q bench of the driver's N = 2 run with peer swaps and with staged NCCL swaps, sharded gradients
mkdir -p gpurun_out
O=gpurun_out
QCB_DIST_NPROC=2 timeout -k 10 1500 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > $O/r02ag_2gpu_pytest_dist_2.log 2>&1; echo "dist rc=$?"; tail -n 25 $O/r02ag_2gpu_pytest_dist_2.log | cut -c1-220
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout -k 10 900 $TR bench.py --gpus 2 --nqubits 31 --steps 2 --warmup 1 > $O/r02ag_2gpu_bench_2gpu_31q.json 2> $O/r02ag_2gpu_bench_2gpu_31q.err; echo "bench rc=$?"; cut -c1-300 $O/r02ag_2gpu_bench_2gpu_31q.json; grep -v "^W\|^\*\|OMP_NUM" $O/r02ag_2gpu_bench_2gpu_31q.err | tail -n 5
timeout -k 10 1500 $TR bench.py --gpus 2 --steps 2 --warmup 1 > $O/r02ag_2gpu_bench_2gpu_34q.json 2> $O/r02ag_2gpu_bench_2gpu_34q.err; echo "bench34 rc=$?"; cut -c1-300 $O/r02ag_2gpu_bench_2gpu_34q.json; grep -v "^W\|^\*\|OMP_NUM" $O/r02ag_2gpu_bench_2gpu_34q.err | tail -n 5
QCB_PEER_SWAP=0 timeout -k 10 1500 $TR bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e > $O/r02ag_2gpu_bench_2gpu_34q_nccl.json 2> $O/r02ag_2gpu_bench_2gpu_34q_nccl.err; echo "bench34 nccl rc=$?"; cut -c1-300 $O/r02ag_2gpu_bench_2gpu_34q_nccl.json
