#!/bin/bash
# 2-GPU verification session (gpurun --gpus 2): multi-GPU parity tests, sharded bench at 31 q with its reference arm, sharded gradients
mkdir -p gpurun_out
O=gpurun_out
QCB_DIST_NPROC=2 timeout -k 10 1500 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > $O/r02v_2gpu_pytest_dist_2.log 2>&1; echo "dist rc=$?"; tail -n 25 $O/r02v_2gpu_pytest_dist_2.log | cut -c1-220
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout -k 10 1500 $TR bench.py --gpus 2 --nqubits 31 --steps 2 --warmup 1 > $O/r02v_2gpu_bench_2gpu_31q.json 2> $O/r02v_2gpu_bench_2gpu_31q.err; echo "bench rc=$?"; cut -c1-500 $O/r02v_2gpu_bench_2gpu_31q.json; grep -v "^W\|^\*\|OMP_NUM" $O/r02v_2gpu_bench_2gpu_31q.err | tail -n 5
timeout -k 10 600 $TR bench.py --impl reference --gpus 2 --nqubits 31 --steps 2 --warmup 1 > $O/r02v_2gpu_bench_reference_2gpu.json 2> $O/r02v_2gpu_bench_reference_2gpu.err; cut -c1-300 $O/r02v_2gpu_bench_reference_2gpu.json
timeout -k 10 900 $TR tools/bench_gradients_sharded.py > $O/r02v_2gpu_gradients_sharded_2.json 2> $O/r02v_2gpu_gradients_sharded_2.err; cat $O/r02v_2gpu_gradients_sharded_2.json | cut -c1-400
