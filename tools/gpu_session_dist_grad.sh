#!/bin/bash
# usage (gpurun --gpus N): tools/gpu_session_dist_grad.sh N [NLOC]  -- multi-GPU parity tests (incl. sharded gradients) + sharded gradient timing
N=${1:-2}; NLOC=${2:-28}
mkdir -p gpurun_out
echo "== dist tests"; QCB_DIST_NPROC=$N timeout 1200 python -m pytest tests/test_gpu_dist.py -x -q -rP -m gpu > gpurun_out/pytest_dist_$N.log 2>&1; echo "rc=$?"; grep -a "PASS\|FAIL\|passed\|failed\|Error\|error" gpurun_out/pytest_dist_$N.log | tail -40 | cut -c1-220
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29677"
echo "== sharded gradient timing"; timeout 900 $TR tools/bench_gradients_sharded.py $NLOC > gpurun_out/gradients_sharded_$N.json 2> gpurun_out/gradients_sharded_$N.err; echo "rc=$?"; tail -1 gpurun_out/gradients_sharded_$N.json | cut -c1-600; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/gradients_sharded_$N.err | tail -5
