#!/bin/bash
# multi-GPU session: usage tools/gpu_session_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/nvsmi_multi_$N.txt 2>&1; free -g >> gpurun_out/nvsmi_multi_$N.txt; nproc >> gpurun_out/nvsmi_multi_$N.txt
nvidia-smi topo -m >> gpurun_out/nvsmi_multi_$N.txt 2>&1
echo "== dist tests"; QCB_DIST_NPROC=$N timeout 1500 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/pytest_dist_$N.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_dist_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
echo "== bench 34q"; timeout 1500 $TR bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_multi_${N}_34q.json 2> gpurun_out/bench_multi_${N}_34q.err; echo "rc=$?"; cat gpurun_out/bench_multi_${N}_34q.json; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/bench_multi_${N}_34q.err | tail -15
