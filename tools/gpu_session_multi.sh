#!/bin/bash
# usage: tools/gpu_session_multi2.sh N NQ   -- dist tests for N ranks + bench at NQ qubits (fused swaps need the second buffer)
N=${1:-2}; NQ=${2:-31}
mkdir -p gpurun_out
echo "== dist tests"; QCB_DIST_NPROC=$N timeout 1500 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/pytest_dist_$N.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_dist_$N.log | cut -c1-250
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
for fuse in 1 0; do
echo "== bench ${NQ}q fuse=$fuse"; QCB_FUSE_SWAPS=$fuse timeout 1500 $TR bench.py --gpus $N --nqubits $NQ --steps 2 --warmup 1 > gpurun_out/bench_multi_${N}_${NQ}q_fuse$fuse.json 2> gpurun_out/bench_multi_${N}_${NQ}q_fuse$fuse.err; echo "rc=$?"; python -c "
import json
L=open('gpurun_out/bench_multi_${N}_${NQ}q_fuse$fuse.json').read().strip().splitlines(); print(len(L),'stdout lines')
d=json.loads(L[-1]); print(d['value'], d['ms_per_step'], d['config']['sharding'], d['config'].get('swaps_per_circuit'), d.get('e2e'))"; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/bench_multi_${N}_${NQ}q_fuse$fuse.err | tail -5
done
