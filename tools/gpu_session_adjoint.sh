#!/bin/bash
# One single-GPU gpurun call for the tensor-core backward pass (adjoint_mma.cuh): parity tests, racecheck, timings of
# the scalar vs tensor-core pass (C5: 28 q), ncu launch list + full capture, default bench.
# usage: gpurun -- 'bash tools/gpu_session_adjoint.sh'
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-400
echo "== racecheck (12 q gradient, both passes)"
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_gradient.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log | cut -c1-300
echo "== gradients C5"
for v in "0 0" "1 4" "1 8"; do set -- $v
  REPS=3 ADJ_MMA=$1 ADJ_WARPS=$2 timeout 600 python tools/bench_gradients.py C5 >> gpurun_out/gradients_adj.jsonl 2>> gpurun_out/gradients_adj.err
done
REPS=3 ADJ_MMA=1 ADJ_WARPS=8 timeout 600 python tools/bench_gradients.py C5 11 >> gpurun_out/gradients_adj.jsonl 2>> gpurun_out/gradients_adj.err
REPS=3 ADJ_MMA=1 ADJ_WARPS=4 timeout 600 python tools/bench_gradients.py C5 11 >> gpurun_out/gradients_adj.jsonl 2>> gpurun_out/gradients_adj.err
cut -c1-330 gpurun_out/gradients_adj.jsonl
echo "== all gradient configs"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-260 gpurun_out/gradients.jsonl
echo "== ncu"
CMD2="python tools/bench_gradients.py C5"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_gradient_c5.csv $CMD2 > gpurun_out/ncu_grad_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"adjoint_pass" -s 6 -c 3 -f -o gpurun_out/prof_adjoint_mma $CMD2 > gpurun_out/ncu_adj.log 2>&1; echo "ncu adjoint rc=$?"
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
