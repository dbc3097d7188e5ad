#!/bin/bash
# Quick single-GPU iteration on the tensor-core backward pass: gradient parity tests, C5 timings of the variants, ncu.
mkdir -p gpurun_out
echo "== pytest gpu (gradients)"; timeout 900 python -m pytest tests -m gpu -q -k "gradient or optimise or autograd" > gpurun_out/pytest_gpu_grad.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_grad.log | cut -c1-300
rm -f gpurun_out/gradients_adj.jsonl
for v in "0 0 0" "1 4 0" "1 8 0" "1 4 11" "1 8 11"; do set -- $v
  REPS=3 ADJ_MMA=$1 ADJ_WARPS=$2 timeout 600 python tools/bench_gradients.py C5 $3 >> gpurun_out/gradients_adj.jsonl 2>> gpurun_out/gradients_adj.err
done
cut -c1-230 gpurun_out/gradients_adj.jsonl
CMD2="python tools/bench_gradients.py C5"
ncu --set full --clock-control none --import-source on -k regex:"adjoint_pass" -s 7 -c 2 -f -o gpurun_out/prof_adjoint_mma $CMD2 > gpurun_out/ncu_adj.log 2>&1; echo "ncu adjoint rc=$?"
