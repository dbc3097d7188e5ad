#!/bin/bash
# Single-GPU verification call of round 1, session 2: smoke, all GPU parity tests, gradient configs (C1, C2, C5) with
# the scalar / tensor-core backward pass, expectation-value timings, default bench + reference arm.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-330 gpurun_out/gradients.jsonl
ADJ_MMA=0 REPS=3 timeout 600 python tools/bench_gradients.py C5 >> gpurun_out/gradients.jsonl 2>> gpurun_out/gradients.err; tail -1 gpurun_out/gradients.jsonl | cut -c1-330
echo "== measure"; timeout 600 python tools/bench_measure.py 30 > gpurun_out/measure_30q.txt 2>&1; cat gpurun_out/measure_30q.txt
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
