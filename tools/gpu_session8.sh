#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
for tb in 9 10 11 12; do echo "adjoint tb=$tb"; timeout 300 python tools/bench_gradients.py C5 $tb 2>&1 | cut -c1-200; done
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-260 gpurun_out/gradients.jsonl
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_reference.json
