#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== examples"; timeout 300 python examples/test_gpu.py 2>&1 | tail -4; timeout 300 python examples/brickwork_unitary.py 12 8 20 2>&1 | tail -5
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('hbm_bound_regime'), d['cpu_baseline'])"; tail -3 gpurun_out/bench.err
for tb in 11 12; do echo "== bench c64 tb=$tb"; timeout 600 python bench.py --steps 2 --warmup 2 --dtype c64 --no-cpu-baseline --tile-bits $tb > gpurun_out/bench_c64_$tb.json 2> gpurun_out/bench_c64.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c64_$tb.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['fma_roof']['achieved_tflops'], d['roofline'].get('hbm_bound_regime'))"; done
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-200 gpurun_out/gradients.jsonl
