#!/bin/bash
# End-of-round single-GPU verification: smoke, all GPU parity tests, examples, gradient configs, bench (ours + reference arm),
# ncu launch list of one C5 gradient (kernel shares of the step).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== examples"; for e in test_gpu.py test_gradients.py; do timeout 300 python examples/$e 2>&1 | tail -1; done; timeout 300 python examples/brickwork_unitary.py 12 8 20 2>&1 | tail -1
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-280 gpurun_out/gradients.jsonl
REPS=3 timeout 600 python tools/bench_gradients.py C5 11 >> gpurun_out/gradients.jsonl 2>> gpurun_out/gradients.err; tail -1 gpurun_out/gradients.jsonl | cut -c1-280
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/bench_reference.json
echo "== ncu launch list (C5 gradient)"
CMD2="python tools/bench_gradients.py C5"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_gradient_c5.csv $CMD2 > gpurun_out/ncu_grad_launches.log 2>&1; echo "ncu launches rc=$?"
