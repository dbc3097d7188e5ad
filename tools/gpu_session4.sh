#!/bin/bash
# session 4: validate fused adjoint, time gradients, ncu of the backward pass
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cat gpurun_out/gradients.jsonl | cut -c1-420; tail -3 gpurun_out/gradients.err
for tb in 9 10 11 12; do echo "adjoint tb=$tb"; timeout 300 python tools/bench_gradients.py C5 $tb 2>&1 | cut -c1-200; done
echo "== bench c64"; timeout 600 python bench.py --steps 2 --warmup 2 --dtype c64 --no-cpu-baseline > gpurun_out/bench_c64.json 2> gpurun_out/bench_c64.err; cut -c1-400 gpurun_out/bench_c64.json
echo "== ncu adjoint"
CMD="python tools/bench_gradients.py C5"
$CMD > gpurun_out/plain_adj.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:adjoint_pass -s 0 -c 2 -f -o gpurun_out/prof_adjoint $CMD > gpurun_out/ncu_adj.log 2>&1
echo "ncu adjoint rc=$?"; tail -2 gpurun_out/ncu_adj.log
$CMD > gpurun_out/plain_adj2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_gradient_c5.csv $CMD > gpurun_out/ncu_adj_launches.log 2>&1
echo "ncu launches rc=$?"
