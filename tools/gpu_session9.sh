#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (gradients)"; timeout 1800 python -m pytest tests -m gpu -q -k "gradient or golden or optimise or hamiltonian or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
for tb in 10 11; do echo "adjoint tb=$tb"; timeout 300 python tools/bench_gradients.py C5 $tb 2>&1 | cut -c1-200; done
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-260 gpurun_out/gradients.jsonl
echo "== ncu adjoint"
CMD="python tools/bench_gradients.py C5"
$CMD > gpurun_out/plain_adj.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:adjoint_pass -s 0 -c 2 -f -o gpurun_out/prof_adjoint $CMD > gpurun_out/ncu_adj.log 2>&1
echo "ncu adjoint rc=$?"; tail -2 gpurun_out/ncu_adj.log
$CMD > gpurun_out/plain_adj2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_gradient_c5.csv $CMD > gpurun_out/ncu_adj_launches.log 2>&1
echo "ncu launches rc=$?"
