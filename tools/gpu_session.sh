#!/bin/bash
# One single-GPU gpurun call: smoke, FP roof probe, GPU parity tests, examples, bench (both dtypes, reference arm),
# gradient / measure timings, ncu launch list + full captures of the three hot kernels.
# usage: gpurun -- 'bash tools/gpu_session.sh [quick]'
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1; free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== fp roofs"; timeout 120 ./tools/fp64_peak > gpurun_out/fp64_peak.json 2>&1; cat gpurun_out/fp64_peak.json
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== examples"; timeout 300 python examples/test_gpu.py 2>&1 | tail -2; timeout 300 python examples/brickwork_unitary.py 12 8 20 2>&1 | tail -2
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
echo "== bench c64"; timeout 600 python bench.py --steps 2 --warmup 2 --dtype c64 --no-cpu-baseline > gpurun_out/bench_c64.json 2> gpurun_out/bench_c64.err; cut -c1-300 gpurun_out/bench_c64.json
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-260 gpurun_out/gradients.jsonl
if [ "$1" != "quick" ]; then
echo "== ncu"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 30 -c 3 -f -o gpurun_out/prof_tile $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu tile rc=$?"
CMD2="python tools/bench_gradients.py C5"
$CMD2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"measure_tile|adjoint_pass" -s 0 -c 5 -f -o gpurun_out/prof_grad $CMD2 > gpurun_out/ncu_grad.log 2>&1
echo "ncu grad rc=$?"
fi
if [ "$2" == "big" ]; then
echo "== 34q ComplexF32 on one GPU (128 GiB state)"; timeout 900 python bench.py --nqubits 34 --dtype c64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_34q_c64_1gpu.json 2> gpurun_out/bench_34q_c64_1gpu.err; cut -c1-300 gpurun_out/bench_34q_c64_1gpu.json; tail -2 gpurun_out/bench_34q_c64_1gpu.err
fi
