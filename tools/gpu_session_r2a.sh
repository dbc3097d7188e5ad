#!/bin/bash
# round 2, call A: TMA kernels first (bounded), the whole GPU suite, then timing of both tile movers and the expectation value
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_nvsmi.txt 2>&1
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma or right_apply or measure" > gpurun_out/r02a_pytest_tma.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_pytest_tma.log
timeout -k 10 300 python tools/bench_measure_quick.py 30 > gpurun_out/r02a_measure_30q.txt 2>&1; echo "rc=$?" >> gpurun_out/r02a_measure_30q.txt
timeout -k 10 300 python tools/bench_measure_quick.py 28 > gpurun_out/r02a_measure_28q.txt 2>&1; echo "rc=$?" >> gpurun_out/r02a_measure_28q.txt
timeout -k 10 600 python tools/bench_pass_modes.py 30 100 > gpurun_out/r02a_pass_modes_30q.jsonl 2> gpurun_out/r02a_pass_modes_30q.err; echo "rc=$?" >> gpurun_out/r02a_pass_modes_30q.err
timeout -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_tma.log gpurun_out/r02a_pytest_gpu.log
cat gpurun_out/r02a_measure_30q.txt
cat gpurun_out/r02a_pass_modes_30q.jsonl | cut -c1-400
