#!/bin/bash
# session AM: cluster gradient kernel with tensor-core outer products in the local gates
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -x -m gpu -k "cluster or c1 or C1" > $O/r02am_pytest.log 2>&1; echo "rc=$?" >> $O/r02am_pytest.log; tail -n 12 $O/r02am_pytest.log | cut -c1-220
{
timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=30
timeout -k 10 200 python tools/prof_case.py grad 10 layers=20 reps=30
timeout -k 10 200 python tools/prof_case.py grad 8 layers=8 reps=30
timeout -k 10 200 python tools/prof_case.py grad32 12 layers=20 reps=30
} > $O/r02am_cases.jsonl 2> $O/r02am_cases.err
cut -c1-200 $O/r02am_cases.jsonl; tail -n 3 $O/r02am_cases.err
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cluster_gradient" -c 2 --csv --log-file $O/r02am_ncu.csv python tools/prof_case.py grad 12 layers=20 reps=1 > /dev/null 2>&1
tail -n 2 $O/r02am_ncu.csv | awk -F'","' '{print $NF}'
