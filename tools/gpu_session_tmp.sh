#!/bin/bash
# session AK: cluster kernels, quad-owner form of the gates on the rank bits
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -x -m gpu -k "cluster or c1 or C1" > $O/r02ak_pytest.log 2>&1; echo "rc=$?" >> $O/r02ak_pytest.log; tail -n 3 $O/r02ak_pytest.log | cut -c1-220
{
timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=30
timeout -k 10 200 python tools/prof_case.py grad 10 layers=20 reps=30
timeout -k 10 200 python tools/prof_case.py pass 12 layers=20 reps=50
timeout -k 10 200 python tools/prof_case.py pass 14 layers=20 reps=50 small_cluster=1
timeout -k 10 200 python tools/prof_case.py pass 15 layers=20 reps=50 small_cluster=1
} > $O/r02ak_cases.jsonl 2> $O/r02ak_cases.err
cut -c1-200 $O/r02ak_cases.jsonl; tail -n 3 $O/r02ak_cases.err
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cluster_gradient" -c 2 --csv --log-file $O/r02ak_ncu.csv python tools/prof_case.py grad 12 layers=20 reps=1 > /dev/null 2>&1
tail -n 2 $O/r02ak_ncu.csv | awk -F'","' '{print $NF}'
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cluster_circuit" -c 2 --csv --log-file $O/r02ak_ncu2.csv python tools/prof_case.py pass 12 layers=20 reps=1 > /dev/null 2>&1
tail -n 2 $O/r02ak_ncu2.csv | awk -F'","' '{print $NF}'
