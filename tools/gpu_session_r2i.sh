#!/bin/bash
# round 2, call I: expectation kernel with 16 warps on one tile
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "measure" > $O/r02i_pytest_measure.log 2>&1; echo "rc=$?" >> $O/r02i_pytest_measure.log
tail -n 3 $O/r02i_pytest_measure.log
{
timeout -k 10 200 python tools/prof_case.py measure 30; timeout -k 10 200 python tools/prof_case.py measure32 30
timeout -k 10 200 python tools/prof_case.py measure 28; timeout -k 10 200 python tools/prof_case.py measure32 28
timeout -k 10 200 python tools/prof_case.py measure 24; timeout -k 10 200 python tools/prof_case.py measure32 24
timeout -k 10 200 python tools/prof_case.py grad 28; timeout -k 10 200 python tools/prof_case.py grad32 28
} > $O/r02i_cases.jsonl 2> $O/r02i_cases.err
cut -c1-330 $O/r02i_cases.jsonl; tail -n 3 $O/r02i_cases.err
timeout -k 10 300 python tools/bench_measure_quick.py 30 > $O/r02i_measure_30q.txt 2>&1; cat $O/r02i_measure_30q.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:flip_measure -s 1 -c 1 -o $O/r02i_prof_flip_measure -f python tools/prof_case.py measure 28 reps=1 > $O/r02i_ncu3.log 2>&1
timeout -k 10 300 $NCU -k regex:flip_measure -s 1 -c 1 -o $O/r02i_prof_flip_measure_c64 -f python tools/prof_case.py measure32 28 reps=1 > $O/r02i_ncu3b.log 2>&1
tail -n 2 $O/r02i_ncu*.log
timeout -k 10 1200 python -m pytest tests -q -m gpu > $O/r02i_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02i_pytest_gpu.log
tail -n 4 $O/r02i_pytest_gpu.log
