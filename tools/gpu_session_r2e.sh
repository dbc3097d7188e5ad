#!/bin/bash
# round 2, call E: expectation kernel without a producer warp, deferred reduction of the backward sweep, 5 CTAs per SM
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "measure or gradient or optimise or polar" > $O/r02e_pytest_new.log 2>&1; echo "rc=$?" >> $O/r02e_pytest_new.log
tail -n 3 $O/r02e_pytest_new.log
{
timeout -k 10 200 python tools/prof_case.py measure 30; timeout -k 10 200 python tools/prof_case.py measure32 30
timeout -k 10 200 python tools/prof_case.py measure 28; timeout -k 10 200 python tools/prof_case.py measure32 28
timeout -k 10 200 python tools/prof_case.py measure 30 measure_tma=0; timeout -k 10 200 python tools/prof_case.py measure32 30 measure_tma=0
timeout -k 10 200 python tools/prof_case.py grad32 28
timeout -k 10 200 python tools/prof_case.py grad 28
timeout -k 10 200 python tools/prof_case.py grad 28 adjoint_defer=0
for d in 1 0; do timeout -k 10 200 python tools/prof_case.py grad 20 reps=20 adjoint_defer=$d; timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=20 adjoint_defer=$d; done
} > $O/r02e_cases.jsonl 2> $O/r02e_cases.err
cut -c1-330 $O/r02e_cases.jsonl; tail -n 3 $O/r02e_cases.err
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:flip_measure -s 1 -c 1 -o $O/r02e_prof_flip_measure -f python tools/prof_case.py measure 28 reps=1 > $O/r02e_ncu3.log 2>&1
timeout -k 10 300 $NCU -k regex:flip_measure -s 1 -c 1 -o $O/r02e_prof_flip_measure_c64 -f python tools/prof_case.py measure32 28 reps=1 > $O/r02e_ncu3b.log 2>&1
timeout -k 10 300 $NCU -k regex:adjoint_pass_f32 -s 2 -c 1 -o $O/r02e_prof_adjoint_f32 -f python tools/prof_case.py grad32 28 reps=1 > $O/r02e_ncu4.log 2>&1
tail -n 2 $O/r02e_ncu*.log
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $O/r02e_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02e_pytest_gpu.log
tail -n 3 $O/r02e_pytest_gpu.log
