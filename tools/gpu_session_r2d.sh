#!/bin/bash
# round 2, call D: expectation kernel with two consumer groups (112 registers), ComplexF32 backward pass without operand broadcasts,
# ComplexF64 backward pass at 5 CTAs per SM, launch lists of the small gradient configs
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "measure or gradients_f32 or polar_optimise_inv" > $O/r02d_pytest_new.log 2>&1; echo "rc=$?" >> $O/r02d_pytest_new.log
tail -n 3 $O/r02d_pytest_new.log
{
timeout -k 10 200 python tools/prof_case.py measure 30; timeout -k 10 200 python tools/prof_case.py measure32 30
timeout -k 10 200 python tools/prof_case.py measure 28; timeout -k 10 200 python tools/prof_case.py measure32 28
for v in 1 0; do timeout -k 10 200 python tools/prof_case.py grad32 28 adjoint_f32=$v; done
timeout -k 10 200 python tools/prof_case.py grad 28
timeout -k 10 200 python tools/prof_case.py grad 28 adjoint_ctas=5
timeout -k 10 200 python tools/prof_case.py grad 20 reps=20
timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=20
timeout -k 10 200 python tools/prof_case.py pass 12 layers=20 reps=20
} > $O/r02d_cases.jsonl 2> $O/r02d_cases.err
cut -c1-330 $O/r02d_cases.jsonl; tail -n 3 $O/r02d_cases.err
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:flip_measure -s 1 -c 1 -o $O/r02d_prof_flip_measure -f python tools/prof_case.py measure 28 reps=1 > $O/r02d_ncu3.log 2>&1
timeout -k 10 300 $NCU -k regex:adjoint_pass_f32 -s 2 -c 1 -o $O/r02d_prof_adjoint_f32 -f python tools/prof_case.py grad32 28 reps=1 > $O/r02d_ncu4.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02d_launches_c2.csv python tools/prof_case.py grad 20 reps=1 > $O/r02d_ncu_c2.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02d_launches_c1.csv python tools/prof_case.py grad 12 layers=20 reps=1 > $O/r02d_ncu_c1.log 2>&1
tail -n 2 $O/r02d_ncu*.log
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $O/r02d_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02d_pytest_gpu.log
tail -n 3 $O/r02d_pytest_gpu.log
