#!/bin/bash
# round 2, call C: ComplexF32 backward pass (register windows + transpose-reduction), two consumer groups in the expectation kernel,
# per-pass timing log of the gate passes, ncu of the TMA gate pass
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "gradient or optimise or measure or polar" > $O/r02c_pytest_new.log 2>&1; echo "rc=$?" >> $O/r02c_pytest_new.log
tail -n 3 $O/r02c_pytest_new.log
{
for v in 1 0; do timeout -k 10 200 python tools/prof_case.py grad32 28 adjoint_f32=$v; done
timeout -k 10 200 python tools/prof_case.py grad32 28 adjoint_f32=1 adjoint_tile_bits=10
timeout -k 10 200 python tools/prof_case.py grad32 28 adjoint_f32=1 adjoint_tile_bits=12
timeout -k 10 200 python tools/prof_case.py grad 28
timeout -k 10 200 python tools/prof_case.py measure 30; timeout -k 10 200 python tools/prof_case.py measure32 30
timeout -k 10 200 python tools/prof_case.py measure 28; timeout -k 10 200 python tools/prof_case.py measure32 28
timeout -k 10 300 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=3 pass_tma=0 passlog=$O/r02c_passlog.txt
timeout -k 10 300 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=3 pass_tma=1 passlog=$O/r02c_passlog.txt
timeout -k 10 300 python tools/prof_case.py pass 30 layers=40 pass_tma=0 passlog=$O/r02c_passlog.txt
timeout -k 10 300 python tools/prof_case.py pass 30 layers=40 pass_tma=1 passlog=$O/r02c_passlog.txt
} > $O/r02c_cases.jsonl 2> $O/r02c_cases.err
cut -c1-330 $O/r02c_cases.jsonl
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:tile_pass_tma -s 3 -c 1 -o $O/r02c_prof_tma_pass_full -f python tools/prof_case.py pass 28 layers=20 reps=1 pass_tma=1 > $O/r02c_ncu1.log 2>&1
timeout -k 10 300 $NCU -k regex:tile_pass_tma -s 10 -c 1 -o $O/r02c_prof_tma_pass_cap3 -f python tools/prof_case.py pass 28 layers=20 reps=1 max_gates_per_pass=3 pass_tma=1 > $O/r02c_ncu2.log 2>&1
timeout -k 10 300 $NCU -k regex:tile_pass_kernel -s 3 -c 1 -o $O/r02c_prof_ldg_pass_full -f python tools/prof_case.py pass 28 layers=20 reps=1 pass_tma=0 > $O/r02c_ncu5.log 2>&1
timeout -k 10 300 $NCU -k regex:flip_measure -s 1 -c 1 -o $O/r02c_prof_flip_measure -f python tools/prof_case.py measure 28 reps=1 > $O/r02c_ncu3.log 2>&1
timeout -k 10 300 $NCU -k regex:adjoint_pass_f32 -s 2 -c 1 -o $O/r02c_prof_adjoint_f32 -f python tools/prof_case.py grad32 28 reps=1 > $O/r02c_ncu4.log 2>&1
tail -n 2 $O/r02c_ncu*.log
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $O/r02c_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02c_pytest_gpu.log
tail -n 3 $O/r02c_pytest_gpu.log
