#!/bin/bash
# round 2, call B: ComplexF32 backward pass on TF32-split MMAs, L2 promotion of the strided tensor maps, probes, ncu captures
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 200 ./tools/fp64_peak > $O/r02b_peaks.json 2>&1
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "gradient or optimise or out_of_place_gate or measure_vs or hamiltonian_layer" > $O/r02b_pytest_new.log 2>&1; echo "rc=$?" >> $O/r02b_pytest_new.log
tail -3 $O/r02b_pytest_new.log
{
for tf in 1 0; do timeout -k 10 200 python tools/prof_case.py grad32 28 adjoint_tf32=$tf; done
for tf in 1 0; do timeout -k 10 200 python tools/prof_case.py grad32 28 adjoint_tf32=$tf adjoint_tile_bits=10; done
timeout -k 10 200 python tools/prof_case.py grad 28
for pr in 1 2 0; do timeout -k 10 200 python tools/prof_case.py measure 30 tma_l2_promotion=$pr; timeout -k 10 200 python tools/prof_case.py measure32 30 tma_l2_promotion=$pr; done
for cap in 1073741824 3 2; do for mode in "pass_tma=0" "pass_tma=1 tma_l2_promotion=1" "pass_tma=1 tma_l2_promotion=2"; do
  timeout -k 10 300 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=$cap $mode; done; done
} > $O/r02b_cases.jsonl 2> $O/r02b_cases.err
cat $O/r02b_cases.jsonl | cut -c1-330
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:tile_pass_tma -s 3 -c 1 -o $O/r02b_prof_tma_pass_full -f python tools/prof_case.py pass 28 layers=20 reps=1 > $O/r02b_ncu1.log 2>&1
timeout -k 10 300 $NCU -k regex:tile_pass_tma -s 10 -c 1 -o $O/r02b_prof_tma_pass_cap3 -f python tools/prof_case.py pass 28 layers=20 reps=1 max_gates_per_pass=3 > $O/r02b_ncu2.log 2>&1
timeout -k 10 300 $NCU -k regex:flip_measure -s 1 -c 1 -o $O/r02b_prof_flip_measure -f python tools/prof_case.py measure 28 reps=1 > $O/r02b_ncu3.log 2>&1
timeout -k 10 300 $NCU -k regex:adjoint_pass_f32 -s 2 -c 1 -o $O/r02b_prof_adjoint_f32 -f python tools/prof_case.py grad32 28 reps=1 > $O/r02b_ncu4.log 2>&1
timeout -k 10 300 $NCU -k regex:tile_pass_kernel -s 10 -c 1 -o $O/r02b_prof_ldg_pass_cap3 -f python tools/prof_case.py pass 28 layers=20 reps=1 max_gates_per_pass=3 pass_tma=0 > $O/r02b_ncu5.log 2>&1
tail -2 $O/r02b_ncu*.log
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $O/r02b_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02b_pytest_gpu.log
tail -3 $O/r02b_pytest_gpu.log
ls -la $O | grep r02b
