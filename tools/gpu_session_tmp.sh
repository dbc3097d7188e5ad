#!/bin/bash
# session AS: partner loads in flight per thread in the H psi kernels (batched gathers), A/B builds
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "tiled_hamiltonian" > $O/r02as_pytest.log 2>&1; echo "rc=$?" >> $O/r02as_pytest.log; tail -n 3 $O/r02as_pytest.log | cut -c1-220
run() { L=$1; shift; QCB200_LIB=$PWD/quantumcircuits.jl_b200/libqcb200$L.so timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"k_apply_ham" -c 1 --csv --log-file $O/r02as_tmp.csv python tools/prof_case.py $1 28 reps=1 ${@:2} > /dev/null 2>&1
echo "lib$L $@ :" $(grep -E "k_apply_ham" $O/r02as_tmp.csv | awk -F'","' '{print substr($5,1,30), $(NF-2), $NF}' | tr -d '"' | tr '\n' ' '); }
{
run _old grad32
run "" grad32
run _b16 grad32
run _b4 grad32
run _old grad
run _b16 grad
run "" grad ham_tile=2
run _old grad32 ham_tile=0
run _b16 grad32 ham_tile=0
} > $O/r02as_ham_batch.txt 2>&1
cat $O/r02as_ham_batch.txt
{
QCB200_LIB=$PWD/quantumcircuits.jl_b200/libqcb200_old.so timeout -k 10 200 python tools/prof_case.py grad32 28 reps=5
timeout -k 10 200 python tools/prof_case.py grad32 28 reps=5
QCB200_LIB=$PWD/quantumcircuits.jl_b200/libqcb200_old.so timeout -k 10 200 python tools/prof_case.py grad32 24 reps=10
timeout -k 10 200 python tools/prof_case.py grad32 24 reps=10
} > $O/r02as_cases.jsonl 2> $O/r02as_cases.err
cut -c1-200 $O/r02as_cases.jsonl; tail -n 3 $O/r02as_cases.err
