#!/bin/bash
# session AE: 128 vs 256 threads per CTA in the small-state kernels
mkdir -p gpurun_out
O=gpurun_out
V=quantumcircuits.jl_b200/variants
{
for lib in "" $V/libqcb200_t128.so; do
echo "# lib=$lib"
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=30
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 12 layers=20 reps=50
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py grad 10 layers=20 reps=30
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py grad 8 layers=8 reps=30
done
} > $O/r02ae_cases.jsonl 2> $O/r02ae_cases.err
cut -c1-200 $O/r02ae_cases.jsonl; tail -n 3 $O/r02ae_cases.err
QCB200_LIB=$V/libqcb200_t128.so timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "cluster" 2>&1 | tail -n 2
