#!/bin/bash
mkdir -p gpurun_out
QCB200_LIB=quantumcircuits.jl_b200/variants/libqcb200_timing.so timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=2 2>&1 | grep cycles | tail -n 2
QCB200_LIB=quantumcircuits.jl_b200/variants/libqcb200_timing.so timeout -k 10 200 python tools/prof_case.py grad 9 layers=20 reps=2 2>&1 | grep cycles | tail -n 1
