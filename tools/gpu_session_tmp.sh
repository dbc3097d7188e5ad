#!/bin/bash
# session T: code-size variants of the 3M kernel on one box (V2 = straight-line full diamond + generic ordinary path, V1 = one path of four
# conditional 3M gates, V4 = one path with three gate bodies)
mkdir -p gpurun_out
O=gpurun_out
V=quantumcircuits.jl_b200/variants
rm -f $O/r02t_passlog_v4.txt
{
for lib in "" $V/libqcb200_v1.so $V/libqcb200_v4.so; do
echo "# lib=$lib"
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 30 layers=100
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 30 layers=100 f64_3m_min_gates=9
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 28 layers=4
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 28 layers=4 f64_3m_min_gates=13
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py grad 28
done
QCB200_LIB=$V/libqcb200_v4.so timeout -k 10 200 python tools/prof_case.py pass 30 layers=100 passlog=$O/r02t_passlog_v4.txt
} > $O/r02t_cases.jsonl 2> $O/r02t_cases.err
cut -c1-200 $O/r02t_cases.jsonl; tail -n 3 $O/r02t_cases.err
