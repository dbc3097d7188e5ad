#!/bin/bash
# session Q: per-pass choice of the ComplexF64 arithmetic (f64_3m_min_gates) -- A/B against the previous commit on one box, parity, bench
mkdir -p gpurun_out
O=gpurun_out
V=quantumcircuits.jl_b200/variants
rm -f $O/r02q_passlog.txt
{
for lib in $V/libqcb200_old.so ""; do
echo "# lib=$lib"
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 30 layers=40
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=3
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=1
QCB200_LIB=$lib timeout -k 10 200 python tools/prof_case.py pass 28 layers=4
done
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 f64_3m_min_gates=0
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 f64_3m_min_gates=9
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 f64_3m_min_gates=1000
timeout -k 10 200 python tools/prof_case.py pass 30 layers=100 passlog=$O/r02q_passlog.txt
} > $O/r02q_cases.jsonl 2> $O/r02q_cases.err
cut -c1-200 $O/r02q_cases.jsonl; tail -n 3 $O/r02q_cases.err
timeout -k 10 1200 python -m pytest tests -q -x -m gpu > $O/r02q_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02q_pytest_gpu.log
tail -n 3 $O/r02q_pytest_gpu.log
timeout -k 10 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $O/r02q_bench.json 2> $O/r02q_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r02q_bench.json; tail -n 3 $O/r02q_bench.err
