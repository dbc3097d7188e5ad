#!/bin/bash
# session O: form 3 + straight-line full diamonds -- timings, parity, bench
mkdir -p gpurun_out
O=gpurun_out
{
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=3
timeout -k 10 200 python tools/prof_case.py pass32 30 layers=40
timeout -k 10 200 python tools/prof_case.py grad 28
timeout -k 10 200 python tools/prof_case.py grad32 28
} > $O/r02o_cases.jsonl 2> $O/r02o_cases.err
cut -c1-200 $O/r02o_cases.jsonl; tail -n 3 $O/r02o_cases.err
timeout -k 10 1200 python -m pytest tests -q -x -m gpu > $O/r02o_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02o_pytest_gpu.log
tail -n 3 $O/r02o_pytest_gpu.log
timeout -k 10 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $O/r02o_bench.json 2> $O/r02o_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r02o_bench.json; tail -n 3 $O/r02o_bench.err
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tile_pass_kernel -s 8 -c 1 -o $O/r02o_prof_tile_pass -f python tools/prof_case.py pass 30 layers=40 reps=1 > $O/r02o_ncu.log 2>&1; echo "ncu rc=$?"
