#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma_gate or out_of_place" > $O/r02k_pytest.log 2>&1; echo "rc=$?" >> $O/r02k_pytest.log
tail -n 3 $O/r02k_pytest.log
{
for cap in 1073741824 3 2 1; do for mode in 0 1; do timeout -k 10 300 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=$cap pass_tma=$mode; done; done
for mode in 0 2 1; do timeout -k 10 200 python tools/prof_case.py grad 28 pass_tma=$mode; done
for mode in 0 1; do timeout -k 10 200 python tools/prof_case.py pass 28 layers=4 reps=10 pass_tma=$mode; done
timeout -k 10 300 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=3 pass_tma=1 passlog=$O/r02k_passlog.txt
} > $O/r02k_cases.jsonl 2> $O/r02k_cases.err
cut -c1-330 $O/r02k_cases.jsonl; tail -n 3 $O/r02k_cases.err
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:tile_pass_tma -s 3 -c 1 -o $O/r02k_prof_tma_pass_full -f python tools/prof_case.py pass 28 layers=20 reps=1 pass_tma=1 > $O/r02k_ncu1.log 2>&1
timeout -k 10 300 $NCU -k regex:tile_pass_tma -s 10 -c 1 -o $O/r02k_prof_tma_pass_cap3 -f python tools/prof_case.py pass 28 layers=20 reps=1 max_gates_per_pass=3 pass_tma=1 > $O/r02k_ncu2.log 2>&1
tail -n 2 $O/r02k_ncu*.log
