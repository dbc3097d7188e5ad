#!/bin/bash
# session Z: kernel-only durations of the cluster circuit kernel (ncu launch list)
mkdir -p gpurun_out
O=gpurun_out
for n in 9 10 12 14 15; do
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cluster_circuit -c 4 --csv --log-file $O/r02z_ncu_n$n.csv python tools/prof_case.py pass $n layers=20 reps=2 small_cluster=1 > /dev/null 2>&1
tail -n 2 $O/r02z_ncu_n$n.csv | cut -c1-260
done
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02z_ncu_tile12.csv python tools/prof_case.py pass 12 layers=20 reps=2 small_cluster=0 > /dev/null 2>&1
tail -n 6 $O/r02z_ncu_tile12.csv | cut -c1-260
