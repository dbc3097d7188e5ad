#!/bin/bash
# session AH: compute-sanitizer (memcheck + racecheck) on the kernels of session 3: cluster-resident circuit / gradient kernels, 3M gate passes,
# cp.async persistent pass
mkdir -p gpurun_out
O=gpurun_out
for tool in memcheck racecheck; do
timeout -k 10 500 compute-sanitizer --tool $tool --error-exitcode 7 python tools/prof_case.py grad 12 layers=6 reps=1 > $O/r02ah_${tool}_cluster_grad.log 2>&1; echo "$tool cluster grad rc=$?"; tail -n 3 $O/r02ah_${tool}_cluster_grad.log | cut -c1-200
timeout -k 10 500 compute-sanitizer --tool $tool --error-exitcode 7 python tools/prof_case.py pass 13 layers=8 reps=1 small_cluster=1 > $O/r02ah_${tool}_cluster_circuit.log 2>&1; echo "$tool cluster circuit rc=$?"; tail -n 3 $O/r02ah_${tool}_cluster_circuit.log | cut -c1-200
timeout -k 10 500 compute-sanitizer --tool $tool --error-exitcode 7 python tools/prof_case.py pass 16 layers=12 reps=1 > $O/r02ah_${tool}_pass3m.log 2>&1; echo "$tool 3M pass rc=$?"; tail -n 3 $O/r02ah_${tool}_pass3m.log | cut -c1-200
timeout -k 10 500 compute-sanitizer --tool $tool --error-exitcode 7 python tools/prof_case.py pass 16 layers=12 reps=1 pass_async=1 > $O/r02ah_${tool}_async.log 2>&1; echo "$tool async pass rc=$?"; tail -n 3 $O/r02ah_${tool}_async.log | cut -c1-200
done
