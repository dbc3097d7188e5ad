#!/bin/bash
# session AP: tiled H psi sweep, tile shapes (ham_tile 1 / 2 / 3) and CTAs per SM
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "tiled_hamiltonian" > $O/r02ap_pytest.log 2>&1; echo "rc=$?" >> $O/r02ap_pytest.log; tail -n 6 $O/r02ap_pytest.log | cut -c1-220
{
for ht in 0 1 2 3; do
timeout -k 10 200 python tools/prof_case.py grad 28 reps=5 ham_tile=$ht
timeout -k 10 200 python tools/prof_case.py grad32 28 reps=5 ham_tile=$ht
done
timeout -k 10 200 python tools/prof_case.py grad 28 reps=5 ham_tile=2 ham_ctas=3
timeout -k 10 200 python tools/prof_case.py grad32 28 reps=5 ham_tile=2 ham_ctas=3
timeout -k 10 200 python tools/prof_case.py grad 28 reps=5 ham_tile=3 ham_ctas=3
timeout -k 10 200 python tools/prof_case.py grad32 28 reps=5 ham_tile=1 ham_ctas=3
timeout -k 10 200 python tools/prof_case.py grad32 24 reps=10 ham_tile=1
timeout -k 10 200 python tools/prof_case.py grad32 24 reps=10 ham_tile=2
timeout -k 10 200 python tools/prof_case.py grad 24 reps=10 ham_tile=0
timeout -k 10 200 python tools/prof_case.py grad 24 reps=10 ham_tile=2
} > $O/r02ap_cases.jsonl 2> $O/r02ap_cases.err
cut -c1-200 $O/r02ap_cases.jsonl; tail -n 3 $O/r02ap_cases.err
for ht in 2 3; do
for cs in grad grad32; do
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:"k_apply_ham" -c 1 --csv --log-file $O/r02ap_ncu_${cs}_ham$ht.csv python tools/prof_case.py $cs 28 reps=1 ham_tile=$ht > /dev/null 2>&1
grep -E "k_apply_ham" $O/r02ap_ncu_${cs}_ham$ht.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $NF}' | cut -c1-160
done
done
