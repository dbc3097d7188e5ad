#!/bin/bash
# session AQ: H psi gather (ComplexF64): CTAs per SM and traversal order against the L2 misses of the top-bit partner streams
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "tiled_hamiltonian" > $O/r02aq_pytest.log 2>&1; echo "rc=$?" >> $O/r02aq_pytest.log; tail -n 6 $O/r02aq_pytest.log | cut -c1-220
run() { timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"k_apply_ham" -c 1 --csv --log-file $O/r02aq_tmp.csv python tools/prof_case.py $1 28 reps=1 ${@:2} > /dev/null 2>&1
echo "$@ :" $(grep -E "k_apply_ham" $O/r02aq_tmp.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tr '\n' ' '); }
{
run grad ham_ctas=8
run grad ham_ctas=6
run grad ham_ctas=4
run grad ham_ctas=3
run grad ham_ctas=2
run grad ham_order=712
run grad ham_order=912
run grad ham_order=714
run grad ham_order=514
run grad ham_order=716
run grad ham_order=1010
run grad ham_order=712 ham_ctas=4
run grad32 ham_tile=0 ham_order=712
run grad32 ham_tile=0 ham_order=714
run grad32 ham_tile=1 ham_ctas=6
run grad32 ham_tile=1 ham_ctas=4
} > $O/r02aq_ham_variants.txt 2>&1
cat $O/r02aq_ham_variants.txt
