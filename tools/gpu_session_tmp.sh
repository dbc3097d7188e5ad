#!/bin/bash
# session W: persistent gate pass with asynchronous tile loads (pass_async) against the one-tile-per-CTA form
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "async or 3m_form" > $O/r02w_pytest.log 2>&1; echo "rc=$?" >> $O/r02w_pytest.log; tail -n 4 $O/r02w_pytest.log
rm -f $O/r02w_passlog.txt
{
for a in 0 1; do
timeout -k 10 200 python tools/prof_case.py pass 30 layers=100 pass_async=$a
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=3 pass_async=$a
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=1 pass_async=$a
timeout -k 10 200 python tools/prof_case.py pass 28 layers=4 pass_async=$a
timeout -k 10 200 python tools/prof_case.py pass32 30 layers=40 pass_async=$a
timeout -k 10 200 python tools/prof_case.py grad 28 pass_async=$a
timeout -k 10 200 python tools/prof_case.py grad 20 reps=20 pass_async=$a
timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=20 pass_async=$a
done
timeout -k 10 200 python tools/prof_case.py pass 30 layers=100 pass_async=1 passlog=$O/r02w_passlog.txt
} > $O/r02w_cases.jsonl 2> $O/r02w_cases.err
cut -c1-220 $O/r02w_cases.jsonl; tail -n 3 $O/r02w_cases.err
