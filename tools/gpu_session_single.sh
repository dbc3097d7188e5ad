#!/bin/bash
# single-GPU verification session: case timings, the whole GPU suite, smoke, the driver's bench commands, ncu evidence of the bench (launch list + full capture)
mkdir -p gpurun_out
O=gpurun_out
{
timeout -k 10 200 python tools/prof_case.py pass 30 layers=100
timeout -k 10 200 python tools/prof_case.py pass 28 layers=4
timeout -k 10 200 python tools/prof_case.py pass 28 layers=4 f64_3m_min_gates=3
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=4 f64_3m_min_gates=4
timeout -k 10 200 python tools/prof_case.py pass 30 layers=40 max_gates_per_pass=4
timeout -k 10 200 python tools/prof_case.py grad32 28
timeout -k 10 200 python tools/prof_case.py grad 28
timeout -k 10 200 python tools/prof_case.py grad 20 reps=20
timeout -k 10 200 python tools/prof_case.py grad 12 layers=20 reps=20
} > $O/r02at_cases.jsonl 2> $O/r02at_cases.err
cut -c1-330 $O/r02at_cases.jsonl; tail -n 3 $O/r02at_cases.err
timeout -k 10 1200 python -m pytest tests -q -m gpu > $O/r02at_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02at_pytest_gpu.log
tail -n 4 $O/r02at_pytest_gpu.log
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02at_smoke.log 2>&1; echo "rc=$?" >> $O/r02at_smoke.log; tail -n 2 $O/r02at_smoke.log
timeout -k 10 1500 python bench.py --steps 5 --warmup 3 > $O/r02at_bench.json 2> $O/r02at_bench.err; echo "bench rc=$?"; cut -c1-400 $O/r02at_bench.json; tail -n 3 $O/r02at_bench.err
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02at_bench_reference.json 2> $O/r02at_bench_reference.err; cut -c1-300 $O/r02at_bench_reference.json
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02at_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02at_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ls -la $O | grep r02at
