#!/bin/bash
# One gpurun call: smoke, fp64 roof probe, GPU parity tests, bench, tile-shape sweep, ncu launch list + full capture.
# usage: tools/gpu_session.sh [quick]
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== fp64 peak"; timeout 120 ./tools/fp64_peak > gpurun_out/fp64_peak.json 2>&1; cat gpurun_out/fp64_peak.json
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "quick" ]; then
echo "== sweep"
for tb in 11 12; do for lb in 3; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --tile-bits $tb --low-bits $lb >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done; done
for mg in 3 5; do
  timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --max-gates-per-pass $mg >> gpurun_out/sweep_maxg.jsonl 2>> gpurun_out/sweep.err
done
for tb in 11 12; do timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --dtype c64 --tile-bits $tb >> gpurun_out/sweep_c64.jsonl 2>> gpurun_out/sweep.err; done
python - <<'PY'
import json
for f in ("sweep.jsonl","sweep_maxg.jsonl","sweep_c64.jsonl"):
    try:
        for ln in open("gpurun_out/"+f):
            d=json.loads(ln); r=d.get("roofline",{})
            print(f, d["config"]["tile_bits"], d["config"]["low_bits"], "value=%.1f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "passes=%s"%r.get("launches_per_step"), "hbm_frac=%.3f"%r.get("frac",0), "fma TF=%.2f"%r.get("fma_roof",{}).get("achieved_tflops",0), d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
fi
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cat gpurun_out/gradients.jsonl; tail -3 gpurun_out/gradients.err
echo "== ncu"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 30 -c 3 -f -o gpurun_out/prof_tile $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --dtype c64"
$CMD2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 30 -c 3 -f -o gpurun_out/prof_tile_c64 $CMD2 > gpurun_out/ncu_full_c64.log 2>&1
echo "ncu full c64 rc=$?"; tail -3 gpurun_out/ncu_full_c64.log
