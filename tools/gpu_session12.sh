#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-330 gpurun_out/gradients.jsonl; tail -3 gpurun_out/gradients.err
echo "== measure 30q"; timeout 600 python - <<'PY'
import time, numpy as np, torch, qcb200 as qc
qc.init(0)
for dt in (np.complex128, np.complex64):
    st = qc.B200State(30, dt)
    c = qc.GenericBrickworkCircuit(30, 4); c.gate_angles[...] = np.random.default_rng(0).uniform(0, 6.28, c.gate_angles.shape)
    qc.apply_(st, st, c)
    for H in (qc.TFIMHamiltonian(1.0, 0.1, 0.5), qc.EastModelHamiltonian(0.1, -0.1)):
        for mode in (0, 1):
            qc.set_option("force_simple", mode)
            E = qc.measure_(None, H, st); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3): E = qc.measure_(None, H, st)
            t = (time.perf_counter() - t0) / 3
            print(np.dtype(dt).name, type(H).__name__, "gather" if mode else "tiled", "E=%.12f" % E, "ms=%.2f" % (t * 1e3), flush=True)
        qc.set_option("force_simple", 0)
PY
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench c64"; timeout 600 python bench.py --steps 2 --warmup 2 --dtype c64 --no-cpu-baseline > gpurun_out/bench_c64.json 2> gpurun_out/bench_c64.err; cut -c1-260 gpurun_out/bench_c64.json
echo "== ncu"
CMD="python tools/bench_gradients.py C5"
$CMD > gpurun_out/plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"measure_tile|adjoint_pass" -s 0 -c 5 -f -o gpurun_out/prof_grad $CMD > gpurun_out/ncu_grad.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_grad.log
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 30 -c 2 -f -o gpurun_out/prof_tile $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu tile rc=$?"
