#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-330 gpurun_out/gradients.jsonl; tail -3 gpurun_out/gradients.err
echo "== measure 30q"; timeout 600 python - <<'PY'
import time, numpy as np, torch, qcb200 as qc
qc.init(0)
for dt in (np.complex128, np.complex64):
    st = qc.B200State(30, dt)
    c = qc.GenericBrickworkCircuit(30, 4); c.gate_angles[...] = np.random.default_rng(0).uniform(0, 6.28, c.gate_angles.shape)
    qc.apply_(st, st, c)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    for mode in (0, 1):
        qc.set_option("force_simple", mode)
        E = qc.measure_(None, H, st); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): E = qc.measure_(None, H, st)
        t = (time.perf_counter() - t0) / 3
        print(np.dtype(dt).name, "gather" if mode else "tiled", "E=%.12f" % E, "ms=%.2f" % (t * 1e3), "GB/s of one read=%.0f" % ((np.dtype(dt).itemsize << 30) / t / 1e9))
    qc.set_option("force_simple", 0)
PY
echo "== ncu measure"
CMD="python tools/bench_gradients.py C5"
$CMD > gpurun_out/plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:measure_tile -s 0 -c 3 -f -o gpurun_out/prof_measure $CMD > gpurun_out/ncu_measure.log 2>&1
echo "ncu measure rc=$?"; tail -2 gpurun_out/ncu_measure.log
