#!/usr/bin/env python
"""One kernel scenario per invocation, for timing (CUDA events) or for wrapping in ncu.
usage: prof_case.py <case> [n] [key=value options ...]
cases: pass (brickwork circuit, ComplexF64), pass32 (ComplexF32), measure, measure32, grad, grad32
options are qcb_set_option pairs (e.g. pass_tma=0 max_gates_per_pass=3 tma_l2_promotion=2 adjoint_f32=0); layers=K sets the depth."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qcb200 as qc

case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 28
opts = dict(kv.split("=") for kv in sys.argv[2:] if "=" in kv)
layers = int(opts.pop("layers", 4 if case.startswith(("grad", "measure")) else 20))
reps = int(opts.pop("reps", 3))
passlog = opts.pop("passlog", None)  # file: per-launch times of one circuit application with the pass geometry (QCB_PASS_LOG)
if passlog:
    os.environ["QCB_PASS_LOG"] = passlog
qc.init(0)
for k, v in opts.items():
    qc.set_option(k, int(v))
dt = np.complex64 if case.endswith("32") else np.complex128
rng = np.random.default_rng(n)
c = qc.GenericBrickworkCircuit(n, layers)
c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape) if case.startswith("pass") else rng.standard_normal(c.gate_angles.shape) * 0.01
H = qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
st = qc.B200State(n, dt)


def run():
    if case.startswith("pass"):
        qc.apply_(st, st, c)
        return qc.get_stat("passes")
    if case.startswith("measure"):
        return qc.measure_(None, H, st)
    return qc.gradients(H, st, c, calculate_energy=True)[0]


if case.startswith("measure"):
    qc.apply_(st, st, c)
out = run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
if passlog and case.startswith("pass"):
    with open(passlog, "a") as f:
        f.write(f"# {case} n={n} layers={layers} {opts}\n")
    qc.set_option("time_passes", 1)
    run()
    qc.get_stat("pass_ms")
    qc.set_option("time_passes", 0)
print(json.dumps({"case": case, "n": n, "layers": layers, "options": opts, "ms_median": float(np.median(ts)), "ms_min": float(min(ts)), "result": float(out),
                  "passes": qc.get_stat("passes"), "tma_passes": qc.get_stat("tma_passes"), "launches": qc.get_stat("launches")}), flush=True)
