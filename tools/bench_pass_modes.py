#!/usr/bin/env python
"""30-qubit brickwork circuit (BASELINE C3) under the two tile movers and several fusion caps: per-launch event timing of
the gate passes (option time_passes), HBM GB/s per launch and gate-apps/s."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qcb200 as qc
qc.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(1)
c = qc.GenericBrickworkCircuit(n, layers)
c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
st = qc.B200State(n, np.complex128)
bytes_pass = 2 * 16.0 * 2 ** n
for cap in (1 << 30, 8, 5, 4, 3, 2):
    for mode in (0, 1):
        qc.set_option("pass_tma", mode)
        qc.set_option("max_gates_per_pass", cap)
        qc.apply_(st, st, c)
        torch.cuda.synchronize()
        qc.set_option("time_passes", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); qc.apply_(st, st, c); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        k, kms = qc.get_stat("pass_count"), qc.get_stat("pass_ms")
        qc.set_option("time_passes", 0)
        print(json.dumps({"n": n, "layers": layers, "max_gates_per_pass": cap if cap < (1 << 30) else None, "pass_tma": mode, "passes": k, "tma_passes": qc.get_stat("tma_passes"),
                          "circuit_ms": ms, "avg_launch_ms": kms / k, "hbm_gbs": bytes_pass / (kms / k * 1e-3) / 1e9, "frac_of_6457": bytes_pass / (kms / k * 1e-3) / 1e9 / 6457.1,
                          "gate_apps_per_s": c.ngates / (ms * 1e-3), "norm2": st.norm2()}), flush=True)
