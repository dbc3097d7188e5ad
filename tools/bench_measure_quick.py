#!/usr/bin/env python
"""Quick timing of the expectation value: TMA-fed persistent kernel (default) vs the round-1 tiled kernel, both dtypes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qcb200 as qc
qc.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(5)
c = qc.GenericBrickworkCircuit(n, 2)
c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
for Hname, H in (("tfim", qc.TFIMHamiltonian(1.0, 0.9045, 1.4)), ("east", qc.EastModelHamiltonian(0.1, -0.1))):
    for dt in (np.complex128, np.complex64):
        st = qc.B200State(n, dt)
        qc.apply_(st, st, c)
        for tma in (1, 0):
            qc.set_option("measure_tma", tma)
            E = qc.measure_(None, H, st)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); E = qc.measure_(None, H, st); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            gb = np.dtype(dt).itemsize * 2.0 ** n / 1e9
            print(f"{Hname} n={n} {np.dtype(dt).name} tma={tma} ms={np.median(ts):.3f} (min {min(ts):.3f}) read-once GB/s={gb / np.median(ts) * 1e3:.0f} passes={qc.get_stat('passes'):.0f} E={E:.9f}", flush=True)
        qc.set_option("measure_tma", 1)
        del st
