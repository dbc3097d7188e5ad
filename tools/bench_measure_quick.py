#!/usr/bin/env python
"""Shortest possible timing of the tiled expectation value, default vs persistent form (28 qubits, TFIM, both dtypes)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qcb200 as qc
qc.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
for dt in (np.complex128, np.complex64):
    st = qc.B200State(n, dt)
    for pers in (0, 1):
        qc.set_option("measure_persistent", pers)
        qc.measure_(None, H, st); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): E = qc.measure_(None, H, st)
        torch.cuda.synchronize()
        print(f"{np.dtype(dt).name} persistent={pers} ms={(time.perf_counter() - t0) / 5 * 1e3:.3f} E={E:.6f}", flush=True)
    del st
