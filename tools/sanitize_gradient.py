#!/usr/bin/env python
"""Small gradient through both backward passes, meant to run under compute-sanitizer (racecheck / memcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcb200 as qc

qc.init(0)
rng = np.random.default_rng(0)
for n, layers, atb in [(12, 3, 0), (13, 3, 9), (13, 2, 11), (13, 2, 12)]:
    c = qc.GenericBrickworkCircuit(n, layers)
    c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    psi0 = qc.B200State(n, np.complex128)
    qc.set_option("adjoint_tile_bits", atb)
    out = []
    for mma in (0, 1):
        qc.set_option("adjoint_mma", mma)
        out.append(qc.gradients(H, psi0, c, calculate_energy=True))
    print(n, layers, atb, "max |dgrad| =", float(np.max(np.abs(out[0][1] - out[1][1]))))
