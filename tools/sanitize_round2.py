#!/usr/bin/env python
"""The kernels added in round 2 on small inputs, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):
TMA-fed expectation value (mbarrier ring with two consumer groups), ComplexF32 backward pass (register windows +
transpose-reduction, tile prefetch), deferred reduction of the backward sweep, TMA-fed gate pass."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcb200 as qc

qc.init(0)
rng = np.random.default_rng(0)
for n, dt in [(15, np.complex128), (16, np.complex64), (22, np.complex128)]:
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    st = qc.B200State((v / np.linalg.norm(v)).astype(dt))
    for H in (qc.TFIMHamiltonian(1.0, 0.9045, 1.4), qc.EastModelHamiltonian(0.1, -0.1)):
        qc.set_option("measure_tma", 1)
        a = qc.measure_(None, H, st)
        qc.set_option("measure_tma", 0)
        b = qc.measure_(None, H, st)
        qc.set_option("measure_tma", 1)
        print("measure", n, np.dtype(dt).name, type(H).__name__, "tma - tiled =", a - b)
for n, layers, atb in [(12, 3, 0), (13, 3, 9), (14, 2, 11), (13, 2, 12)]:
    c = qc.GenericBrickworkCircuit(n, layers)
    c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    qc.set_option("adjoint_tile_bits", atb)
    for dt in (np.complex64, np.complex128):
        psi0 = qc.B200State(n, dt)
        out = []
        for defer in (1, 0):
            qc.set_option("adjoint_defer", defer)
            out.append(qc.gradients(H, psi0, c, calculate_energy=True))
        qc.set_option("adjoint_defer", 1)
        print("gradient", n, layers, atb, np.dtype(dt).name, "max |deferred - per pass| =", float(np.max(np.abs(out[0][1] - out[1][1]))))
for n, layers in [(12, 6), (14, 5)]:
    c = qc.GenericBrickworkCircuit(n, layers)
    c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
    outs = []
    for mode in (1, 0):
        qc.set_option("pass_tma", mode)
        st = qc.B200State(n, np.complex128)
        qc.apply_(st, st, c)
        outs.append(st.to_numpy())
    qc.set_option("pass_tma", 0)
    print("gate pass", n, layers, "tma == ldg:", bool(np.array_equal(outs[0], outs[1])))
