#!/usr/bin/env python
"""Mirror of the reference's examples/brickwork_unitary.jl (without the plotting): gradient descent of a 12-qubit,
8-half-layer brickwork circuit on the TFIM energy, three random restarts, on the B200 backend.

    python examples/brickwork_unitary.py [nbits] [nlayers] [epochs]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcb200 as qc  # noqa: E402


def main():
    nbits = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    nlayers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 80
    J, h, g = 1.0, 0.2, 0.5
    H = qc.TFIMHamiltonian(J, h, g)
    circuit = qc.GenericBrickworkCircuit(nbits, nlayers)
    lr = 0.01
    rng = np.random.default_rng(0)
    for rep in range(3):
        circuit.gate_angles[...] = rng.standard_normal(circuit.gate_angles.shape) * 0.01  # randn! .* 0.01
        psi0 = qc.B200State(qc.zero_state_tensor(nbits))
        t0 = time.perf_counter()
        energies = qc.optimise_(circuit, H, psi0, epochs, lr)
        dt = time.perf_counter() - t0
        print(f"run {rep}: E[1] = {energies[1]:.6f} -> E[end] = {energies[-1]:.6f}   ({epochs} epochs, {dt:.2f} s, "
              f"{circuit.ngates} gates, {15 * circuit.ngates} angles)")
    psi = qc.apply(qc.zero_state_tensor(nbits), circuit)
    print("final state norm:", float(np.linalg.norm(psi.reshape(-1))))


if __name__ == "__main__":
    main()
