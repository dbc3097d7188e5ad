#!/usr/bin/env python
"""Mirror of the reference's examples/test_gpu.jl: energy and gradients of a 14-qubit, 4-half-layer circuit computed by
the reference algorithm on the CPU (the oracle's restatement of src/gradients.jl:88-152) and by the B200 backend.
The only change against the reference script is `B200State(psi0)` where it writes `CuArray(psi0)`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (test infrastructure: stands in for the Julia CPU path here)
import qcb200 as qc  # noqa: E402


def main():
    nbits, nlayers = 14, 4
    J, h, g = 1.0, 0.2, 0.5
    H = qc.TFIMHamiltonian(J, h, g)
    circuit = qc.GenericBrickworkCircuit(nbits, nlayers)
    circuit.gate_angles[...] = np.random.default_rng(1).standard_normal(circuit.gate_angles.shape) * 0.01
    psi0 = qc.zero_state_tensor(nbits)
    psi_gpu = qc.B200State(psi0)
    E_actual, correct_grads = oracle.gradients_brickwork(psi0.reshape(-1, order="F"), nbits, nlayers, circuit.gate_angles, 0, H.params())
    E_test, gpu_grads = qc.gradients(H, psi_gpu, circuit, calculate_energy=True)
    print("E (reference algorithm, CPU):", E_actual)
    print("E (B200 backend):            ", E_test)
    print("max |grad difference|:       ", float(np.max(np.abs(correct_grads - gpu_grads))))
    assert np.isclose(E_actual, E_test, rtol=1e-12) and np.allclose(correct_grads, gpu_grads, rtol=0, atol=1e-12)
    print("ok")


if __name__ == "__main__":
    main()
