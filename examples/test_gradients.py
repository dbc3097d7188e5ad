#!/usr/bin/env python
"""Mirror of the reference's examples/test_gradients.jl: the kernel TFIM Hamiltonian and the same Hamiltonian as an
explicit sparse matrix (examples/matrix_tfim.jl) give the same energy and gradients, on a host array and on a device
state.  `B200State(psi0)` stands where the reference writes `CuArray(psi0)`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcb200 as qc  # noqa: E402


def main():
    nbits, nlayers = 8, 4
    J, h, g = 1.0, 0.1, 0.5
    H = qc.build_sparse_tfim_hamiltonian(nbits, J, h, g)  # sparse(build_hamiltonian(nbits, J, h, g))
    Heff = qc.TFIMHamiltonian(J, h, g)
    circuit = qc.GenericBrickworkCircuit(nbits, nlayers)
    circuit.gate_angles[...] = np.random.default_rng(0).standard_normal(circuit.gate_angles.shape) * 0.01
    psi0 = qc.zero_state_tensor(nbits)
    E, grads = qc.gradients(Heff, psi0, circuit, calculate_energy=True)
    E_m, grads_m = qc.gradients(H, psi0, circuit, calculate_energy=True)
    assert np.isclose(E, E_m) and np.allclose(grads, grads_m)
    psi_gpu = qc.B200State(psi0)
    E_gpu, grads_gpu = qc.gradients(Heff, psi_gpu, circuit, calculate_energy=True)
    assert np.isclose(E, E_gpu) and np.allclose(grads, grads_gpu)
    E0, gs = qc.find_tfim_ground_state(nbits, J, h, g)
    print(f"E = {E:.12f} (kernel), {E_m:.12f} (sparse matrix); ground state energy {E0:.12f}; max |grad| = {np.max(np.abs(grads)):.3e}")
    print("ok")


if __name__ == "__main__":
    main()
