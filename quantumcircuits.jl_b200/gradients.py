"""gradients / gradients! / optimise! / propagate_forwards! (src/gradients.jl).

The numbers are the reference's exact parameter-shift gradients (src/gradients.jl:120-139); they are
computed with the O(G) adjoint sweep instead of the reference's O(G^2) un-apply / re-propagate loop
(SURVEY.md appendix A).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .apply import B200ApplyCache, _apply_circuit_inplace, construct_apply_cache
from .hamiltonians import AbstractKernelHamiltonian, _csr_arrays, _is_matrix
from .state import B200State


def construct_grads_cache(psi):
    """src/gradients.jl:73-81 : (cache, psi', psi'', psi), splattable into gradients_.

    The adjoint sweep keeps its two work states inside the library, so the three buffers are only
    place-holders that preserve the reference's tuple shape (examples/hpc/utils.jl:57)."""
    cache = construct_apply_cache(psi)
    return (cache, None, None, None)


def gradients_(cache, psi_p, psi_pp, psi, H, psi0, circuit, calculate_energy=False):
    """gradients!(cache, psi', psi'', psi, H, psi0, circuit; calculate_energy) -- src/gradients.jl:88-152."""
    # (a dist.ShardedState works too: collective call, every rank receives the energy and the full gradient)
    st0 = psi0 if isinstance(psi0, B200State) or hasattr(psi0, "_handle") else B200State(psi0)
    a = circuit.angles_abi()
    E = C.c_double(0.0)
    grads = np.zeros(15 * circuit.ngates, dtype=np.float64)
    if _is_matrix(H):
        # H::AbstractMatrix (dense or sparse, Hermitian): the reference's generic gradients! over measure!(psi', H, psi),
        # src/circuits.jl:56-58, examples/test_gradients.jl:25
        n, rowptr, colidx, vals = _csr_arrays(H)
        _lib.check(_lib.lib().qcb_gradients_brickwork_csr(st0._handle, C.c_int(circuit.nbits), C.c_int(circuit.nlayers),
                                                          a.ctypes.data_as(C.c_void_p), C.c_longlong(n), rowptr.ctypes.data_as(C.c_void_p),
                                                          colidx.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p),
                                                          C.c_int(1 if calculate_energy else 0), C.byref(E), grads.ctypes.data_as(C.c_void_p)))
        grads = grads.reshape(15, circuit.ngates, order="F")
        return (E.value, grads) if calculate_energy else grads
    if not isinstance(H, AbstractKernelHamiltonian):
        raise TypeError(f"gradients: unsupported Hamiltonian type {type(H)}")
    hp = H.params()
    _lib.check(_lib.lib().qcb_gradients_brickwork(st0._handle, C.c_int(circuit.nbits), C.c_int(circuit.nlayers),
                                                  a.ctypes.data_as(C.c_void_p), C.c_int(H.kind),
                                                  hp.ctypes.data_as(C.c_void_p), C.c_int(1 if calculate_energy else 0),
                                                  C.byref(E), grads.ctypes.data_as(C.c_void_p)))
    grads = grads.reshape(15, circuit.ngates, order="F")  # similar(circuit.gate_angles)  :89
    return (E.value, grads) if calculate_energy else grads


def gradients(H, psi, circuit, **kwargs):
    """src/gradients.jl:83-86"""
    return gradients_(*construct_grads_cache(psi), H, psi, circuit, **kwargs)


def propagate_forwards_(cache, psi_p, psi, circuit, start_layer, layer_gate_idx, gate_idx):
    """propagate_forwards!(cache, psi', psi, circuit, start_layer, layer_gate_idx, gate_idx) -- src/gradients.jl:54-71.

    Applies gates gate_idx..G (1-based, as in Julia) and returns (result, other) like the reference."""
    psi_p.assign(psi)
    _apply_circuit_inplace(psi_p, circuit, first=gate_idx - 1, last=circuit.ngates)
    return psi_p, psi


def optimise_(circuit, H, psi0, epochs, lr, use_progress=False):
    """optimise!(circuit, H, psi0, epochs, lr) -- src/gradients.jl:1-17, with quirk Q1 resolved to the intended
    semantics (calculate_energy=true, theta <- theta - lr * grad).  circuit.gate_angles is updated in place;
    returns `energies` laid out like the reference's (length epochs+1, [0] never written, [-1] = final measure)."""
    # (a dist.ShardedState works too: collective call, every rank receives the energy and the full gradient)
    st0 = psi0 if isinstance(psi0, B200State) or hasattr(psi0, "_handle") else B200State(psi0)
    energies = np.zeros(epochs + 1, dtype=np.float64)
    if _is_matrix(H):
        # optimise! is generic over H (src/gradients.jl:1-17); a matrix Hamiltonian runs the same loop on the host, one
        # qcb_gradients_brickwork_csr call per epoch (the reference's small-n cross-check path)
        from .hamiltonians import measure

        for e in range(1, epochs + 1):
            E, g = gradients_(None, None, None, None, H, st0, circuit, calculate_energy=True)
            energies[e] = E                      # :11
            circuit.gate_angles[...] -= lr * g   # :13
        if epochs >= 0:
            energies[epochs] = measure(H, st0, circuit)  # :15
        return energies
    if not isinstance(H, AbstractKernelHamiltonian):
        raise TypeError(f"optimise!: unsupported Hamiltonian type {type(H)}")
    a = circuit.angles_abi().copy()
    hp = H.params()
    _lib.check(_lib.lib().qcb_optimise_brickwork(st0._handle, C.c_int(circuit.nbits), C.c_int(circuit.nlayers),
                                                 a.ctypes.data_as(C.c_void_p), C.c_int(H.kind), hp.ctypes.data_as(C.c_void_p),
                                                 C.c_int(epochs), C.c_double(lr), energies.ctypes.data_as(C.c_void_p)))
    circuit.gate_angles[...] = a.reshape(15, circuit.ngates, order="F")
    return energies
