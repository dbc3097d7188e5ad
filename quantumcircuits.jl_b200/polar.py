"""Statevector polar optimiser (ext/MPSExt/polar_optimise.jl:83-147 of the reference).

Sweeps the gates of a brickwork circuit and replaces each by the polar factor of its 4x4 environment so that
U_G ... U_1 |0...0> approaches a target state.  The inner reduction is the same 4x4 environment as in the adjoint
gradient sweep (csrc/kernels.cuh: k_env_2q); the 4x4 SVD stays on the host like in the reference.  States live in HBM:
`lower` = U_g ... U_1 |0...0>, `chi` = conj(upper) = U_{g+1}^dagger ... U_G^dagger |target>.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .apply import _apply_list_inplace
from .gates import Generic2SpinGate, Localised2SpinAdjGate
from .hamiltonians import measure_
from .state import B200State


def _sites(N):
    return [2 * idx - 1 for idx in range(1, N // 2 + 1)] + [2 * idx - N for idx in range(N // 2 + 1, N)]


def _loc(M, site):
    return Localised2SpinAdjGate(Generic2SpinGate(np.asarray(M).reshape(2, 2, 2, 2, order="F")), site)


def environment_2q(lam: B200State, psi: B200State, K: int):
    """E[r, c] = sum_rest conj(lam[r, rest]) psi[c, rest] on the qubit pair (K, K+1) -- qcb_environment_2q."""
    out = np.empty(16, dtype=np.complex128)
    _lib.check(_lib.lib().qcb_environment_2q(lam._handle, psi._handle, C.c_int(K), out.ctypes.data_as(C.c_void_p)))
    return out.reshape(4, 4, order="F")


def inner(a: B200State, b: B200State):
    out = np.zeros(2)
    _lib.check(_lib.lib().qcb_state_inner(a._handle, b._handle, out.ctypes.data_as(C.c_void_p)))
    return complex(out[0], out[1])


def polar_optimise(circuit, psi_GS, H, N, iterations=100, use_progress=False):
    """polar_optimise(circuit, psi_GS, H, N; iterations) -- ext/MPSExt/polar_optimise.jl:105-147.

    circuit: list of brick layers, each a list of N-1 gates (2x2x2x2 arrays u[i,j,a,b] or 4x4 matrices); H: a kernel
    Hamiltonian (the reference passes the sparse TFIM matrix; the kernel TFIM gives the same energies).
    Returns (circuit, overlaps, energies) with the gates as 2x2x2x2 arrays, like the reference."""
    assert N % 2 == 0, "the reference's layer layout needs an even number of qubits"
    sites = _sites(N)
    mats = [[np.asarray(g, dtype=np.complex128).reshape(4, 4, order="F") if np.ndim(g) == 4 else np.array(g, dtype=np.complex128)
             for g in layer] for layer in circuit]
    target = psi_GS if isinstance(psi_GS, B200State) else B200State(np.asarray(psi_GS, dtype=np.complex128))
    overlaps, energies = [], []
    for _ in range(iterations):
        lower = B200State(N, np.complex128)
        chi = B200State(target)  # create_upper_state (:52-78): un-apply the whole circuit from the target, one fused call
        undo = [_loc(layer[idx - 1].conj().T, sites[idx - 1]) for layer in reversed(mats) for idx in range(N - 1, 0, -1)]
        _apply_list_inplace(chi, undo)
        for layer in mats:
            for idx in range(1, N):
                site = sites[idx - 1]
                _apply_list_inplace(chi, [_loc(layer[idx - 1], site)])  # upper <- conj(gate) . upper   (:85-86)
                E = environment_2q(chi, lower, site)  # (:88-92)
                U_, _, Vt = np.linalg.svd(E)
                new = np.conj(U_ @ Vt)  # (:93-94)
                layer[idx - 1] = new
                _apply_list_inplace(lower, [_loc(new, site)])  # (:97-98)
        overlaps.append(abs(inner(target, lower)))  # (:140)
        energies.append(measure_(None, H, lower))  # (:141)
    out = [[m.reshape(2, 2, 2, 2, order="F") for m in layer] for layer in mats]
    return out, np.array(overlaps), np.array(energies)
