"""Kernel Hamiltonians and measure / measure! (src/hamiltonians/*.jl)."""
from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass

import numpy as np

from . import _lib
from .apply import _apply_circuit_inplace
from .state import B200State


class AbstractKernelHamiltonian:  # src/hamiltonians/hamiltonians.jl:5
    kind = None

    def params(self):
        raise NotImplementedError("Unimplemented kernel")  # :7


def imaginary_energy_limit(H):  # :8
    return 1e-8


@dataclass
class TFIMHamiltonian(AbstractKernelHamiltonian):
    """H = -J (sum Z_i Z_{i+1} + g sum X_i + h sum Z_i), open chain (src/hamiltonians/tfim.jl:3-15).
    Field order is the reference's: TFIMHamiltonian(J, h, g); g multiplies X, h multiplies Z."""
    J: float
    h: float
    g: float
    kind = _lib.QCB_HAM_TFIM

    def params(self):
        return np.array([self.J, self.h, self.g], dtype=np.float64)


class EastModelHamiltonian(AbstractKernelHamiltonian):
    """src/hamiltonians/east_model.jl:16-27 : EastModelHamiltonian(c, s) -> A = -e^{-s} sqrt(c(1-c)), B = 1-2c."""
    kind = _lib.QCB_HAM_EAST

    def __init__(self, c, s, A=None, B=None):
        self.c, self.s = float(c), float(s)
        self.A = float(-np.exp(-s) * np.sqrt(c * (1 - c))) if A is None else float(A)
        self.B = float(1 - 2 * c) if B is None else float(B)

    def params(self):
        return np.array([self.c, self.A, self.B], dtype=np.float64)


def _measure_state(H, st):
    E = np.zeros(2, dtype=np.float64)
    p = H.params()
    fn = _lib.lib().qcb_measure_tfim if H.kind == _lib.QCB_HAM_TFIM else _lib.lib().qcb_measure_east
    _lib.check(fn(st._handle, C.c_double(p[0]), C.c_double(p[1]), C.c_double(p[2]), E.ctypes.data_as(C.c_void_p)))
    if E[1] > imaginary_energy_limit(H):  # src/hamiltonians/hamiltonians.jl:43-45
        warnings.warn(f"Imaginary energy of {E[1]} -> above threshold!")
    return float(E[0])


def _is_matrix(H):
    return isinstance(H, np.ndarray) or (hasattr(H, "tocsr") and hasattr(H, "shape"))


def _csr_arrays(H):
    """(nrows, rowptr, colidx, vals) of a numpy / scipy-sparse matrix as the ABI wants them (0-based, int64, complex128);
    a dense matrix travels as the CSR that holds every entry."""
    if len(H.shape) != 2:  # @assert ndims(H) == 2   src/circuits.jl:50
        raise AssertionError("ndims(H) == 2")
    if H.shape[0] != H.shape[1]:
        raise ValueError("H must be square")
    if isinstance(H, np.ndarray):
        n = H.shape[0]
        vals = np.ascontiguousarray(H, dtype=np.complex128).reshape(-1)
        rowptr = np.arange(0, n * H.shape[1] + 1, H.shape[1], dtype=np.int64)
        colidx = np.tile(np.arange(H.shape[1], dtype=np.int64), n)
    else:
        M = H.tocsr()
        n = M.shape[0]
        vals = np.ascontiguousarray(M.data, dtype=np.complex128)
        rowptr = np.ascontiguousarray(M.indptr, dtype=np.int64)
        colidx = np.ascontiguousarray(M.indices, dtype=np.int64)
    return n, rowptr, colidx, vals


def _measure_matrix(H, st):
    """measure(H::AbstractMatrix, psi) = real(dot(psi, H, psi)) -- src/circuits.jl:49-55, on the device through
    qcb_measure_csr."""
    n, rowptr, colidx, vals = _csr_arrays(H)
    E = np.zeros(2, dtype=np.float64)
    _lib.check(_lib.lib().qcb_measure_csr(st._handle, C.c_longlong(n), rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p),
                                          vals.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p)))
    if not E[1] < 1e-12:  # @assert imag(measurement) < 1e-12   :53
        raise AssertionError(f"imag(measurement) = {E[1]} >= 1e-12")
    return float(E[0])


def measure_(psi_scratch, H, psi):
    """measure!(psi', H, psi) -- src/hamiltonians/hamiltonians.jl:31-50 (kernel Hamiltonians) and src/circuits.jl:56-58
    (matrix Hamiltonians).  The scratch state psi' is accepted for signature compatibility and left untouched (the fused
    sweep-and-reduce kernel needs no scratch)."""
    st = psi if isinstance(psi, B200State) else B200State(psi)
    if _is_matrix(H):
        return _measure_matrix(H, st)
    if not isinstance(H, AbstractKernelHamiltonian):
        raise TypeError(f"measure: unsupported Hamiltonian type {type(H)}")
    return _measure_state(H, st)


def measure(H, psi, circuit=None):
    """measure(H, psi) | measure(H, psi, circuit) -- src/hamiltonians/hamiltonians.jl:11-30; matrix Hamiltonians
    (numpy arrays, scipy sparse matrices): src/circuits.jl:49-62."""
    if circuit is None:
        return measure_(None, H, psi)
    if isinstance(psi, B200State):  # the input is not modified (:19): psi' = circuit(psi) out of place, no copy
        from .apply import _apply_circuit_from

        st = psi.similar()
        _apply_circuit_from(st, psi, circuit)
    else:
        st = B200State(psi)
        _apply_circuit_inplace(st, circuit)
    return measure_(None, H, st)


def build_sparse_tfim_hamiltonian(n, J, h, g):
    """src/hamiltonians/tfim.jl:58-84 : the TFIM as an explicit sparse matrix (scipy CSR), diagonal -(J zz + J h z),
    -J g on every single-bit flip."""
    import scipy.sparse as sp

    Cs = np.arange(1 << n, dtype=np.int64)
    zz = np.zeros(1 << n)
    for i in range(n - 1):
        r = (Cs >> i) & 3
        zz += np.where((r == 3) | (r == 0), 1.0, -1.0)  # :66-69
    z = n - 2 * sum(((Cs >> i) & 1) for i in range(n))  # :72
    diag = -(J * zz + (J * h) * z)  # :74
    rows = np.concatenate([Cs ^ (1 << k) for k in range(n)] + [Cs])
    cols = np.concatenate([Cs] * (n + 1))
    vals = np.concatenate([np.full((1 << n) * n, -(J * g), dtype=np.float64), diag])  # :78-83
    return sp.csr_matrix((vals, (rows, cols)), shape=(1 << n, 1 << n))


def find_tfim_ground_state(nbits, J, h, g):
    """src/hamiltonians/tfim.jl:86-92 : smallest eigenpair of the sparse TFIM matrix (host eigensolver, as in the
    reference: KrylovKit there, scipy's Lanczos here).  Returns (energy, state vector)."""
    import scipy.sparse.linalg as spl

    H = build_sparse_tfim_hamiltonian(nbits, J, h, g)
    if nbits <= 2:
        w, v = np.linalg.eigh(H.toarray())
        return float(w[0]), v[:, 0]
    w, v = spl.eigsh(H, k=1, which="SA")
    return float(w[0]), v[:, 0]


def east_model_matrix(H, nbits):
    """mat(H::EastModelHamiltonian, nbits) -- src/hamiltonians/east_model.jl:57-81 : dense 2^n x 2^n matrix built from
    Kronecker products of X and N = |0><0| on the reference's qubit order."""
    from .gates import Localised1SpinGate, NGate, XGate
    from .utils import convert_gates_to_matrix

    X = lambda j: Localised1SpinGate(XGate(), j)  # noqa: E731
    N = lambda j: Localised1SpinGate(NGate(), j)  # noqa: E731
    M = convert_gates_to_matrix(nbits, [X(1)]).astype(np.float64)
    for j in range(2, nbits + 1):
        M = M + convert_gates_to_matrix(nbits, [N(j - 1), X(j)]).real
    M = M * H.A
    M = M + H.B * convert_gates_to_matrix(nbits, [N(1)]).real
    for j in range(2, nbits + 1):
        m1 = convert_gates_to_matrix(nbits, [N(j - 1), N(j)]).real
        m2 = convert_gates_to_matrix(nbits, [N(j - 1)]).real
        M = M + H.B * (m1 + H.c * m2)
    return M + H.c * np.eye(1 << nbits)
