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


def measure_(psi_scratch, H, psi):
    """measure!(psi', H, psi) -- src/hamiltonians/hamiltonians.jl:31-50.  The scratch state psi' is accepted for
    signature compatibility and left untouched (the fused sweep-and-reduce kernel needs no scratch)."""
    if not isinstance(H, AbstractKernelHamiltonian):
        raise TypeError("only kernel Hamiltonians are accelerated (matrix Hamiltonians: SURVEY.md 8f, not built)")
    st = psi if isinstance(psi, B200State) else B200State(psi)
    return _measure_state(H, st)


def measure(H, psi, circuit=None):
    """measure(H, psi) | measure(H, psi, circuit) -- src/hamiltonians/hamiltonians.jl:11-30."""
    if circuit is None:
        return measure_(None, H, psi)
    st = B200State(psi)  # copy: the input is not modified (:19)
    _apply_circuit_inplace(st, circuit)
    return _measure_state(H, st)
