"""Gate types of the reference (src/gates.jl), same names and matrices.

`mat(gate)` returns the Julia array's contents as a NumPy array indexed the same way
(u[out, in] for 1-qubit gates, u[i, j, a, b] for 2-qubit gates).  All algebra on 2x2 / 4x4
matrices is host work in the reference too; the 15-angle builder goes through the C ABI
(`qcb_build_general_unitary_gate`) so Julia and Python hosts share one implementation.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


class AbstractGate:  # src/gates.jl:1
    pass


class Abstract1SpinGate(AbstractGate):
    pass


class Abstract2SpinGate(AbstractGate):
    pass


def _t4(m4):
    """4x4 (row i+2j, col a+2b) -> (2,2,2,2) array u[i,j,a,b]."""
    return np.asarray(m4).reshape(2, 2, 2, 2, order="F")


_H = np.array([[1, 1], [1, -1]], dtype=np.float64) / np.sqrt(2)
_X = np.array([[0, 1], [1, 0]], dtype=np.float64)
_Z = np.array([[1, 0], [0, -1]], dtype=np.float64)
_I = np.eye(2)
_N = np.array([[1, 0], [0, 0]], dtype=np.float64)
_CNOT = _t4(np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.float64))
_RCNOT = _t4(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.float64))


class XGate(Abstract1SpinGate):
    def mat(self):
        return _X


class ZGate(Abstract1SpinGate):
    def mat(self):
        return _Z


class NGate(Abstract1SpinGate):
    def mat(self):
        return _N


class HadamardGate(Abstract1SpinGate):
    def mat(self):
        return _H


class IdentityGate(Abstract1SpinGate):
    def mat(self):
        return _I


class CNOTGate(Abstract2SpinGate):
    def mat(self):
        return _CNOT


class ReverseCNOTGate(Abstract2SpinGate):
    def mat(self):
        return _RCNOT


@dataclass
class RZGate(Abstract1SpinGate):  # src/gates.jl:36-43
    angle: float

    def mat(self):
        c, s = np.cos(self.angle / 2), np.sin(self.angle / 2)
        return np.array([[c - 1j * s, 0], [0, c + 1j * s]])


@dataclass
class RYGate(Abstract1SpinGate):  # src/gates.jl:44-51
    angle: float

    def mat(self):
        c, s = np.cos(self.angle / 2), np.sin(self.angle / 2)
        return np.array([[c, -s], [s, c]])


@dataclass
class RXGate(Abstract1SpinGate):  # src/gates.jl:52-59
    angle: float

    def mat(self):
        c, s = np.cos(self.angle / 2), -1j * np.sin(self.angle / 2)
        return np.array([[c, s], [s, c]])


@dataclass
class RotationGate(Abstract1SpinGate):  # src/gates.jl:61-73
    theta: float
    phi: float
    lam: float

    def mat(self):
        c, s = np.cos(self.theta / 2), np.sin(self.theta / 2)
        p1, p2 = np.exp(1j * self.lam), np.exp(1j * self.phi)
        return np.array([[c, -p1 * s], [p2 * s, c * (p1 * p2)]])


class Generic1SpinGate(Abstract1SpinGate):  # src/gates.jl:112-115
    def __init__(self, array):
        self.array = np.asarray(array)
        assert self.array.shape == (2, 2)

    def mat(self):
        return self.array


class Generic2SpinGate(Abstract2SpinGate):  # src/gates.jl:116-119
    def __init__(self, array):
        self.array = np.asarray(array)
        assert self.array.shape == (2, 2, 2, 2), "2-qubit gates are 2x2x2x2 arrays u[i,j,a,b]"

    def mat(self):
        return self.array

    def adjoint(self):  # src/gradients.jl:27-29
        m = self.array.reshape(4, 4, order="F").conj().T
        return Generic2SpinGate(m.reshape(2, 2, 2, 2, order="F"))


class Localised1SpinGate(Abstract1SpinGate):  # src/gates.jl:121-125
    def __init__(self, gate, target_gate_dim):
        self.gate, self.target_gate_dim = gate, int(target_gate_dim)

    def mat(self):
        return self.gate.mat()


class Localised2SpinAdjGate(Abstract2SpinGate):  # src/gates.jl:127-131 ; acts on K and K+1
    def __init__(self, gate, target_gate_dim):
        self.gate, self.target_gate_dim = gate, int(target_gate_dim)

    def mat(self):
        return self.gate.mat()

    def adjoint(self):  # src/gradients.jl:30-32
        return Localised2SpinAdjGate(adjoint(self.gate), self.target_gate_dim)


def mat(gate, nbits=None):
    if nbits is not None:  # mat(H::EastModelHamiltonian, nbits)  src/hamiltonians/east_model.jl:57
        from .hamiltonians import EastModelHamiltonian, east_model_matrix

        if isinstance(gate, EastModelHamiltonian):
            return east_model_matrix(gate, nbits)
        raise TypeError(f"Unimplemented matrix representation for {type(gate)}")
    if not isinstance(gate, AbstractGate) or not hasattr(gate, "mat"):
        raise TypeError(f"Unimplemented matrix representation for {type(gate)}")  # src/gates.jl:6-8
    return gate.mat()


def adjoint(gate):
    if isinstance(gate, (Generic2SpinGate, Localised2SpinAdjGate)):
        return gate.adjoint()
    if isinstance(gate, Abstract2SpinGate):
        return Generic2SpinGate(np.asarray(gate.mat())).adjoint()
    raise TypeError(type(gate))


def build_general_unitary_gate(angles):
    """src/gates.jl:80-109 : 15 angles -> Generic2SpinGate (computed by the C ABI host routine)."""
    a = np.ascontiguousarray(angles, dtype=np.float64)
    if a.size != 15:
        raise ValueError("Must supply exactly 15 angles.")
    out = np.empty(16, dtype=np.complex128)
    _lib.check(_lib.lib().qcb_build_general_unitary_gate(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return Generic2SpinGate(out.reshape(2, 2, 2, 2, order="F"))


def gate_abi_block(gate):
    """(arity, K, 32 doubles) as qcb_apply_gates expects."""
    m = np.asarray(mat(gate))
    blk = np.zeros(16, dtype=np.complex128)
    if isinstance(gate, Abstract2SpinGate):
        blk[:] = m.astype(np.complex128).reshape(-1, order="F")
        return 2, gate.target_gate_dim, blk
    blk[:4] = m.astype(np.complex128).reshape(-1, order="F")
    return 1, gate.target_gate_dim, blk
