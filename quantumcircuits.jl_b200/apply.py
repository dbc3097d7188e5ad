"""apply / apply! entry points of the reference (src/apply.jl, src/circuits.jl:19-42).

Julia's `apply!` is spelled `apply_` here.  The (psi', psi) ping-pong signature is kept: kernels
run in place on a copy made into psi', and psi' is returned, so callers that swap buffers keep
working (SURVEY.md 8b "In-place vs out-of-place").  Host NumPy arrays are accepted wherever
the reference accepts an `Array`: they are uploaded, processed on the GPU and downloaded -- there is
no CPU arithmetic path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .circuits import GenericBrickworkCircuit
from .gates import AbstractGate, Abstract2SpinGate, Localised1SpinGate, Localised2SpinAdjGate, gate_abi_block
from .state import B200State


class B200ApplyCache:
    """Stands where CPUApplyCache / GPUApplyCache stand (src/apply.jl:89-104).  Stateless: gate
    matrices travel as kernel parameters, so there is no staging array to cache."""


def construct_apply_cache(psi):
    _lib.init()
    return B200ApplyCache()


def _check_pair(psi_out, psi):
    if not isinstance(psi_out, B200State) or not isinstance(psi, B200State):
        raise AssertionError("The backend of each array must be the same")  # src/apply.jl:107
    if psi_out.nqubits != psi.nqubits or psi_out.dtype != psi.dtype:
        raise ValueError("ψ′ must be the same size as ψ.")  # src/apply.jl:108


def _apply_list_inplace(state, gates, src=None):
    """gates on `state` in place, or (src given) state = gates(src) with src untouched: the first fused pass reads src."""
    if not gates:
        if src is not None and src is not state:
            state.assign(src)
        return
    G = len(gates)
    K = np.empty(G, dtype=np.int32)
    ar = np.empty(G, dtype=np.int32)
    mats = np.empty((G, 16), dtype=np.complex128)
    for i, g in enumerate(gates):
        if not isinstance(g, (Localised1SpinGate, Localised2SpinAdjGate)):
            raise TypeError("gates must be Localised1SpinGate or Localised2SpinAdjGate")
        ar[i], K[i], mats[i] = gate_abi_block(g)
    if src is not None and src is not state:
        _lib.check(_lib.lib().qcb_apply_gates_from(state._handle, src._handle, C.c_int(G), K.ctypes.data_as(C.c_void_p),
                                                   ar.ctypes.data_as(C.c_void_p), mats.ctypes.data_as(C.c_void_p)))
        return
    _lib.check(_lib.lib().qcb_apply_gates(state._handle, C.c_int(G), K.ctypes.data_as(C.c_void_p),
                                          ar.ctypes.data_as(C.c_void_p), mats.ctypes.data_as(C.c_void_p)))


def _apply_circuit_inplace(state, circuit, first=0, last=None, adjoint=False):
    last = circuit.ngates if last is None else last
    a = circuit.angles_abi()
    _lib.check(_lib.lib().qcb_apply_brickwork_range(state._handle, C.c_int(circuit.nbits), C.c_int(circuit.nlayers),
                                                    a.ctypes.data_as(C.c_void_p), C.c_int(first), C.c_int(last),
                                                    C.c_int(1 if adjoint else 0)))


def _apply_circuit_from(dst, src, circuit, first=0, last=None, adjoint=False):
    last = circuit.ngates if last is None else last
    a = circuit.angles_abi()
    _lib.check(_lib.lib().qcb_apply_brickwork_from(dst._handle, src._handle, C.c_int(circuit.nbits), C.c_int(circuit.nlayers),
                                                   a.ctypes.data_as(C.c_void_p), C.c_int(first), C.c_int(last), C.c_int(1 if adjoint else 0)))


def apply_(*args):
    """apply!(psi', psi, gate) | apply!(cache, psi', psi, gate) | apply!(cache, psi', psi, circuit)

    src/apply.jl:20,39,117,121 and src/circuits.jl:19.  Returns the state that holds the result (psi')."""
    if isinstance(args[0], B200ApplyCache):
        args = args[1:]
    psi_out, psi, op = args
    _check_pair(psi_out, psi)
    if isinstance(op, GenericBrickworkCircuit) and psi_out is not psi and isinstance(psi, B200State):
        _apply_circuit_from(psi_out, psi, op)  # psi' = circuit(psi): the first pass reads psi, no copy
        return psi_out
    if isinstance(op, GenericBrickworkCircuit):
        if psi_out is not psi:
            psi_out.assign(psi)
        _apply_circuit_inplace(psi_out, op)
    else:
        _apply_list_inplace(psi_out, [op], src=psi)  # psi' = gate(psi): no copy, no zero fill (src/apply.jl:135)
    return psi_out


def apply(psi, op):
    """apply(psi, gate) | apply(psi, gates) | apply(psi, circuit) -- src/apply.jl:1-18, src/circuits.jl:37-42.

    The input is not modified.  NumPy in -> NumPy out (same shape), B200State in -> B200State out."""
    host = not isinstance(psi, B200State)
    if host and isinstance(op, GenericBrickworkCircuit):
        return apply_circuit_host(psi, op)
    if not host and isinstance(op, GenericBrickworkCircuit):
        st = psi.similar()
        _apply_circuit_from(st, psi, op)
        return st
    gates = [op] if isinstance(op, AbstractGate) else list(op)
    if not host:
        st = psi.similar()
        _apply_list_inplace(st, gates, src=psi)  # the first fused pass reads psi: no copy of the state
        return st
    st = B200State(psi)
    _apply_list_inplace(st, gates)
    if host:
        out = st.to_numpy()
        return out.reshape(np.shape(psi), order="F") if np.ndim(psi) > 1 else out
    return st


def apply_circuit_host(psi, circuit, out=None):
    """apply(psi::Array, circuit) through the one-shot host-buffer entry point (H2D, circuit, D2H)."""
    _lib.init()
    src = np.asarray(psi)
    flat = src.reshape(-1, order="F") if src.ndim > 1 else src
    flat = np.ascontiguousarray(flat)
    if flat.dtype not in (np.complex64, np.complex128):
        flat = flat.astype(np.complex128)
    if flat.size != 1 << circuit.nbits:
        raise ValueError("state size does not match the circuit")
    res = np.empty_like(flat) if out is None else out
    a = circuit.angles_abi()
    _lib.check(_lib.lib().qcb_apply_brickwork_host(C.c_int(1 if flat.dtype == np.complex128 else 0), C.c_int(circuit.nbits),
                                                   C.c_int(circuit.nlayers), a.ctypes.data_as(C.c_void_p),
                                                   flat.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p)))
    return res.reshape(src.shape, order="F") if src.ndim > 1 else res


def right_apply_gate_(M_out, M, gate):
    """right_apply_gate!(M', M, gate::Localised2SpinAdjGate) -- src/apply.jl:144-182 (operator evolution: M' = M * U).

    The reference contracts the gate into the column index of the 2^n x 2^n matrix,
    M'[pre, i, j, post] += M[pre, a, b, post] * u[a, b, i, j] on dims K + n, K + n + 1 of the 2n-dim tensor.  Seen as a
    2n-qubit state (column-major: row bits first), that is the 2-qubit gate v[i, j, a, b] = u[a, b, i, j] on qubits
    (K + n, K + n + 1): one fused tile pass on the device."""
    M = np.asarray(M)
    if M.ndim != 2:
        raise ValueError("Must be a matrix")
    if M.shape[0] != M.shape[1]:
        raise ValueError("Must be a square matrix")
    if np.shape(M_out) != M.shape:
        raise ValueError("M′ must be the same size as M")
    nbits = int(M.shape[0]).bit_length() - 1
    if M.shape[0] != 1 << nbits:
        raise AssertionError("The matrix must have a power of two edge")
    if not isinstance(gate, Localised2SpinAdjGate):
        raise TypeError("gate must be a Localised2SpinAdjGate")
    K = gate.target_gate_dim
    if not (0 < K < nbits):
        raise _lib.GateOutsideQubitSpace("The gate cannot be applied as it sits outside the qubit space.")
    from .gates import Generic2SpinGate

    u = np.asarray(gate.mat(), dtype=np.complex128)            # u[i, j, a, b]
    v = np.ascontiguousarray(u.transpose(2, 3, 0, 1))          # v[i, j, a, b] = u[a, b, i, j]
    dt = np.complex64 if M.dtype in (np.complex64, np.float32) else np.complex128
    st = B200State(np.asarray(M, dtype=dt).reshape(-1, order="F"))
    _apply_list_inplace(st, [Localised2SpinAdjGate(Generic2SpinGate(v), K + nbits)])
    res = st.to_numpy().reshape(M.shape, order="F")
    M_out[...] = res if np.iscomplexobj(M_out) else res.real
    return M_out
