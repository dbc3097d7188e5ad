"""ctypes front-end of the C oracle (oracle/qc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle_np as np_oracle  # noqa: F401  (re-export for tests)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libqc_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.qco_brickwork_num_gates.restype = C.c_int
        _lib.qco_get_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _dt(psi):
    if psi.dtype == np.complex128:
        return 1
    if psi.dtype == np.complex64:
        return 0
    raise TypeError(psi.dtype)


def _nq(psi):
    n = int(psi.size).bit_length() - 1
    assert psi.size == 1 << n
    return n


def _gate(u, k):
    u = np.ascontiguousarray(np.asarray(u, dtype=np.complex128).reshape(-1, order="F"))
    assert u.size == k
    return u


def set_threads(t: int):
    lib().qco_set_threads(C.c_int(t))


def max_threads() -> int:
    return lib().qco_get_max_threads()


def build_general_unitary_gate(angles):
    a = np.ascontiguousarray(angles, dtype=np.float64)
    assert a.size == 15
    out = np.empty(16, dtype=np.complex128)
    lib().qco_build_general_unitary_gate(_p(a), _p(out))
    return out.reshape(4, 4, order="F")


def brickwork_num_gates(nbits, nlayers):
    return lib().qco_brickwork_num_gates(C.c_int(nbits), C.c_int(nlayers))


def apply_2q(psi, M, K, scatter=False, out=None):
    """M: 4x4 (row = i+2j, col = a+2b) or (2,2,2,2) Julia-ordered array.  `out`: preallocated psi' (the reference's
    apply!(psi', psi, gate) contract; bench.py ping-pongs two buffers like src/circuits.jl:19-31)."""
    psi = np.ascontiguousarray(psi).reshape(-1)
    if out is None:
        out = np.empty_like(psi)
    assert out.dtype == psi.dtype and out.size == psi.size and out.flags.c_contiguous and out is not psi
    M = np.asarray(M)
    u = _gate(M if M.ndim == 2 else M.reshape(4, 4, order="F"), 16)
    fn = lib().qco_apply_2q_scatter if scatter else lib().qco_apply_2q
    rc = fn(_dt(psi), _nq(psi), _p(out), _p(psi), _p(u), C.c_int(K))
    if rc:
        raise AssertionError("The gate cannot be applied as it sits outside the qubit space.")
    return out


def apply_1q(psi, u, K):
    psi = np.ascontiguousarray(psi).reshape(-1)
    out = np.empty_like(psi)
    rc = lib().qco_apply_1q(_dt(psi), _nq(psi), _p(out), _p(psi), _p(_gate(u, 4)), C.c_int(K))
    if rc:
        raise AssertionError("The gate cannot be applied as it sits outside the qubit space.")
    return out


def _angles(angles, nbits, nlayers):
    a = np.asarray(angles, dtype=np.float64)
    G = brickwork_num_gates(nbits, nlayers)
    if a.ndim == 2:
        assert a.shape == (15, G), a.shape
        a = a.reshape(-1, order="F")
    assert a.size == 15 * G
    return np.ascontiguousarray(a), G


def apply_brickwork(psi, nbits, nlayers, angles, g_begin=0, g_end=None, inplace=False, scratch=None):
    psi = np.ascontiguousarray(psi).reshape(-1)
    if not inplace:
        psi = psi.copy()
    a, G = _angles(angles, nbits, nlayers)
    if scratch is None:
        scratch = np.empty_like(psi)
    assert scratch.dtype == psi.dtype and scratch.size == psi.size
    rc = lib().qco_apply_brickwork_range(_dt(psi), C.c_int(nbits), C.c_int(nlayers), _p(a), _p(psi), _p(scratch),
                                         C.c_int(g_begin), C.c_int(G if g_end is None else g_end))
    assert rc == 0
    return psi


def measure_tfim(psi, J, h, g, scratch=None):
    psi = np.ascontiguousarray(psi).reshape(-1)
    if scratch is None:
        scratch = np.empty_like(psi)
    E = np.zeros(2)
    lib().qco_measure_tfim(_dt(psi), _nq(psi), _p(scratch), _p(psi), C.c_double(J), C.c_double(h), C.c_double(g), _p(E))
    return complex(E[0], E[1])


def measure_east(psi, c, A, B):
    psi = np.ascontiguousarray(psi).reshape(-1)
    scratch = np.empty_like(psi)
    E = np.zeros(2)
    lib().qco_measure_east(_dt(psi), _nq(psi), _p(scratch), _p(psi), C.c_double(c), C.c_double(A), C.c_double(B), _p(E))
    return complex(E[0], E[1])


def gradients_brickwork(psi0, nbits, nlayers, angles, ham_kind, ham_params, want_energy=True):
    """The reference's O(G^2) parameter-shift sweep (src/gradients.jl:88-152)."""
    psi0 = np.ascontiguousarray(psi0).reshape(-1)
    a, G = _angles(angles, nbits, nlayers)
    hp = np.ascontiguousarray(ham_params, dtype=np.float64)
    E = C.c_double(0.0)
    grads = np.zeros(15 * G)
    rc = lib().qco_gradients_brickwork(_dt(psi0), C.c_int(nbits), C.c_int(nlayers), _p(a), _p(psi0), C.c_int(ham_kind),
                                       _p(hp), C.c_int(1 if want_energy else 0), C.byref(E), _p(grads))
    assert rc == 0
    grads = grads.reshape(15, G, order="F")
    return (E.value, grads) if want_energy else grads


def optimise_brickwork(psi0, nbits, nlayers, angles, ham_kind, ham_params, epochs, lr):
    psi0 = np.ascontiguousarray(psi0).reshape(-1)
    a, G = _angles(angles, nbits, nlayers)
    a = a.copy()
    hp = np.ascontiguousarray(ham_params, dtype=np.float64)
    energies = np.zeros(epochs + 1)
    rc = lib().qco_optimise_brickwork(_dt(psi0), C.c_int(nbits), C.c_int(nlayers), _p(a), _p(psi0), C.c_int(ham_kind),
                                      _p(hp), C.c_int(epochs), C.c_double(lr), _p(energies))
    assert rc == 0
    return energies, a.reshape(15, G, order="F")
