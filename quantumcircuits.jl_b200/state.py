"""Device-resident statevectors.

`B200State` plays the role the `CuArray` plays in the reference's GPU usage
(examples/test_gpu.jl:21-24: `psi_gpu = CuArray(psi0); gradients(H, psi_gpu, circuit)`).
Storage is a torch CUDA tensor (torch = allocator + stream plumbing) wrapped by a
`qcb_state` handle; all arithmetic happens in libqcb200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_NP2QCB = {np.dtype(np.complex64): _lib.QCB_C64, np.dtype(np.complex128): _lib.QCB_C128}


def _nqubits_of(arr):
    size = int(np.prod(arr.shape)) if hasattr(arr, "shape") else int(arr)
    n = size.bit_length() - 1
    if size != 1 << n or n < 1:
        raise ValueError("a state needs 2^n amplitudes, n >= 1")
    return n


def zero_state_vec(*args):
    """zero_state_vec([type=ComplexF64], n) -- src/utils.jl:47-53 (host array, like the reference)."""
    dtype, n = (np.complex128, args[0]) if len(args) == 1 else args
    assert n >= 1, "Must have at least 1 qubit in the initial state"
    psi = np.zeros(2 ** n, dtype=dtype)
    psi[0] = 1
    return psi


def zero_state_tensor(*args):
    """zero_state_tensor([type=ComplexF64], n) -- src/utils.jl:61-66.

    Returned as a column-major 2x2x...x2 NumPy array so that `psi.reshape(-1, order="F")` is the
    reference's linear index (qubit 1 = least significant bit)."""
    dtype, n = (np.complex128, args[0]) if len(args) == 1 else args
    assert n >= 1
    psi = np.zeros((2,) * n, dtype=dtype, order="F")
    psi[(0,) * n] = 1
    return psi


def _flat(host):
    host = np.asarray(host)
    if host.ndim > 1:
        host = host.reshape(-1, order="F")  # Julia linear index
    return np.ascontiguousarray(host)


class B200State:
    """2^n amplitudes in HBM.  `B200State(psi_host)` uploads, `.to_numpy()` downloads."""

    def __init__(self, src, dtype=None, _empty=False):
        import torch

        _lib.init()
        self._handle = C.c_void_p()
        if isinstance(src, B200State):
            n, npdt = src.nqubits, src.dtype
        elif isinstance(src, (int, np.integer)):
            n, npdt = int(src), np.dtype(dtype or np.complex128)
        else:
            src_arr = np.asarray(src)
            n = _nqubits_of(src_arr)
            npdt = np.dtype(dtype) if dtype is not None else src_arr.dtype
        if npdt not in _NP2QCB:
            raise TypeError(f"state dtype must be complex64 or complex128, got {npdt}")
        self.nqubits, self.dtype = n, npdt
        self.tensor = torch.empty(1 << n, dtype=torch.complex64 if npdt == np.complex64 else torch.complex128,
                                  device=torch.device("cuda", torch.cuda.current_device()))
        _lib.check(_lib.lib().qcb_state_wrap(C.c_int(n), C.c_int(_NP2QCB[npdt]), C.c_void_p(self.tensor.data_ptr()),
                                             C.byref(self._handle)))
        if _empty:
            return
        if isinstance(src, B200State):
            _lib.check(_lib.lib().qcb_state_copy(self._handle, src._handle))
        elif isinstance(src, (int, np.integer)):
            _lib.check(_lib.lib().qcb_state_set_zero_state(self._handle))
        else:
            host = _flat(src).astype(npdt, copy=False)
            _lib.check(_lib.lib().qcb_state_upload(self._handle, host.ctypes.data_as(C.c_void_p)))

    # -- the array operations the reference uses on states --
    def similar(self):
        return B200State(self.nqubits, self.dtype, _empty=True)

    def copy(self):
        return B200State(self)

    def assign(self, other):
        """psi .= other"""
        if isinstance(other, B200State):
            _lib.check(_lib.lib().qcb_state_copy(self._handle, other._handle))
        else:
            host = _flat(other).astype(self.dtype, copy=False)
            if host.size != 1 << self.nqubits:
                raise ValueError("ψ′ must be the same size as ψ.")
            _lib.check(_lib.lib().qcb_state_upload(self._handle, host.ctypes.data_as(C.c_void_p)))
        return self

    def to_numpy(self, tensor_shape=False):
        out = np.empty(1 << self.nqubits, dtype=self.dtype)
        _lib.check(_lib.lib().qcb_state_download(self._handle, out.ctypes.data_as(C.c_void_p)))
        return out.reshape((2,) * self.nqubits, order="F") if tensor_shape else out

    def norm2(self):
        out = C.c_double(0)
        _lib.check(_lib.lib().qcb_state_norm2(self._handle, C.byref(out)))
        return out.value

    @property
    def shape(self):
        return (2,) * self.nqubits

    @property
    def ndim(self):
        return self.nqubits

    def __len__(self):
        return 1 << self.nqubits

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().qcb_state_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass
