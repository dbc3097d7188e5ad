"""ctypes binding of libqcb200.so (include/qcb200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
compute entry point fails with QCB_ECUDA when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# QCB200_LIB: another build of the same library (A/B timing of compile-time variants); the default is the in-tree build
SO_PATH = os.environ.get("QCB200_LIB") or os.path.join(_HERE, "libqcb200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "qcb200.h")

QCB_OK, QCB_EINVAL, QCB_ENOMEM, QCB_ECUDA, QCB_ENCCL, QCB_EUNSUPPORTED = range(6)
QCB_C64, QCB_C128 = 0, 1
QCB_HAM_TFIM, QCB_HAM_EAST = 0, 1


class QCBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libqcb200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class GateOutsideQubitSpace(AssertionError):
    """The reference's `@assert K < nqubits && K > 0` (src/apply.jl:27,45,112)."""


_lib = None


def declared_symbols():
    """Every function name include/qcb200.h declares."""
    with open(HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(qcb_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This backend has no CPU fallback.")
        # torch first so that its bundled libnccl/libcudart are the ones the process resolves
        try:
            import torch  # noqa: F401
        except Exception:  # pragma: no cover
            pass
        _lib = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
        _lib.qcb_last_error.restype = C.c_char_p
        _lib.qcb_version.restype = C.c_char_p
        _lib.qcb_brickwork_num_gates.restype = C.c_int
        _lib.qcb_set_option.argtypes = [C.c_char_p, C.c_long]
        _lib.qcb_get_stat.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
    return _lib


def check(rc):
    if rc != QCB_OK:
        msg = lib().qcb_last_error().decode("utf-8", "replace")
        if "outside the qubit space" in msg:
            raise GateOutsideQubitSpace(msg)
        if "must be the same size" in msg:  # the reference's error("ψ′ must be the same size as ψ.")  src/apply.jl:108
            raise ValueError(msg)
        raise QCBError(rc, msg)


_initialised_device = None


def init(device=None):
    """Bind this process to one GPU (LOCAL_RANK by default) and share torch's current stream."""
    global _initialised_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised_device == device:
        return
    L = lib()
    check(L.qcb_init(C.c_int(device)))
    try:
        import torch

        torch.cuda.set_device(device)
        check(L.qcb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    except ImportError:  # pragma: no cover
        pass
    _initialised_device = device


def set_option(key, value):
    check(lib().qcb_set_option(key.encode(), C.c_long(int(value))))


def get_stat(key):
    out = C.c_double(0)
    check(lib().qcb_get_stat(key.encode(), C.byref(out)))
    return out.value


def sync():
    check(lib().qcb_sync())


def trim():
    """Free the library's cached work buffers (gradient work states, host-call staging, partials)."""
    check(lib().qcb_trim())
