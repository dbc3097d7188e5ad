"""Sharded statevectors: one process per GPU, amplitudes split by the top log2(P) qubits.

No reference counterpart (the reference keeps one array in one memory).  torch.distributed is
only the bootstrap channel (it ships the 128-byte NCCL id and sums scalars); the qubit swaps are NCCL
all-to-alls issued by libqcb200.so itself (csrc/dist.cu).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .state import _NP2QCB

_dist_ready = False


def init_from_torch():
    """Every rank of an initialised torch.distributed job calls this once."""
    global _dist_ready
    import torch
    import torch.distributed as dist

    if _dist_ready:
        return
    _lib.init(int(os.environ.get("LOCAL_RANK", "0")))
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _lib.check(_lib.lib().qcb_dist_unique_id(uid.ctypes.data_as(C.c_void_p)))
    box = [uid.tobytes()]
    dist.broadcast_object_list(box, src=0)
    uid = np.frombuffer(box[0], dtype=np.uint8).copy()
    _lib.check(_lib.lib().qcb_dist_init(C.c_int(rank), C.c_int(world), uid.ctypes.data_as(C.c_void_p)))
    torch.cuda.synchronize()
    _dist_ready = True


def finalize():
    global _dist_ready
    if _dist_ready:
        _lib.check(_lib.lib().qcb_dist_finalize())
        _dist_ready = False


class ShardedState:
    """2^n amplitudes over all ranks; this rank holds the canonical slice [rank * 2^nloc, (rank+1) * 2^nloc)
    when `canonical` (after upload / download); between swaps the library tracks the permuted layout."""

    def __init__(self, nqubits, dtype=np.complex128):
        self.nqubits, self.dtype = int(nqubits), np.dtype(dtype)
        self._handle = C.c_void_p()
        _lib.check(_lib.lib().qcb_state_create_sharded(C.c_int(self.nqubits), C.c_int(_NP2QCB[self.dtype]), C.byref(self._handle)))
        nloc = C.c_int(0)
        _lib.check(_lib.lib().qcb_state_local_qubits(self._handle, C.byref(nloc)))
        self.local_qubits = nloc.value
        self.peer_swap_mapped = False
        self.peers_mapped = self._map_peers()
        self.set_zero_state()

    def _map_peers(self):
        """Exchange CUDA IPC handles so that qubit swaps can be fused into gate passes as NVLink peer loads."""
        import torch.distributed as dist

        mine = np.zeros(128, dtype=np.uint8)
        rc = _lib.lib().qcb_state_ipc_export(self._handle, mine.ctypes.data_as(C.c_void_p))
        table = [None] * dist.get_world_size()
        dist.all_gather_object(table, mine.tobytes() if rc == 0 else bytes(128))
        if rc != 0 or any(t[:64] == bytes(64) for t in table):
            return False  # export failed somewhere: swaps stay on NCCL
        blob = np.frombuffer(b"".join(table), dtype=np.uint8).copy()
        ok = _lib.lib().qcb_state_ipc_import(self._handle, blob.ctypes.data_as(C.c_void_p)) == 0
        flags = [None] * dist.get_world_size()
        dist.all_gather_object(flags, ok)
        if not all(flags):
            _lib.lib().qcb_state_ipc_close(self._handle)
            return False
        # second buffers everywhere: swaps ride on gate passes (peer loads); nowhere: in-place peer swaps (k_swap_peer);
        # mixed: nothing was mapped
        self.peer_swap_mapped = all(t[64:] == bytes(64) for t in table)
        return all(t[64:] != bytes(64) for t in table)

    def close(self):
        """Collective: unmap peers on every rank, synchronise, free."""
        import torch
        import torch.distributed as dist

        if self._handle:
            torch.cuda.synchronize()
            _lib.lib().qcb_state_ipc_close(self._handle)
            dist.barrier()
            _lib.lib().qcb_state_destroy(self._handle)
            self._handle = C.c_void_p()

    def set_zero_state(self):
        _lib.check(_lib.lib().qcb_state_set_zero_state(self._handle))

    def upload_shard(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == 1 << self.local_qubits
        _lib.check(_lib.lib().qcb_state_upload_shard(self._handle, host.ctypes.data_as(C.c_void_p)))

    def download_shard(self, out=None):
        """Canonical slice of this rank; `out` may be a (pinned) host array to receive it without an extra copy."""
        if out is None:
            out = np.empty(1 << self.local_qubits, dtype=self.dtype)
        assert out.dtype == self.dtype and out.size == 1 << self.local_qubits and out.flags.c_contiguous
        _lib.check(_lib.lib().qcb_state_download_shard(self._handle, out.ctypes.data_as(C.c_void_p)))
        return out

    def apply_circuit(self, circuit, first=0, last=None, adjoint=False):
        last = circuit.ngates if last is None else last
        a = circuit.angles_abi()
        _lib.check(_lib.lib().qcb_apply_brickwork_range(self._handle, C.c_int(circuit.nbits), C.c_int(circuit.nlayers),
                                                        a.ctypes.data_as(C.c_void_p), C.c_int(first), C.c_int(last),
                                                        C.c_int(1 if adjoint else 0)))

    def apply_gates(self, gates):
        from .apply import _apply_list_inplace

        _apply_list_inplace(self, list(gates))

    def measure(self, H):
        """measure(H, psi) on the sharded state (collective; every rank gets the value)."""
        from .hamiltonians import _measure_state

        return _measure_state(H, self)

    def gradients(self, H, circuit, calculate_energy=False):
        """gradients(H, psi, circuit; calculate_energy) on the sharded state (collective; every rank gets E and the full
        15 x G gradient): forward circuit with qubit swaps, lambda = H psi in two sweeps around one swap, backward sweep on
        the (psi, lambda) pair, one all-reduce."""
        from .gradients import gradients

        return gradients(H, self, circuit, calculate_energy=calculate_energy)

    def norm2(self):
        import torch
        import torch.distributed as dist

        out = C.c_double(0)
        _lib.check(_lib.lib().qcb_state_norm2(self._handle, C.byref(out)))
        t = torch.tensor([out.value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().qcb_state_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass
