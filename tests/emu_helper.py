"""ctypes access to tests/emu/libqcb_emu.so -- CPU replay of the device tile-pass code."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_SO = os.path.join(_HERE, "libqcb_emu.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "..", "quantumcircuits.jl_b200", "csrc")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "emu_tile.cu"), os.path.join(_CSRC, "tile.cuh"), os.path.join(_CSRC, "schedule.hpp"), os.path.join(_CSRC, "dist_plan.hpp"), os.path.join(_CSRC, "adjoint.cuh"), os.path.join(_CSRC, "adjoint_mma.cuh"), os.path.join(_CSRC, "adjoint_f32.cuh"), os.path.join(_CSRC, "gates_hd.cuh"), os.path.join(_CSRC, "measure_tile.cuh"), os.path.join(_CSRC, "measure_plan.hpp"), os.path.join(_CSRC, "kernels.cuh"), os.path.join(_CSRC, "flip_tile.cuh"), os.path.join(_CSRC, "tma.cuh"), os.path.join(_CSRC, "tile_tma.cuh"), os.path.join(_CSRC, "tile_async.cuh"), os.path.join(_CSRC, "small_circuit.cuh")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                               "-fPIC", "-shared", "-o", _SO, srcs[0]], cwd=_HERE)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def set_m3(on):
    """ComplexF64 arithmetic of the replayed gate passes: True = 3M form + straight-line full diamonds (default), False = ordinary."""
    lib().emu_set_m3(C.c_int(1 if on else 0))


def set_async(grid):
    """grid > 0: replay the persistent gate pass with asynchronous tile loads (tile_async.cuh) on that many CTAs; 0: one tile per CTA."""
    lib().emu_set_async(C.c_int(int(grid)))


def apply_gates(psi, q0, mats, tile_bits, low_bits, phys=None, nloc=None, max_gates=0, adjoint=False):
    """psi: flat complex array (modified copy returned); q0: 0-based logical low qubits; mats: (G,4,4)."""
    psi = np.ascontiguousarray(psi).copy()
    n = int(psi.size).bit_length() - 1
    nloc = n if nloc is None else nloc
    phys = np.arange(n, dtype=np.int32) if phys is None else np.ascontiguousarray(phys, dtype=np.int32)
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    stats = np.zeros(4, dtype=np.int32)
    rc = lib().emu_apply_gates(C.c_int(1 if psi.dtype == np.complex128 else 0), C.c_int(nloc), C.c_int(len(phys)),
                               phys.ctypes.data_as(C.c_void_p), psi.ctypes.data_as(C.c_void_p), C.c_int(len(q0)),
                               q0.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_int(tile_bits),
                               C.c_int(low_bits), C.c_int(max_gates), C.c_int(1 if adjoint else 0),
                               stats.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_apply_gates rc={rc}")
    return psi, (int(stats[0]), int(stats[1]))


def apply_gates_tma(psi, q0, mats, low_bits=3, max_gates=0, adjoint=False):
    """The TMA-fed gate pass (tile_tma.cuh) on a canonical ComplexF64 state with emulated box loads / stores."""
    psi = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    n = int(psi.size).bit_length() - 1
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    stats = np.zeros(4, dtype=np.int32)
    rc = lib().emu_apply_gates_tma(C.c_int(n), psi.ctypes.data_as(C.c_void_p), C.c_int(len(q0)), q0.ctypes.data_as(C.c_void_p),
                                   m.ctypes.data_as(C.c_void_p), C.c_int(low_bits), C.c_int(max_gates), C.c_int(1 if adjoint else 0),
                                   stats.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_apply_gates_tma rc={rc}")
    return psi, (int(stats[0]), int(stats[1]))


def plan_brickwork(nbits, nlayers, tile_bits, low_bits, max_gates=0):
    stats = np.zeros(4, dtype=np.int32)
    lib().emu_plan_brickwork(C.c_int(nbits), C.c_int(nlayers), C.c_int(tile_bits), C.c_int(low_bits), C.c_int(max_gates),
                             stats.ctypes.data_as(C.c_void_p))
    return dict(passes=int(stats[0]), stages=int(stats[1]), gates=int(stats[2]))


def sharded_apply_gates(psi, g, q0, mats, tile_bits, low_bits, adjoint=False, canonical_out=True, fuse_swaps=False):
    """Simulated 2^g ranks.  psi: full canonical complex128 state.  Returns (state, phys, stats)."""
    psi = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    n = int(psi.size).bit_length() - 1
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    stats = np.zeros(8, dtype=np.int32)
    phys = np.zeros(n, dtype=np.int32)
    rc = lib().emu_sharded_apply_gates(C.c_int(n), C.c_int(g), psi.ctypes.data_as(C.c_void_p), C.c_int(len(q0)),
                                       q0.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_int(tile_bits),
                                       C.c_int(low_bits), C.c_int(1 if adjoint else 0), C.c_int((1 if canonical_out else 0) | (2 if fuse_swaps else 0)),
                                       phys.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_sharded_apply_gates rc={rc}")
    return psi, phys, dict(gate_passes=int(stats[0]), stages=int(stats[1]), swaps=int(stats[2]), passes=int(stats[3]), fused=int(stats[4]))


def plan_sharded_brickwork(nbits, nlayers, g, tile_bits, low_bits):
    stats = np.zeros(8, dtype=np.int32)
    rc = lib().emu_plan_sharded_brickwork(C.c_int(nbits), C.c_int(nlayers), C.c_int(g), C.c_int(tile_bits), C.c_int(low_bits),
                                          stats.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_plan_sharded_brickwork rc={rc}")
    return dict(passes=int(stats[0]), stages=int(stats[1]), swaps=int(stats[2]), canon_passes=int(stats[3]), permute_passes=int(stats[4]))


def adjoint_sweep(psi, lam, q0, mats, tile_bits, low_bits):
    """Backward sweep replay: returns (psi_0, lam_0, env[G, 4, 4]) with env[g][r, c] = sum conj(lam_g[r]) psi_{g-1}[c]."""
    psi = np.ascontiguousarray(psi).copy()
    lam = np.ascontiguousarray(lam).astype(psi.dtype).copy()
    n = int(psi.size).bit_length() - 1
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    env = np.zeros((len(q0), 32), dtype=np.float64)
    rc = lib().emu_adjoint_sweep(C.c_int(1 if psi.dtype == np.complex128 else 0), C.c_int(n), psi.ctypes.data_as(C.c_void_p),
                                 lam.ctypes.data_as(C.c_void_p), C.c_int(len(q0)), q0.ctypes.data_as(C.c_void_p),
                                 m.ctypes.data_as(C.c_void_p), C.c_int(tile_bits), C.c_int(low_bits), env.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_adjoint_sweep rc={rc}")
    envc = (env[:, 0::2] + 1j * env[:, 1::2]).reshape(len(q0), 4, 4).transpose(0, 2, 1)  # stored column-major r + 4c
    return psi, lam, envc


def adjoint_sweep_f32(psi, lam, q0, mats, tile_bits, low_bits):
    """ComplexF32 backward sweep with register windows and TF32-split reduction MMAs (adjoint_f32.cuh): returns
    (psi_0, lam_0, env[G, 4, 4])."""
    psi = np.ascontiguousarray(psi, dtype=np.complex64).copy()
    lam = np.ascontiguousarray(lam, dtype=np.complex64).copy()
    n = int(psi.size).bit_length() - 1
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    env = np.zeros((len(q0), 32), dtype=np.float64)
    rc = lib().emu_adjoint_sweep_f32(C.c_int(n), psi.ctypes.data_as(C.c_void_p), lam.ctypes.data_as(C.c_void_p), C.c_int(len(q0)),
                                     q0.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_int(tile_bits), C.c_int(low_bits),
                                     env.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_adjoint_sweep_f32 rc={rc}")
    envc = (env[:, 0::2] + 1j * env[:, 1::2]).reshape(len(q0), 4, 4).transpose(0, 2, 1)
    return psi, lam, envc


def adjoint_sweep_mma(psi, lam, q0, mats, tile_bits, low_bits, nwarps):
    """Tensor-core form of the backward sweep (adjoint_mma.cuh), complex128 only.  Returns (psi_0, lam_0, env, (wavefronts,
    conflict-free wavefronts) of the stage loads/stores)."""
    psi = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    lam = np.ascontiguousarray(lam, dtype=np.complex128).copy()
    n = int(psi.size).bit_length() - 1
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    env = np.zeros((len(q0), 32), dtype=np.float64)
    conf = np.zeros(2, dtype=np.int64)
    rc = lib().emu_adjoint_sweep_mma(C.c_int(n), psi.ctypes.data_as(C.c_void_p), lam.ctypes.data_as(C.c_void_p), C.c_int(len(q0)),
                                     q0.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_int(tile_bits), C.c_int(low_bits),
                                     C.c_int(nwarps), env.ctypes.data_as(C.c_void_p), conf.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_adjoint_sweep_mma rc={rc}")
    envc = (env[:, 0::2] + 1j * env[:, 1::2]).reshape(len(q0), 4, 4).transpose(0, 2, 1)
    return psi, lam, envc, (int(conf[0]), int(conf[1]))


def sharded_gradient_env(psi0, g, kind, params, q0, mats, tile_bits, low_bits):
    """Sharded adjoint gradient with 2^g simulated ranks, from the full canonical input state psi_0 (forward circuit, lambda =
    H psi, backward sweep: all sharded): returns (env[G, 4, 4], E, stats)."""
    psi = np.ascontiguousarray(psi0, dtype=np.complex128)
    n = int(psi.size).bit_length() - 1
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    env = np.zeros((len(q0), 32), dtype=np.float64)
    E = np.zeros(2, dtype=np.float64)
    stats = np.zeros(4, dtype=np.int32)
    rc = lib().emu_sharded_gradient_env(C.c_int(n), C.c_int(g), psi.ctypes.data_as(C.c_void_p), C.c_int(kind), C.c_double(params[0]),
                                        C.c_double(params[1]), C.c_double(params[2]), C.c_int(len(q0)), q0.ctypes.data_as(C.c_void_p),
                                        m.ctypes.data_as(C.c_void_p), C.c_int(tile_bits), C.c_int(low_bits), env.ctypes.data_as(C.c_void_p),
                                        E.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_sharded_gradient_env rc={rc}")
    envc = (env[:, 0::2] + 1j * env[:, 1::2]).reshape(len(q0), 4, 4).transpose(0, 2, 1)
    return envc, complex(E[0], E[1]), dict(passes=int(stats[0]), swaps=int(stats[1]), canonical_after_forward=bool(stats[2]))


def measure(psi, kind, params, tile_bits, low_bits):
    """Tiled expectation value replay.  kind 0: TFIM (J, h, g); 1: East (c, A, B).  Returns (E, number of passes)."""
    psi = np.ascontiguousarray(psi)
    n = int(psi.size).bit_length() - 1
    E = C.c_double(0)
    npass = C.c_int(0)
    rc = lib().emu_measure(C.c_int(1 if psi.dtype == np.complex128 else 0), C.c_int(n), psi.ctypes.data_as(C.c_void_p), C.c_int(kind),
                           C.c_double(params[0]), C.c_double(params[1]), C.c_double(params[2]), C.c_int(tile_bits), C.c_int(low_bits),
                           C.byref(E), C.byref(npass))
    if rc:
        raise RuntimeError(f"emu_measure rc={rc}")
    return E.value, npass.value


def apply_ham_tiled(psi, kind, params, grid=3):
    """lambda = H psi with the flips of the low index bits served from a staged tile (kernels.cuh: k_apply_ham_tiled, the tile
    sizes the library launches: 2^11 ComplexF64 / 2^12 ComplexF32) on `grid` emulated CTAs.  Returns (lambda, <psi|lambda>)."""
    psi = np.ascontiguousarray(psi)
    n = int(psi.size).bit_length() - 1
    lam = np.zeros_like(psi)
    E2 = np.zeros(2)
    rc = lib().emu_apply_ham_tiled(C.c_int(1 if psi.dtype == np.complex128 else 0), C.c_int(n), psi.ctypes.data_as(C.c_void_p), C.c_int(kind),
                                   C.c_double(params[0]), C.c_double(params[1]), C.c_double(params[2]), C.c_int(grid),
                                   lam.ctypes.data_as(C.c_void_p), E2.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"emu_apply_ham_tiled rc={rc}")
    return lam, complex(E2[0], E2[1])


def flip_measure(psi, kind, params, grid=3, n_total=None, c_hi=0):
    """TMA-fed persistent expectation value (flip_tile.cuh) with emulated tile loads.  psi: the local amplitudes; for a
    simulated shard pass n_total and c_hi = rank << nloc.  Returns (diagonal sum, flip sum, levels); E = diag + coef * flips."""
    psi = np.ascontiguousarray(psi)
    nloc = int(psi.size).bit_length() - 1
    out = np.zeros(2)
    nl = C.c_int(0)
    rc = lib().emu_flip_measure(C.c_int(1 if psi.dtype == np.complex128 else 0), C.c_int(nloc), C.c_int(nloc if n_total is None else n_total),
                                psi.ctypes.data_as(C.c_void_p), C.c_int(kind), C.c_double(params[0]), C.c_double(params[1]), C.c_double(params[2]),
                                C.c_ulonglong(c_hi), C.c_int(grid), out.ctypes.data_as(C.c_void_p), C.byref(nl))
    if rc:
        raise RuntimeError(f"emu_flip_measure rc={rc}")
    return out[0], out[1], nl.value


def measure_persistent(psi, kind, params, tile_bits, low_bits, grid):
    """Persistent form of the tiled expectation value (grid CTAs walking the tiles)."""
    psi = np.ascontiguousarray(psi)
    n = int(psi.size).bit_length() - 1
    E = C.c_double(0)
    rc = lib().emu_measure_persistent(C.c_int(1 if psi.dtype == np.complex128 else 0), C.c_int(n), psi.ctypes.data_as(C.c_void_p), C.c_int(kind),
                                      C.c_double(params[0]), C.c_double(params[1]), C.c_double(params[2]), C.c_int(tile_bits), C.c_int(low_bits),
                                      C.c_int(grid), C.byref(E))
    if rc:
        raise RuntimeError(f"emu_measure_persistent rc={rc}")
    return E.value


def sharded_measure(psi, g, kind, params, tile_bits, low_bits):
    """Expectation value of a state split over 2^g simulated ranks (canonical layout)."""
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    n = int(psi.size).bit_length() - 1
    E = C.c_double(0)
    rc = lib().emu_sharded_measure(C.c_int(n), C.c_int(g), psi.ctypes.data_as(C.c_void_p), C.c_int(kind), C.c_double(params[0]),
                                   C.c_double(params[1]), C.c_double(params[2]), C.c_int(tile_bits), C.c_int(low_bits), C.byref(E))
    if rc:
        raise RuntimeError(f"emu_sharded_measure rc={rc}")
    return E.value


def contract_env_gate(angles15, env, post):
    """gates_hd.cuh (the algebra of k_contract_env) on the host: returns (out15, U as 4x4)."""
    a = np.ascontiguousarray(angles15, dtype=np.float64)
    e = np.ascontiguousarray(np.asarray(env, dtype=np.complex128).reshape(-1, order="F"))
    out = np.zeros(15)
    u = np.zeros(16, dtype=np.complex128)
    rc = lib().emu_contract_env_gate(a.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), C.c_int(1 if post else 0), out.ctypes.data_as(C.c_void_p),
                                     u.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError("per-layer and whole-gate contraction differ")
    return out, u.reshape(4, 4, order="F")


def small_circuit(psi, q0, mats, gbits=-1, adjoint=False):
    """Replay of the cluster-resident small-circuit kernel (small_circuit.cuh) with 2^gbits simulated CTAs (-1: the host's choice)."""
    psi = np.ascontiguousarray(psi).copy()
    n = int(psi.size).bit_length() - 1
    q0 = np.ascontiguousarray(q0, dtype=np.int32)
    m = np.ascontiguousarray(np.stack([np.asarray(M, dtype=np.complex128).reshape(-1, order="F") for M in mats]))
    rc = lib().emu_small_circuit(C.c_int(1 if psi.dtype == np.complex128 else 0), C.c_int(n), C.c_int(gbits), psi.ctypes.data_as(C.c_void_p),
                                 C.c_int(len(q0)), q0.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_int(1 if adjoint else 0))
    if rc:
        raise RuntimeError(f"emu_small_circuit rc={rc}")
    return psi
