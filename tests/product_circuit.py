"""Analytic reference for brickwork circuits whose gates do not entangle (TEST INFRASTRUCTURE).

With the three entangling angles (rows 13..15 of the reference's 15 angles) of a gate at zero, `build_general_unitary_gate`
(src/gates.jl:80-109) is  (r2 x r1 rz(-pi/2)) . SWAP . (rz(-pi/2) r4 x r3)  because RCNOT . CNOT . RCNOT = SWAP: a product
gate after a swap.  A brickwork circuit of such gates maps |0...0> to a product state whose 2-vectors are tracked here gate
by gate, so that the amplitude at any index is a product of n numbers -- a known answer at ANY size (30 qubits included)
that involves neither the oracle nor the device code."""
import numpy as np

from oracle import oracle_np as onp


def product_circuit_vectors(n, layers, angles):
    """angles: (15, G) with rows 12..14 zero.  Returns the list of n per-qubit 2-vectors (qubit q + 1 = index bit q)."""
    assert not np.any(angles[12:15])
    v = [np.array([1.0, 0.0], dtype=np.complex128) for _ in range(n)]
    rzm = onp.rz(-np.pi / 2)
    for g, K in enumerate(onp.brickwork_gate_positions(n, layers)):  # the gate acts on qubits K, K + 1 (1-based)
        r1, r2, r3, r4 = (onp.rotation(*angles[3 * i:3 * i + 3, g]) for i in range(4))
        lo, hi = v[K - 1], v[K]
        v[K - 1] = r1 @ rzm @ rzm @ r4 @ hi
        v[K] = r2 @ r3 @ lo
    return v


def product_amplitudes(v, idx):
    """amplitudes of the product state at the (integer array) indices idx"""
    idx = np.asarray(idx, dtype=np.int64)
    out = np.ones(idx.shape, dtype=np.complex128)
    for q, vq in enumerate(v):
        out *= vq[(idx >> q) & 1]
    return out
