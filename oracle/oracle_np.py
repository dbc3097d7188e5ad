"""NumPy mirror of the oracle + the dense checkers the reference's own tests use.

TEST INFRASTRUCTURE ONLY -- see oracle/qc_oracle.c for the rules.  Nothing under
quantumcircuits.jl_b200/ imports this module.

Parity pin: Julia is not installed here, the reference cannot be run and holds no stored
vectors; this module re-implements the *independent* definitions the reference tests
compare against (dense Kronecker operators, dense Hamiltonians) so that the C oracle and
the CUDA path are checked the way test/correctness_tests.jl checks `apply`/`measure`.

All citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------- gates (src/gates.jl)
HADAMARD = np.array([[1, 1], [1, -1]], dtype=np.float64) / np.sqrt(2)  # :10
XGATE = np.array([[0, 1], [1, 0]], dtype=np.float64)  # :11
ZGATE = np.array([[1, 0], [0, -1]], dtype=np.float64)  # :12
IGATE = np.eye(2)  # :13
NGATE = np.array([[1, 0], [0, 0]], dtype=np.float64)  # :14  |0><0|
# 4x4 forms, qubit K = low bit (src/gates.jl:15-16,33-34)
CNOT4 = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.float64)
RCNOT4 = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.float64)


def rz(a):  # src/gates.jl:36-43
    c, s = np.cos(a / 2), np.sin(a / 2)
    return np.array([[c - 1j * s, 0], [0, c + 1j * s]])


def ry(a):  # src/gates.jl:44-51
    c, s = np.cos(a / 2), np.sin(a / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def rx(a):  # src/gates.jl:52-59
    c, s = np.cos(a / 2), -1j * np.sin(a / 2)
    return np.array([[c, s], [s, c]])


def rotation(theta, phi, lam):  # src/gates.jl:61-73
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    p1, p2 = np.exp(1j * lam), np.exp(1j * phi)
    return np.array([[c, -p1 * s], [p2 * s, c * (p1 * p2)]])


def build_general_unitary_gate(angles):
    """src/gates.jl:80-109 -> 4x4 matrix M[i+2j, a+2b] (== the 2x2x2x2 array, column-major)."""
    a = np.asarray(angles, dtype=np.float64)
    assert a.shape == (15,)
    r1, r2, r3, r4 = (rotation(*a[3 * i:3 * i + 3]) for i in range(4))
    theta, phi, lam = a[12], a[13], a[14]
    rzm = rz(-np.pi / 2)
    l1 = np.kron(rzm @ r4, r3)
    l2 = np.kron(ry(phi), rz(theta))
    l3 = np.kron(ry(lam), IGATE)
    l4 = np.kron(r2, r1 @ rzm)
    return ((l4 @ RCNOT4) @ (l3 @ CNOT4)) @ ((l2 @ RCNOT4) @ l1)


# ----------------------------------------------------- geometry (src/circuits.jl:9-12)
def circuit_layer_starts(layer, nbits):
    """1-based layer -> 1-based first qubits of its gates."""
    return list(range(1 + (layer - 1) % 2, nbits, 2))


def brickwork_num_gates(nbits, nlayers):
    return sum(len(circuit_layer_starts(l, nbits)) for l in range(1, nlayers + 1))


def brickwork_gate_positions(nbits, nlayers):
    return [j for l in range(1, nlayers + 1) for j in circuit_layer_starts(l, nbits)]


# ---------------------------------------------- dense checkers (src/utils.jl:10-40)
def convert_gates_to_matrix(nbits, gates):
    """gates: list of (K, matrix) with matrix 2x2 (1-qubit) or 4x4 (2-qubit on K, K+1),
    non-overlapping.  kron(M_last, ..., M_first): first qubit = least-significant factor."""
    by_pos = {K: m for K, m in gates}
    mats = []
    i = 1
    while i <= nbits:
        if i in by_pos:
            m = np.asarray(by_pos[i])
            mats.append(m)
            i += 1 if m.shape == (2, 2) else 2
        else:
            mats.append(IGATE)
            i += 1
    out = mats[-1]
    for m in reversed(mats[:-1]):
        out = np.kron(out, m)
    return out


def dense_tfim(n, J, h, g):
    """examples/matrix_tfim.jl:7-32 : -J (sum ZZ + g sum X + h sum Z), open chain."""
    H = np.zeros((2 ** n, 2 ** n))
    for i in range(1, n):
        H += convert_gates_to_matrix(n, [(i, ZGATE), (i + 1, ZGATE)])
    for i in range(1, n + 1):
        H += g * convert_gates_to_matrix(n, [(i, XGATE)])
        H += h * convert_gates_to_matrix(n, [(i, ZGATE)])
    return -J * H


def east_params(c, s):
    """src/hamiltonians/east_model.jl:23-27"""
    return -np.exp(-s) * np.sqrt(c * (1 - c)), 1 - 2 * c


def dense_east(n, c, s):
    """src/hamiltonians/east_model.jl:57-81"""
    A, B = east_params(c, s)
    H = convert_gates_to_matrix(n, [(1, XGATE)]).astype(np.float64)
    for j in range(2, n + 1):
        H = H + convert_gates_to_matrix(n, [(j - 1, NGATE), (j, XGATE)])
    H = H * A
    H = H + B * convert_gates_to_matrix(n, [(1, NGATE)])
    for j in range(2, n + 1):
        m1 = convert_gates_to_matrix(n, [(j - 1, NGATE), (j, NGATE)])
        m2 = convert_gates_to_matrix(n, [(j - 1, NGATE)])
        H = H + B * (m1 + c * m2)
    return H + c * np.eye(2 ** n)


def measure_matrix(H, psi):
    """src/circuits.jl:49-55 : real(dot(psi, H, psi)); asserts a tiny imaginary part."""
    v = np.asarray(psi).reshape(-1)
    m = np.vdot(v, H @ v)
    assert abs(m.imag) < 1e-9
    return float(m.real)


# ------------------------------------------------ state-vector restatements (small n)
def apply_1q(psi, u, K):
    """src/apply.jl:20-38, vectorised: contraction on axis K (1-based, bit K-1)."""
    n = int(np.log2(psi.size))
    t = psi.reshape((2,) * n)  # C-order axes: axis 0 = highest bit
    ax = n - K
    out = np.tensordot(np.asarray(u), t, axes=([1], [ax]))
    return np.moveaxis(out, 0, ax).reshape(-1)


def apply_2q(psi, M, K):
    """src/apply.jl:62-84, vectorised.  M is the 4x4 form (low bit = qubit K)."""
    n = int(np.log2(psi.size))
    lo = 2 ** (K - 1)
    v = psi.reshape(-1, 4, lo)  # [post, (b a), pre]  with index = a + 2b
    return np.einsum("rc,pcq->prq", np.asarray(M), v).reshape(-1)


def right_apply_gate(M, u4, K):
    """src/apply.jl:144-182, literal: M' [pre, i, j, post] += M[pre, a, b, post] * u[a, b, i, j] on the COLUMN-index
    dims X = K + nbits, X + 1 of the 2^n x 2^n matrix seen as a 2n-dim tensor (column-major: the first n dims are the
    row-index bits).  u4 is the 4x4 form of the gate (row = i + 2j, col = a + 2b)."""
    M = np.asarray(M)
    n = int(np.log2(M.shape[0]))
    assert M.shape == (2 ** n, 2 ** n), "Must be a square matrix with a power of two edge"
    u = np.asarray(u4).reshape(2, 2, 2, 2, order="F")  # u[i, j, a, b]
    T = M.reshape((2,) * (2 * n), order="F")           # dims 0..n-1: row bits, n..2n-1: column bits
    X = K + n - 1                                       # 0-based dim of column qubit K
    out = np.zeros_like(T, dtype=np.result_type(M.dtype, u.dtype))
    for a in range(2):
        for b in range(2):
            src = [slice(None)] * (2 * n)
            src[X], src[X + 1] = a, b
            m = T[tuple(src)]
            for i in range(2):
                for j in range(2):
                    dst = [slice(None)] * (2 * n)
                    dst[X], dst[X + 1] = i, j
                    out[tuple(dst)] += m * u[a, b, i, j]   # "u is reversed": indexed [contract..., i, j]
    return out.reshape(2 ** n, 2 ** n, order="F")


def apply_brickwork(psi, nbits, nlayers, angles):
    """src/circuits.jl:19-31 ; angles is (15, G) as in the reference."""
    angles = np.asarray(angles).reshape(15, -1, order="F") if np.ndim(angles) == 1 else np.asarray(angles)
    for g, K in enumerate(brickwork_gate_positions(nbits, nlayers)):
        psi = apply_2q(psi, build_general_unitary_gate(angles[:, g]), K)
    return psi


def measure_tfim(psi, J, h, g):
    """src/hamiltonians/tfim.jl:17-54 summed (hamiltonians.jl:41), in complex128."""
    n = int(np.log2(psi.size))
    C = np.arange(psi.size, dtype=np.int64)
    zz = np.zeros(psi.size)
    for i in range(n - 1):
        r = (C >> i) & 3
        zz += np.where((r == 0) | (r == 3), 1.0, -1.0)
    ones = np.zeros(psi.size)
    x = np.zeros(psi.size, dtype=np.complex128)
    for i in range(n):
        ones += (C >> i) & 1
        x += np.conj(psi[C ^ (1 << i)])
    p2 = np.abs(psi) ** 2
    dens = -J * (g * x * psi + h * (n - 2 * ones) * p2 + zz * p2)
    return complex(dens.sum())


def measure_east(psi, c, A, B):
    """src/hamiltonians/east_model.jl:30-52 summed."""
    n = int(np.log2(psi.size))
    C = np.arange(psi.size, dtype=np.int64)
    nbit = lambda i: 1.0 - ((C >> i) & 1)  # n_i = 1 iff bit i is 0
    Ac = np.conj(psi[C ^ 1])
    Bc = nbit(0).astype(np.float64)
    for i in range(1, n):
        Ac = Ac + nbit(i - 1) * np.conj(psi[C ^ (1 << i)])
        Bc = Bc + nbit(i - 1) * nbit(i) + c * nbit(i - 1)
    dens = (A * Ac + (B * Bc + c) * np.conj(psi)) * psi
    return complex(dens.sum())


def naive_shift_gradients(nbits, nlayers, angles, psi0, energy_fn):
    """examples/test_gradient_accuracy.jl:25-39 : per-angle reconstruct + measure."""
    angles = np.array(angles, dtype=np.float64, order="F")
    grads = np.zeros_like(angles)
    for g in range(angles.shape[1]):
        for k in range(15):
            a = angles.copy(order="F")
            a[k, g] += np.pi / 2
            ep = energy_fn(apply_brickwork(psi0, nbits, nlayers, a))
            a[k, g] -= np.pi
            em = energy_fn(apply_brickwork(psi0, nbits, nlayers, a))
            grads[k, g] = (ep - em) / 2
    return grads


def zero_state(n, dtype=np.complex128):
    """src/utils.jl:47-66"""
    psi = np.zeros(2 ** n, dtype=dtype)
    psi[0] = 1
    return psi


# ------------------------------------------------ statevector polar optimiser
# ---- ext/MPSExt/polar_optimise.jl restated LITERALLY on 2x2x...x2 tensors (checker of the flat-vector restatement below) ----
def mps_contract(A, B, dimsA, dimsB, free_first="A"):
    """MatrixProductStates.contract(A, B, dimsA, dimsB) (1-based dims).  The package is not vendored; its convention is
    pinned by the reference's own call sites: every `contract` is followed by `permutedims(., perm_idxs(site, N))`
    (polar_optimise.jl:31-33,63-65,85-86,97-98), which puts result dims 1, 2 at (site, site+1) and result dims 3.. at the
    other positions in order.  That restores the qubit order -- so that e.g. an all-identity circuit maps a state to
    itself -- only if the free dims of A come first, then the free dims of B in their original order (numpy.tensordot's
    convention).  free_first="B" is the other conceivable reading, kept so that the test can show it is inconsistent."""
    out = np.tensordot(A, B, axes=([d - 1 for d in dimsA], [d - 1 for d in dimsB]))
    if free_first == "B":
        na = A.ndim - len(dimsA)
        out = np.moveaxis(out, list(range(na)), list(range(out.ndim - na, out.ndim)))
    return out


def perm_idxs(site, N):
    """polar_optimise.jl:5-13 (1-based)."""
    perm = list(range(1, N + 1))
    for i in range(1, site):
        perm[i - 1] = i + 2
    perm[site - 1] = 1
    perm[site] = 2
    return perm


def _permutedims(A, perm):
    """Julia's permutedims(A, perm): result dim i is A's dim perm[i]."""
    return np.transpose(A, [p - 1 for p in perm])


def _gate_tensor(g):
    g = np.asarray(g, dtype=np.complex128)
    return g if g.ndim == 4 else g.reshape(2, 2, 2, 2, order="F")


def circuit_to_state_literal(circuit, N, free_first="A"):
    """polar_optimise.jl:22-46, statement by statement; returns the flat column-major state."""
    psi = np.zeros(2 ** N, dtype=np.complex128)
    psi[0] = 1
    psi = psi.reshape((2,) * N, order="F")
    for layer in circuit:
        for idx in list(range(1, N // 2 + 1)) + list(range(N // 2 + 1, N)):
            site = 2 * idx - 1 if idx <= N // 2 else 2 * idx - N
            psi = mps_contract(_gate_tensor(layer[idx - 1]), psi, [3, 4], [site, site + 1], free_first)
            psi = _permutedims(psi, perm_idxs(site, N))
    return psi.reshape(-1, order="F")


def create_upper_state_literal(circuit, phi, N, free_first="A"):
    """polar_optimise.jl:52-78, statement by statement."""
    psi = np.conj(np.asarray(phi, dtype=np.complex128)).reshape((2,) * N, order="F")
    for layer in reversed(circuit):
        for idx in list(range(N - 1, N // 2, -1)) + list(range(N // 2, 0, -1)):
            site = 2 * idx - N if idx > N // 2 else 2 * idx - 1
            psi = mps_contract(_gate_tensor(layer[idx - 1]), psi, [1, 2], [site, site + 1], free_first)
            psi = _permutedims(psi, perm_idxs(site, N))
    return psi.reshape(-1, order="F")


def polar_optimise_literal(circuit, psi_GS, energy_fn, N, iterations=100):
    """polar_optimise.jl:83-147 on tensors with `mps_contract` / `permutedims`, statement by statement (slow: checker only)."""
    circuit = [[_gate_tensor(g).copy() for g in layer] for layer in circuit]
    overlaps, energies = [], []
    gs = np.asarray(psi_GS, dtype=np.complex128).reshape(-1)
    for _ in range(iterations):
        lower = np.zeros(2 ** N, dtype=np.complex128)
        lower[0] = 1
        lower = lower.reshape((2,) * N, order="F")
        upper = create_upper_state_literal(circuit, gs, N).reshape((2,) * N, order="F")
        for layer in circuit:
            for idx in range(1, N):
                site = 2 * idx - 1 if idx <= N // 2 else 2 * idx - N
                upper = _permutedims(mps_contract(np.conj(layer[idx - 1]), upper, [3, 4], [site, site + 1]), perm_idxs(site, N))
                cd = [i for i in range(1, N + 1) if i not in (site, site + 1)]
                E = mps_contract(upper, lower, cd, cd).reshape(4, 4, order="F")
                U_, _, Vt = np.linalg.svd(E)
                U = np.conj(U_ @ Vt).reshape(2, 2, 2, 2, order="F")
                layer[idx - 1] = U
                lower = _permutedims(mps_contract(U, lower, [3, 4], [site, site + 1]), perm_idxs(site, N))
        flat = lower.reshape(-1, order="F")
        overlaps.append(abs(np.vdot(gs, flat)))
        energies.append(energy_fn(flat))
    return [[g.reshape(4, 4, order="F") for g in layer] for layer in circuit], np.array(overlaps), np.array(energies)


def polar_optimise(circuit, psi_GS, energy_fn, N, iterations=100):
    """ext/MPSExt/polar_optimise.jl:105-147 restated on flat state vectors.

    circuit: list of brick layers, each a list of N-1 gates given as 4x4 matrices M[i+2j, a+2b] (the reshape(.,4,4) of
    the reference's 2x2x2x2 arrays); gates 1..N/2 of a layer sit on sites 1,3,5,.. and gates N/2+1..N-1 on sites 2,4,..
    (:124-137).  `MatrixProductStates.contract(gate, psi, [3,4], [site,site+1])` + `permutedims(perm_idxs)` (:31-33) is
    the ordinary application of the gate on (site, site+1); with [1,2] (:63-65) it is the application of the transpose.
    NOTE: MatrixProductStates.jl is not vendored in the reference tree; the reading of `contract` (contract the listed
    dimensions, free dimensions of the first argument first) is the only one consistent with the `permutedims` that
    follows every call (see mps_contract; tests/test_oracle.py checks this function against the statement-by-statement
    tensor form `polar_optimise_literal`, and that the other reading breaks identity circuits).  Returns (circuit, overlaps, energies)."""
    assert N % 2 == 0
    circuit = [[np.array(g, dtype=np.complex128) for g in layer] for layer in circuit]
    sites = [2 * idx - 1 for idx in range(1, N // 2 + 1)] + [2 * idx - N for idx in range(N // 2 + 1, N)]
    overlaps, energies = [], []
    for _ in range(iterations):
        lower = zero_state(N)
        upper = np.conj(np.asarray(psi_GS, dtype=np.complex128).reshape(-1))  # create_upper_state :52-78
        for layer in reversed(circuit):
            for idx in range(N - 1, 0, -1):
                upper = apply_2q(upper, layer[idx - 1].T, sites[idx - 1])
        for layer in circuit:
            for idx in range(1, N):
                site = sites[idx - 1]
                upper = apply_2q(upper, np.conj(layer[idx - 1]), site)  # :85-86
                lo = 2 ** (site - 1)
                u3 = upper.reshape(-1, 4, lo)
                l3 = lower.reshape(-1, 4, lo)
                E = np.einsum("prq,pcq->rc", u3, l3)  # :88-92  E[(i,j),(a,b)] = sum_rest upper * lower
                U_, _, Vt = np.linalg.svd(E)
                new = np.conj(U_ @ Vt)  # :94
                layer[idx - 1] = new
                lower = apply_2q(lower, new, site)  # :97-98
        overlaps.append(abs(np.vdot(np.asarray(psi_GS).reshape(-1), lower)))  # :140
        energies.append(energy_fn(lower))  # :141
    return circuit, np.array(overlaps), np.array(energies)
