"""Pins the CPU oracle (oracle/qc_oracle.c + oracle/oracle_np.py) against every known-answer
and equivalence test the reference holds for the hot path.

The reference's tests draw unseeded randn; here the same properties are checked with fixed
seeds.  Citations are relative to /root/reference.
"""
import numpy as np
import pytest

from oracle import oracle_np as onp


def rand_state(rng, n, dtype=np.complex128):
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    return (v / np.linalg.norm(v)).astype(dtype)


def rand_nonunitary(rng, d):
    # test/correctness_tests.jl:57-65,90-98 : randn(ComplexF64, d, d) ./ det
    while True:
        u = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
        det = np.linalg.det(u)
        if det != 0:
            return u / det


# ---- test/correctness_tests.jl:3-27 "Test Cat state generation"
@pytest.mark.parametrize("nbits", [4, 5, 7])
@pytest.mark.parametrize("scatter", [True, False])
def test_cat_state(oracle, nbits, scatter):
    psi = onp.zero_state(nbits)
    psi = oracle.apply_1q(psi, onp.HADAMARD, 1)
    for i in range(1, nbits):
        psi = oracle.apply_2q(psi, onp.CNOT4, i, scatter=scatter)
    expected = np.zeros(2 ** nbits, dtype=np.complex128)
    expected[0] = expected[-1] = 1 / np.sqrt(2)
    np.testing.assert_allclose(psi, expected, atol=1e-15)


# ---- test/correctness_tests.jl:30-51 "Cat state ... with matrices" (pins the dense checker)
@pytest.mark.parametrize("nbits", [4, 5, 7])
def test_cat_state_matrices(nbits):
    psi = onp.zero_state(nbits)
    psi = onp.convert_gates_to_matrix(nbits, [(1, onp.HADAMARD)]) @ psi
    for i in range(1, nbits):
        psi = onp.convert_gates_to_matrix(nbits, [(i, onp.CNOT4)]) @ psi
    expected = np.zeros(2 ** nbits)
    expected[0] = expected[-1] = 1 / np.sqrt(2)
    np.testing.assert_allclose(psi, expected, atol=1e-15)


# ---- test/correctness_tests.jl:53-83 "Random single qubit operators"
@pytest.mark.parametrize("nbits", [5, 8])
def test_random_single_qubit(oracle, nbits):
    rng = np.random.default_rng(53)
    psi = rand_state(rng, nbits)
    gates = [(i, rand_nonunitary(rng, 2)) for i in range(1, nbits + 1)]
    expected = onp.convert_gates_to_matrix(nbits, gates) @ psi
    out, out_np = psi, psi
    for K, u in gates:
        out = oracle.apply_1q(out, u, K)
        out_np = onp.apply_1q(out_np, u, K)
    np.testing.assert_allclose(out, expected, rtol=0, atol=1e-12 * np.linalg.norm(expected))
    np.testing.assert_allclose(out_np, expected, rtol=0, atol=1e-12 * np.linalg.norm(expected))


# ---- test/correctness_tests.jl:86-124 "Random brickwork"
@pytest.mark.parametrize("nbits", [5, 6, 9])
def test_random_brickwork(oracle, nbits):
    rng = np.random.default_rng(86)
    psi = rand_state(rng, nbits)
    for layer in range(1, 4):
        gates = [(i, rand_nonunitary(rng, 4)) for i in onp.circuit_layer_starts(layer, nbits)]
        expected = onp.convert_gates_to_matrix(nbits, gates) @ psi
        out_k, out_s, out_np = psi, psi, psi
        for K, M in gates:
            out_k = oracle.apply_2q(out_k, M, K)
            out_s = oracle.apply_2q(out_s, M, K, scatter=True)
            out_np = onp.apply_2q(out_np, M, K)
        tol = 1e-12 * np.linalg.norm(expected)
        np.testing.assert_allclose(out_k, expected, rtol=0, atol=tol)
        np.testing.assert_allclose(out_s, expected, rtol=0, atol=tol)
        np.testing.assert_allclose(out_np, expected, rtol=0, atol=tol)
        psi = out_k


# ---- examples/test_contractions.jl:11-29 mixed gate list
def test_mixed_gate_list(oracle):
    rng = np.random.default_rng(11)
    nbits = 4
    M = rng.standard_normal((4, 4))
    M /= np.linalg.norm(M)
    psi = rand_state(rng, nbits)
    expected = onp.convert_gates_to_matrix(nbits, [(1, onp.HADAMARD), (2, M), (4, onp.HADAMARD)]) @ psi
    out = oracle.apply_1q(psi, onp.HADAMARD, 1)
    out = oracle.apply_2q(out, M, 2)
    out = oracle.apply_1q(out, onp.HADAMARD, 4)
    np.testing.assert_allclose(out, expected, atol=1e-14)


def test_gate_outside_qubit_space(oracle):
    psi = onp.zero_state(4)
    for K in (0, 4, 5):  # src/apply.jl:112 : 0 < K < nqubits
        with pytest.raises(AssertionError):
            oracle.apply_2q(psi, onp.CNOT4, K)
    for K in (0, 5):  # src/apply.jl:27
        with pytest.raises(AssertionError):
            oracle.apply_1q(psi, onp.HADAMARD, K)


# ---- src/gates.jl:80-109
def test_general_unitary_builder(oracle):
    SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    expect0 = SWAP @ np.diag([1j, 1j, -1j, -1j])
    np.testing.assert_allclose(oracle.build_general_unitary_gate(np.zeros(15)), expect0, atol=1e-15)
    np.testing.assert_allclose(onp.build_general_unitary_gate(np.zeros(15)), expect0, atol=1e-15)
    rng = np.random.default_rng(80)
    for _ in range(20):
        a = rng.uniform(0, 2 * np.pi, 15)
        U = oracle.build_general_unitary_gate(a)
        np.testing.assert_allclose(U @ U.conj().T, np.eye(4), atol=1e-14)
        np.testing.assert_allclose(U, onp.build_general_unitary_gate(a), atol=1e-15)


def test_brickwork_gate_count(oracle):
    # 12q x 20 -> 110, 20q x 4 -> 38, 28q x 4 -> 54, 30q x 100 -> 1450, 34q x 100 -> 1650
    for (n, L, G) in [(12, 20, 110), (20, 4, 38), (28, 4, 54), (30, 100, 1450), (34, 100, 1650), (5, 3, 6)]:
        assert oracle.brickwork_num_gates(n, L) == G == onp.brickwork_num_gates(n, L)


@pytest.mark.parametrize("nbits,nlayers", [(4, 3), (5, 4), (9, 6)])
def test_brickwork_apply_c_vs_numpy(oracle, nbits, nlayers):
    rng = np.random.default_rng(19)
    G = onp.brickwork_num_gates(nbits, nlayers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    psi0 = rand_state(rng, nbits)
    a = oracle.apply_brickwork(psi0, nbits, nlayers, angles)
    b = onp.apply_brickwork(psi0, nbits, nlayers, angles)
    np.testing.assert_allclose(a, b, atol=1e-13)
    assert abs(np.linalg.norm(a) - 1) < 1e-12
    # ComplexF32 state: ComplexF64 gates, one rounding per gate (src/apply.jl:72,100)
    a32 = oracle.apply_brickwork(psi0.astype(np.complex64), nbits, nlayers, angles)
    assert a32.dtype == np.complex64
    assert np.linalg.norm(a32 - b) < 1e-5


# ---- test/correctness_tests.jl:126-150 "East Model Hamiltonian"
@pytest.mark.parametrize("nbits", [5, 6, 8])
def test_east_kernel_vs_dense(oracle, nbits):
    rng = np.random.default_rng(126)
    c, s = 0.1, -0.1
    A, B = onp.east_params(c, s)
    assert np.isclose(A, -np.exp(-s) * np.sqrt(c * (1 - c))) and np.isclose(B, 1 - 2 * c)
    psi = rand_state(rng, nbits)
    E_dense = onp.measure_matrix(onp.dense_east(nbits, c, s), psi)
    E_c = oracle.measure_east(psi, c, A, B)
    E_np = onp.measure_east(psi, c, A, B)
    assert abs(E_c.real - E_dense) < 1e-12 * max(1, abs(E_dense))
    assert abs(E_np.real - E_dense) < 1e-12 * max(1, abs(E_dense))
    assert abs(E_c.imag) < 1e-12
    # closed form on |0...0>
    E0 = oracle.measure_east(onp.zero_state(nbits), c, A, B)
    assert abs(E0.real - (B * (1 + (nbits - 1) * (1 + c)) + c)) < 1e-13
    E32 = oracle.measure_east(psi.astype(np.complex64), c, A, B)
    assert abs(E32.real - E_dense) < 1e-5 * max(1, abs(E_dense))


# ---- examples/test_gradients.jl:10-28, examples/matrix_tfim.jl:7-32
@pytest.mark.parametrize("nbits", [4, 7, 9])
@pytest.mark.parametrize("params", [(1.0, 0.1, 0.5), (1.0, 0.2, 0.5), (1.0, 0.9045, 1.4)])
def test_tfim_kernel_vs_dense(oracle, nbits, params):
    J, h, g = params
    rng = np.random.default_rng(10)
    psi = rand_state(rng, nbits)
    E_dense = onp.measure_matrix(onp.dense_tfim(nbits, J, h, g), psi)
    E_c = oracle.measure_tfim(psi, J, h, g)
    E_np = onp.measure_tfim(psi, J, h, g)
    assert abs(E_c.real - E_dense) < 1e-12 * max(1, abs(E_dense))
    assert abs(E_np.real - E_dense) < 1e-12 * max(1, abs(E_dense))
    assert abs(E_c.imag) < 1e-12
    # closed forms: |0..0> -> -J((n-1) + h n);  |+..+> -> -J g n
    assert abs(oracle.measure_tfim(onp.zero_state(nbits), J, h, g).real + J * ((nbits - 1) + h * nbits)) < 1e-13
    plus = np.full(2 ** nbits, 2 ** (-nbits / 2), dtype=np.complex128)
    assert abs(oracle.measure_tfim(plus, J, h, g).real + J * g * nbits) < 1e-12
    E32 = oracle.measure_tfim(psi.astype(np.complex64), J, h, g)
    assert abs(E32.real - E_dense) < 1e-5 * max(1, abs(E_dense))


# ---- an answer from outside the code base: gates without entangling angles are product gates after a swap
@pytest.mark.parametrize("nbits,nlayers", [(2, 1), (5, 3), (9, 6), (12, 9)])
def test_non_entangling_brickwork_gives_the_analytic_product_state(oracle, nbits, nlayers):
    """src/gates.jl:80-109 with the three entangling angles at zero is (r2 x r1 rz) SWAP (rz r4 x r3); a brickwork circuit
    (src/circuits.jl:9-31) of such gates maps |0..0> to a product state that tests/product_circuit.py tracks as n 2-vectors.
    Pins gate builder, qubit order, layer geometry and the application kernel (src/apply.jl:62-84) at once."""
    from product_circuit import product_amplitudes, product_circuit_vectors

    rng = np.random.default_rng(40 + nbits)
    G = onp.brickwork_num_gates(nbits, nlayers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    angles[12:15] = 0
    want = product_amplitudes(product_circuit_vectors(nbits, nlayers, angles), np.arange(2 ** nbits))
    psi0 = onp.zero_state(nbits).reshape(-1)
    assert np.max(np.abs(oracle.apply_brickwork(psi0, nbits, nlayers, angles).reshape(-1) - want)) < 1e-14
    if nbits <= 9:
        assert np.max(np.abs(np.asarray(onp.apply_brickwork(psi0, nbits, nlayers, angles)).reshape(-1, order="F") - want)) < 1e-14


# ---- an answer from outside the code base: the open transverse-field Ising chain is a free-fermion model
@pytest.mark.parametrize("nbits,J,g", [(6, 1.0, 0.5), (9, 1.0, 1.4), (10, 0.7, 1.0)])
def test_tfim_ground_energy_equals_free_fermion_solution(oracle, nbits, J, g):
    """-J (sum ZZ + g sum X) (h = 0; src/hamiltonians/tfim.jl:17-54, examples/matrix_tfim.jl:7-32) is, after X <-> Z, the chain
    -J sum XX - Jg sum Z, which the Jordan-Wigner map turns into free fermions (Lieb-Schultz-Mattis): with A the symmetric and B
    the antisymmetric hopping / pairing matrices, A + B = 2 (Jg I - J S), S the upper shift of the open chain, and
    E0 = -1/2 sum of the singular values of A + B.  The lowest eigenvalue of the oracle's dense TFIM and the kernel
    expectation value of its ground vector must both equal that number."""
    sv = np.linalg.svd(J * g * np.eye(nbits) - J * np.eye(nbits, k=1), compute_uv=False)
    E0 = -float(np.sum(sv))
    w, v = np.linalg.eigh(onp.dense_tfim(nbits, J, 0.0, g))
    assert abs(w[0] - E0) < 1e-11 * nbits
    gs = v[:, 0].astype(np.complex128)
    assert abs(oracle.measure_tfim(gs, J, 0.0, g).real - E0) < 1e-11 * nbits
    assert abs(onp.measure_tfim(gs, J, 0.0, g).real - E0) < 1e-11 * nbits


# ---- examples/test_gradient_accuracy.jl:25-57 : gradients! == naive per-angle shift
@pytest.mark.parametrize("ham", ["tfim", "east"])
def test_gradients_vs_naive_shift(oracle, ham):
    rng = np.random.default_rng(25)
    nbits, nlayers = 4, 3
    G = onp.brickwork_num_gates(nbits, nlayers)
    angles = rng.standard_normal((15, G)) * 0.1
    psi0 = onp.zero_state(nbits)
    if ham == "tfim":
        kind, hp = 0, (1.0, 0.1, 0.5)
        efn = lambda v: onp.measure_tfim(v, *hp).real
    else:
        A, B = onp.east_params(0.1, -0.1)
        kind, hp = 1, (0.1, A, B)
        efn = lambda v: onp.measure_east(v, *hp).real
    E, grads = oracle.gradients_brickwork(psi0, nbits, nlayers, angles, kind, hp, want_energy=True)
    assert abs(E - efn(onp.apply_brickwork(psi0, nbits, nlayers, angles))) < 1e-13
    naive = onp.naive_shift_gradients(nbits, nlayers, angles, psi0, efn)
    assert np.max(np.abs(grads - naive)) < 200 * np.finfo(float).eps  # reference's own threshold
    # the shift rule is exact: compare with a central finite difference
    eps = 1e-6
    for (k, g) in [(0, 0), (7, 2), (14, G - 1), (12, 1)]:
        a = angles.copy()
        a[k, g] += eps
        ep = efn(onp.apply_brickwork(psi0, nbits, nlayers, a))
        a[k, g] -= 2 * eps
        em = efn(onp.apply_brickwork(psi0, nbits, nlayers, a))
        assert abs((ep - em) / (2 * eps) - grads[k, g]) < 1e-8
    only = oracle.gradients_brickwork(psi0, nbits, nlayers, angles, kind, hp, want_energy=False)
    np.testing.assert_array_equal(only, grads)


def test_kernel_tfim_gradients_equal_matrix_tfim_gradients(oracle):
    # examples/test_gradients.jl:10-28 (n=8 there; n=5 keeps the dense shift cheap)
    rng = np.random.default_rng(28)
    nbits, nlayers = 5, 4
    G = onp.brickwork_num_gates(nbits, nlayers)
    angles = rng.standard_normal((15, G)) * 0.01
    psi0 = onp.zero_state(nbits)
    H = onp.dense_tfim(nbits, 1.0, 0.1, 0.5)
    E, grads = oracle.gradients_brickwork(psi0, nbits, nlayers, angles, 0, (1.0, 0.1, 0.5))
    naive = onp.naive_shift_gradients(nbits, nlayers, angles, psi0, lambda v: onp.measure_matrix(H, v))
    assert abs(E - onp.measure_matrix(H, onp.apply_brickwork(psi0, nbits, nlayers, angles))) < 1e-12
    assert np.max(np.abs(grads - naive)) < 100 * np.finfo(float).eps * max(1.0, np.max(np.abs(naive)) * 10)


def test_optimise_descends(oracle):
    # src/gradients.jl:1-17 with quirk Q1 resolved (energy + full gradient step)
    rng = np.random.default_rng(1)
    nbits, nlayers, epochs, lr = 4, 2, 5, 0.01
    G = onp.brickwork_num_gates(nbits, nlayers)
    angles = rng.standard_normal((15, G)) * 0.01
    psi0 = onp.zero_state(nbits)
    hp = (1.0, 0.2, 0.5)
    energies, final = oracle.optimise_brickwork(psi0, nbits, nlayers, angles, 0, hp, epochs, lr)
    assert energies.shape == (epochs + 1,) and energies[0] == 0.0
    # energies[1] is the energy at the initial angles; the last entry is the final measure
    assert abs(energies[1] - onp.measure_tfim(onp.apply_brickwork(psi0, nbits, nlayers, angles), *hp).real) < 1e-13
    assert abs(energies[-1] - onp.measure_tfim(onp.apply_brickwork(psi0, nbits, nlayers, final), *hp).real) < 1e-13
    assert np.all(np.diff(energies[1:]) < 0)
    # one manual step reproduces the update rule theta <- theta - lr * grad
    _, g0 = oracle.gradients_brickwork(psi0, nbits, nlayers, angles, 0, hp)
    e1, a1 = oracle.optimise_brickwork(psi0, nbits, nlayers, angles, 0, hp, 1, lr)
    np.testing.assert_allclose(a1, angles - lr * g0, atol=1e-16)


def test_threads_do_not_change_results(oracle):
    rng = np.random.default_rng(3)
    nbits, nlayers = 10, 4
    G = onp.brickwork_num_gates(nbits, nlayers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    psi0 = rand_state(rng, nbits)
    oracle.set_threads(1)
    a = oracle.apply_brickwork(psi0, nbits, nlayers, angles)
    oracle.set_threads(0)
    b = oracle.apply_brickwork(psi0, nbits, nlayers, angles)
    np.testing.assert_array_equal(a, b)


# ---- src/apply.jl:144-182 : right_apply_gate!(M', M, gate) is M * U (operator evolution), restated literally -----------------
@pytest.mark.parametrize("nbits", [2, 3, 5])
def test_right_apply_gate_is_right_multiplication(nbits):
    rng = np.random.default_rng(144 + nbits)
    M = rng.standard_normal((2 ** nbits, 2 ** nbits)) + 1j * rng.standard_normal((2 ** nbits, 2 ** nbits))
    for K in range(1, nbits):
        u4 = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        got = onp.right_apply_gate(M, u4, K)
        U = onp.convert_gates_to_matrix(nbits, [(K, u4)])
        np.testing.assert_allclose(got, M @ U, atol=1e-12)
    # the identity evolved through a brick layer reproduces the layer's operator (the use the reference has for it:
    # combine_gates / build_identity, src/utils.jl)
    I = np.eye(2 ** nbits, dtype=np.complex128)
    u4 = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    np.testing.assert_allclose(onp.right_apply_gate(I, u4, 1), onp.convert_gates_to_matrix(nbits, [(1, u4)]), atol=1e-13)


# ---- ext/MPSExt/polar_optimise.jl: the convention of the un-vendored MatrixProductStates.contract, pinned by its call sites ----
def test_polar_contract_convention_is_pinned_by_the_call_sites():
    rng = np.random.default_rng(88)
    N = 6
    phi = rng.standard_normal(2 ** N) + 1j * rng.standard_normal(2 ** N)
    ident = [[np.eye(4, dtype=np.complex128) for _ in range(N - 1)] for _ in range(2)]
    # an all-identity circuit must hand back conj(phi) (polar_optimise.jl:52-78) -- true for "free dims of A first" only
    np.testing.assert_allclose(onp.create_upper_state_literal(ident, phi, N), np.conj(phi), atol=0)
    # the other conceivable reading (free dims of the state first) scrambles the qubit order already at a single call site:
    # contract with the identity gate + permutedims(perm_idxs) must be the identity map on the state
    T = phi.reshape((2,) * N, order="F")
    I4 = np.eye(4, dtype=np.complex128).reshape(2, 2, 2, 2, order="F")
    for site in range(1, N):
        ok = onp._permutedims(onp.mps_contract(I4, T, [3, 4], [site, site + 1], "A"), onp.perm_idxs(site, N))
        bad = onp._permutedims(onp.mps_contract(I4, T, [3, 4], [site, site + 1], "B"), onp.perm_idxs(site, N))
        assert np.array_equal(ok, T)
        assert np.linalg.norm(bad - T) > 0.5
    # with that reading, contract(gate, psi, [3,4], [site, site+1]) + permutedims(perm_idxs) IS apply!(psi', psi, gate) of
    # src/apply.jl on (site, site+1), and [1,2] applies the transposed gate
    gates = [[rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)) for _ in range(N - 1)] for _ in range(2)]
    sites = [2 * idx - 1 for idx in range(1, N // 2 + 1)] + [2 * idx - N for idx in range(N // 2 + 1, N)]
    ref = onp.zero_state(N)
    for layer in gates:
        for g, site in zip(layer, sites):
            ref = onp.apply_2q(ref, g, site)
    np.testing.assert_allclose(onp.circuit_to_state_literal(gates, N), ref, atol=1e-9 * np.linalg.norm(ref))
    up = np.conj(phi)
    for layer in reversed(gates):
        for g, site in reversed(list(zip(layer, sites))):
            up = onp.apply_2q(up, g.T, site)
    np.testing.assert_allclose(onp.create_upper_state_literal(gates, phi, N), up, atol=1e-9 * np.linalg.norm(up))


@pytest.mark.parametrize("N,nl", [(4, 1), (6, 2)])
def test_polar_optimise_restatement_equals_literal_tensor_form(N, nl):
    """the flat-vector restatement the GPU path is checked against == the statement-by-statement tensor form"""
    hp = (1.0, 0.9045, 1.4)
    w, v = np.linalg.eigh(onp.dense_tfim(N, *hp))
    gs = v[:, 0].astype(np.complex128)
    ident = [[np.eye(4, dtype=np.complex128) for _ in range(N - 1)] for _ in range(nl)]
    en = lambda s: onp.measure_matrix(onp.dense_tfim(N, *hp), s)  # noqa: E731
    c1, ov1, e1 = onp.polar_optimise(ident, gs, en, N, 4)
    c2, ov2, e2 = onp.polar_optimise_literal(ident, gs, en, N, 4)
    np.testing.assert_allclose(ov1, ov2, atol=1e-12)
    np.testing.assert_allclose(e1, e2, atol=1e-11)
    for la, lb in zip(c1, c2):
        for a, b in zip(la, lb):
            np.testing.assert_allclose(a, b, atol=1e-10)


def test_polar_optimise_invariants():
    """What the algorithm guarantees whatever the data: every polar update maximises Re tr(E U) over unitaries, so the
    overlap with the target never decreases along a sweep sequence; a target that the circuit represents exactly is a fixed
    point (overlap 1 from the first sweep on, gates unchanged up to the sweep's gauge)."""
    rng = np.random.default_rng(91)
    N, nl = 6, 2
    sites = [2 * idx - 1 for idx in range(1, N // 2 + 1)] + [2 * idx - N for idx in range(N // 2 + 1, N)]

    def rand_u():
        q, r = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))
        return q * (np.diag(r) / np.abs(np.diag(r)))
    exact = [[rand_u() for _ in range(N - 1)] for _ in range(nl)]
    target = onp.zero_state(N)
    for layer in exact:
        for g, site in zip(layer, sites):
            target = onp.apply_2q(target, g, site)
    en = lambda s: 0.0  # noqa: E731
    # fixed point
    c, ov, _ = onp.polar_optimise(exact, target, en, N, 3)
    np.testing.assert_allclose(ov, 1.0, atol=1e-12)
    out = onp.zero_state(N)
    for layer in c:
        for g, site in zip(layer, sites):
            out = onp.apply_2q(out, g, site)
    assert abs(abs(np.vdot(target, out)) - 1.0) < 1e-12
    # monotone from the identity start, and from a random start
    for start in ([[np.eye(4, dtype=np.complex128) for _ in range(N - 1)] for _ in range(nl)], [[rand_u() for _ in range(N - 1)] for _ in range(nl)]):
        _, ov, _ = onp.polar_optimise(start, target, en, N, 12)
        assert np.all(np.diff(ov) > -1e-12), ov
        assert ov[-1] > ov[0]
