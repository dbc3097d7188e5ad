"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, fails loudly without a GPU, and its host-side gate algebra matches the oracle."""
import ctypes as C

import numpy as np
import pytest

import qcb200
from qcb200 import _lib
from oracle import oracle_np as onp


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.QCBError) as ei:
        qcb200.B200State(4)
    assert ei.value.code == _lib.QCB_ECUDA
    with pytest.raises(_lib.QCBError):
        qcb200.apply(onp.zero_state(4), qcb200.GenericBrickworkCircuit(4, 2))
    with pytest.raises(_lib.QCBError):
        qcb200.measure(qcb200.TFIMHamiltonian(1.0, 0.1, 0.5), onp.zero_state(4))


def test_general_unitary_gate_matches_oracle(oracle):
    rng = np.random.default_rng(0)
    for _ in range(25):
        a = rng.uniform(-np.pi, 2 * np.pi, 15)
        u = qcb200.build_general_unitary_gate(a).mat().reshape(4, 4, order="F")
        np.testing.assert_allclose(u, oracle.build_general_unitary_gate(a), atol=2e-16 * 8)
        np.testing.assert_allclose(u, onp.build_general_unitary_gate(a), atol=1e-15)


def test_general_unitary_gate_is_the_canonical_two_qubit_gate(oracle):
    """A published answer from outside the code base: src/gates.jl:80-109 is the three-CNOT circuit of Vatan & Williams for a
    general two-qubit gate, so its local invariants (Makhlin: G1 = tr^2 m / 16 det U, G2 = (tr^2 m - tr m^2) / 4 det U with
    m = (Q'UQ)^T (Q'UQ), Q the magic basis) depend on the three entangling angles alone and equal those of
    exp(i (a XX + b YY + c ZZ)) with (a, b, c) = (theta, phi, lambda) / 2 + pi / 4, whatever the twelve local angles are:
    G1 = cos^2 2a cos^2 2b cos^2 2c - sin^2 2a sin^2 2b sin^2 2c + (i / 4) sin 4a sin 4b sin 4c, G2 = cos 4a + cos 4b + cos 4c.
    Checked for the library's host gate builder (qcb_general_unitary_gate) and both oracles; the matrix is unitary."""
    Q = np.array([[1, 0, 0, 1j], [0, 1j, 1, 0], [0, 1j, -1, 0], [1, 0, 0, -1j]]) / np.sqrt(2)
    rng = np.random.default_rng(11)
    for _ in range(20):
        ang = rng.uniform(-np.pi, 2 * np.pi, 15)
        a, b, c = ang[12:15] / 2 + np.pi / 4
        G1 = (np.cos(2 * a) * np.cos(2 * b) * np.cos(2 * c)) ** 2 - (np.sin(2 * a) * np.sin(2 * b) * np.sin(2 * c)) ** 2 \
            + 0.25j * np.sin(4 * a) * np.sin(4 * b) * np.sin(4 * c)
        G2 = np.cos(4 * a) + np.cos(4 * b) + np.cos(4 * c)
        for U in (qcb200.build_general_unitary_gate(ang).mat().reshape(4, 4, order="F"), oracle.build_general_unitary_gate(ang),
                  onp.build_general_unitary_gate(ang)):
            np.testing.assert_allclose(U.conj().T @ U, np.eye(4), atol=1e-14)
            Ub = Q.conj().T @ U @ Q
            m = Ub.T @ Ub
            d = np.linalg.det(U)
            t = np.trace(m)
            assert abs(t * t / (16 * d) - G1) < 1e-12
            assert abs((t * t - np.trace(m @ m)) / (4 * d) - G2) < 1e-12


def test_general_unitary_gate_derivative():
    L = _lib.lib()
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 2 * np.pi, 15)
    eps = 1e-6
    for k in range(15):
        d = np.empty(16, dtype=np.complex128)
        _lib.check(L.qcb_general_unitary_gate_derivative(a.ctypes.data_as(C.c_void_p), C.c_int(k), d.ctypes.data_as(C.c_void_p)))
        ap, am = a.copy(), a.copy()
        ap[k] += eps
        am[k] -= eps
        fd = (onp.build_general_unitary_gate(ap) - onp.build_general_unitary_gate(am)) / (2 * eps)
        np.testing.assert_allclose(d.reshape(4, 4, order="F"), fd, atol=1e-9)


def test_gate_types_match_reference_tables():
    # src/gates.jl:10-16 and the 4x4 forms verified in SURVEY.md 8a a3
    assert np.array_equal(qcb200.CNOTGate().mat().reshape(4, 4, order="F"), onp.CNOT4)
    assert np.array_equal(qcb200.ReverseCNOTGate().mat().reshape(4, 4, order="F"), onp.RCNOT4)
    np.testing.assert_allclose(qcb200.HadamardGate().mat(), onp.HADAMARD)
    np.testing.assert_allclose(qcb200.RZGate(0.3).mat(), onp.rz(0.3))
    np.testing.assert_allclose(qcb200.RYGate(0.3).mat(), onp.ry(0.3))
    np.testing.assert_allclose(qcb200.RXGate(0.3).mat(), onp.rx(0.3))
    np.testing.assert_allclose(qcb200.RotationGate(0.3, 0.4, 0.5).mat(), onp.rotation(0.3, 0.4, 0.5))
    g = qcb200.Generic2SpinGate(np.arange(16).reshape(2, 2, 2, 2, order="F") * (1 + 1j))
    np.testing.assert_allclose(qcb200.adjoint(g).mat().reshape(4, 4, order="F"), g.mat().reshape(4, 4, order="F").conj().T)


def test_circuit_geometry():
    assert list(qcb200.circuit_layer_starts(1, 5)) == [1, 3] and list(qcb200.circuit_layer_starts(2, 5)) == [2, 4]
    for n, L, G in [(12, 20, 110), (20, 4, 38), (28, 4, 54), (30, 100, 1450), (34, 100, 1650)]:
        assert qcb200.brickwork_num_gates(n, L) == G
        assert qcb200.GenericBrickworkCircuit(n, L).gate_angles.shape == (15, G)
    c = qcb200.GenericBrickworkCircuit(4, 2)
    c2 = qcb200.reconstruct(c, 16, 0.5)  # 0-based linear index -> angle [1, 1]
    assert c2.gate_angles[1, 1] == 0.5 and c.gate_angles[1, 1] == 0.0


def test_zero_state_constructors():
    t = qcb200.zero_state_tensor(3)
    assert t.shape == (2, 2, 2) and t[0, 0, 0] == 1 and abs(t).sum() == 1
    v = qcb200.zero_state_vec(np.complex64, 3)
    assert v.dtype == np.complex64 and v[0] == 1


def test_convert_gates_to_matrix_matches_checker():
    # src/utils.jl:10-40 ; test/correctness_tests.jl:30-51 uses it to build the GHZ state with matrices
    rng = np.random.default_rng(2)
    nbits = 5
    u1 = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
    u2 = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    gates = [qcb200.Localised1SpinGate(qcb200.Generic1SpinGate(u1), 1),
             qcb200.Localised2SpinAdjGate(qcb200.Generic2SpinGate(u2.reshape(2, 2, 2, 2, order="F")), 3)]
    M = qcb200.convert_gates_to_matrix(nbits, gates)
    np.testing.assert_allclose(M, onp.convert_gates_to_matrix(nbits, [(1, u1), (3, u2)]), atol=1e-15)
    for n in (4, 5, 7):
        psi = qcb200.zero_state_vec(n)
        psi = qcb200.convert_gates_to_matrix(n, [qcb200.Localised1SpinGate(qcb200.HadamardGate(), 1)]) @ psi
        for i in range(1, n):
            psi = qcb200.convert_gates_to_matrix(n, [qcb200.Localised2SpinAdjGate(qcb200.CNOTGate(), i)]) @ psi
        expected = np.zeros(2 ** n)
        expected[0] = expected[-1] = 1 / np.sqrt(2)
        np.testing.assert_allclose(psi, expected, atol=1e-15)


# ---- matrix Hamiltonians: host builders (src/hamiltonians/tfim.jl:58-92, east_model.jl:57-81) ----------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8])
def test_sparse_tfim_matrix_matches_dense_kronecker_definition(n):
    J, h, g = 1.0, 0.1, 0.5  # examples/test_gradients.jl:12-16
    H = qcb200.build_sparse_tfim_hamiltonian(n, J, h, g)
    assert H.shape == (2 ** n, 2 ** n)
    np.testing.assert_allclose(H.toarray(), onp.dense_tfim(n, J, h, g), atol=1e-13)
    assert abs(H - H.T).max() == 0  # LinearAlgebra.Symmetric(H)   tfim.jl:84


@pytest.mark.parametrize("n", [2, 3, 5])
def test_east_model_matrix_matches_oracle(n):
    H = qcb200.EastModelHamiltonian(0.1, -0.1)  # test/correctness_tests.jl:126-150
    M = qcb200.mat(H, n)
    np.testing.assert_allclose(M, onp.dense_east(n, 0.1, -0.1), atol=1e-13)
    np.testing.assert_allclose(qcb200.east_model_matrix(H, n), M)
    with pytest.raises(TypeError):
        qcb200.mat(qcb200.TFIMHamiltonian(1.0, 0.1, 0.5), n)


def test_find_tfim_ground_state():
    n, J, h, g = 6, 1.0, 0.9045, 1.4  # examples/test_optimise.jl:9-16
    E, v = qcb200.find_tfim_ground_state(n, J, h, g)
    w, V = np.linalg.eigh(onp.dense_tfim(n, J, h, g))
    assert abs(E - w[0]) < 1e-10
    assert abs(abs(np.vdot(V[:, 0], v)) - 1) < 1e-8


def test_general_unitary_gradient_matches_per_angle_derivatives():
    """qcb_general_unitary_gate_gradient (15 angles from shared partial products) == contraction with each
    qcb_general_unitary_gate_derivative, which the oracle pins by finite differences elsewhere in this file."""
    rng = np.random.default_rng(15)
    L = _lib.lib()
    for _ in range(20):
        a = np.ascontiguousarray(rng.uniform(-2 * np.pi, 2 * np.pi, 15))
        env = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        env_abi = np.ascontiguousarray(env.reshape(-1, order="F"))
        out = np.zeros(15)
        _lib.check(L.qcb_general_unitary_gate_gradient(a.ctypes.data_as(C.c_void_p), env_abi.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        for k in range(15):
            d = np.empty(16, dtype=np.complex128)
            _lib.check(L.qcb_general_unitary_gate_derivative(a.ctypes.data_as(C.c_void_p), C.c_int(k), d.ctypes.data_as(C.c_void_p)))
            ref = 2 * np.real(np.sum(env * d.reshape(4, 4, order="F")))
            assert abs(out[k] - ref) < 1e-13 * max(1.0, abs(ref))
