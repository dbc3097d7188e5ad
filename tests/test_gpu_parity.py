"""GPU parity tests: the CUDA path, called through the C ABI (Python mirror of the reference API), against
the CPU oracle and the committed golden fixtures.  Tolerances are BASELINE.json's: 1e-12 relative for
ComplexF64 states, 1e-5 for ComplexF32 (norm-wise; gradients relative to the gradient's max-norm)."""
import os

import numpy as np
import pytest

import qcb200 as qc
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {np.complex128: 1e-12, np.complex64: 1e-5}


HAM_TILE_DEFAULT = 1


def rand_state(rng, n, dtype=np.complex128):
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    return (v / np.linalg.norm(v)).astype(dtype)


def rel(a, b):
    # tensors are column-major 2x2x...x2 (Julia layout): flatten in Fortran order
    a = np.asarray(a, dtype=np.complex128).reshape(-1, order="F")
    b = np.asarray(b, dtype=np.complex128).reshape(-1, order="F")
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def rand_nonunitary(rng, d):
    while True:
        u = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
        det = np.linalg.det(u)
        if det != 0:
            return u / det


def circuit(n, layers, rng, scale=None):
    c = qc.GenericBrickworkCircuit(n, layers)
    c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape) if scale is None else rng.standard_normal(c.gate_angles.shape) * scale
    return c


PASS_ASYNC_DEFAULT = 0


@pytest.fixture(autouse=True)
def _defaults():
    qc.init()
    qc.set_option("force_simple", 0)
    qc.set_option("tile_bits", 11)
    qc.set_option("low_bits", 3)
    qc.set_option("max_gates_per_pass", 1 << 30)
    qc.set_option("adjoint_mma", 1)
    qc.set_option("adjoint_tile_bits", 0)
    qc.set_option("adjoint_prog_host", 0)
    qc.set_option("contract_coop", 1)
    qc.set_option("measure_tma", 1)
    qc.set_option("pass_tma", 0)
    qc.set_option("f64_3m_min_gates", 4)
    qc.set_option("pass_async", PASS_ASYNC_DEFAULT)
    qc.set_option("small_cluster", 0)  # the tests below exercise the tile passes also on small states; the cluster kernel has its own tests
    qc.set_option("adjoint_f32", 1)
    qc.set_option("adjoint_ctas", 5)
    qc.set_option("adjoint_defer", 1)
    qc.set_option("tma_l2_promotion", 1)
    yield


# ---- test/correctness_tests.jl:3-27
@pytest.mark.parametrize("nbits", [1, 2, 4, 5, 7, 12, 20])
def test_cat_state(nbits):
    psi = qc.B200State(qc.zero_state_tensor(nbits))
    psi2 = psi.similar()
    out = qc.apply_(psi2, psi, qc.Localised1SpinGate(qc.HadamardGate(), 1))
    psi, psi2 = out, psi
    for i in range(1, nbits):
        out = qc.apply_(psi2, psi, qc.Localised2SpinAdjGate(qc.CNOTGate(), i))
        psi, psi2 = out, psi
    v = psi.to_numpy()
    expected = np.zeros(2 ** nbits, dtype=np.complex128)
    expected[0] = expected[-1] = 1 / np.sqrt(2)
    if nbits == 1:
        expected[:] = 1 / np.sqrt(2)
    np.testing.assert_allclose(v, expected, atol=1e-15)


# ---- test/correctness_tests.jl:53-83
@pytest.mark.parametrize("nbits", [5, 8, 10, 13])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_random_single_qubit_operators(nbits, dtype):
    rng = np.random.default_rng(53 + nbits)
    psi = rand_state(rng, nbits, dtype)
    mats = [rand_nonunitary(rng, 2) for _ in range(nbits)]
    gates = [qc.Localised1SpinGate(qc.Generic1SpinGate(m), i + 1) for i, m in enumerate(mats)]
    expected = psi.astype(np.complex128)
    for i, m in enumerate(mats):
        expected = onp.apply_1q(expected, m, i + 1)
    if nbits <= 10:
        dense = onp.convert_gates_to_matrix(nbits, [(i + 1, m) for i, m in enumerate(mats)]) @ psi.astype(np.complex128)
        assert rel(expected, dense) < 1e-12
    out = qc.apply(psi.reshape((2,) * nbits, order="F"), gates)
    assert out.shape == (2,) * nbits and out.dtype == dtype
    assert rel(out.reshape(-1, order="F"), expected) < TOL[dtype]


# ---- test/correctness_tests.jl:86-124
@pytest.mark.parametrize("nbits", [5, 6, 9, 10, 14])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_random_brickwork_layers(oracle, nbits, dtype):
    rng = np.random.default_rng(86 + nbits)
    psi = rand_state(rng, nbits, dtype)
    for layer in range(1, 4):
        mats = {i: rand_nonunitary(rng, 4) for i in qc.circuit_layer_starts(layer, nbits)}
        gates = [qc.Localised2SpinAdjGate(qc.Generic2SpinGate(m.reshape(2, 2, 2, 2, order="F")), i) for i, m in mats.items()]
        expected = psi.astype(np.complex128)
        for i, m in mats.items():
            expected = oracle.apply_2q(expected, m, i)
        if nbits <= 10:
            dense = onp.convert_gates_to_matrix(nbits, list(mats.items())) @ psi.astype(np.complex128)
            assert rel(expected, dense) < 1e-12
        out = qc.apply(psi, gates)
        assert rel(out, expected) < TOL[dtype]
        psi = out


# ---- examples/test_contractions.jl:11-29
def test_mixed_gate_list():
    rng = np.random.default_rng(11)
    M = rng.standard_normal((4, 4))
    M /= np.linalg.norm(M)
    psi = rand_state(rng, 4)
    gates = [qc.Localised1SpinGate(qc.HadamardGate(), 1), qc.Localised2SpinAdjGate(qc.Generic2SpinGate(M.reshape(2, 2, 2, 2, order="F")), 2),
             qc.Localised1SpinGate(qc.HadamardGate(), 4)]
    expected = onp.convert_gates_to_matrix(4, [(1, onp.HADAMARD), (2, M), (4, onp.HADAMARD)]) @ psi
    assert rel(qc.apply(psi, gates), expected) < 1e-12
    # a 13-qubit mix goes through the fused tile path
    n = 13
    psi = rand_state(rng, n)
    gl, expected = [], psi
    for _ in range(50):
        if rng.random() < 0.4:
            k, m = int(rng.integers(1, n + 1)), rand_nonunitary(rng, 2)
            gl.append(qc.Localised1SpinGate(qc.Generic1SpinGate(m), k))
            expected = onp.apply_1q(expected, m, k)
        else:
            k, m = int(rng.integers(1, n)), rand_nonunitary(rng, 4) / 3
            gl.append(qc.Localised2SpinAdjGate(qc.Generic2SpinGate(m.reshape(2, 2, 2, 2, order="F")), k))
            expected = onp.apply_2q(expected, m, k)
    assert rel(qc.apply(psi, gl), expected) < 1e-12


# ---- apply!(psi', psi, gate) with distinct states (src/apply.jl:20,39,117): psi' = gate(psi), psi untouched; no copy in front
@pytest.mark.parametrize("n", [1, 3, 8, 12, 15])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_out_of_place_gate_apply(n, dtype):
    rng = np.random.default_rng(170 + n)
    psi0 = rand_state(rng, n, dtype)
    src = qc.B200State(psi0)
    dst = src.similar()
    m1 = rand_nonunitary(rng, 2)
    res = qc.apply_(dst, src, qc.Localised1SpinGate(qc.Generic1SpinGate(m1), n))
    assert res is dst and np.array_equal(src.to_numpy(), psi0)
    assert rel(dst.to_numpy(), onp.apply_1q(psi0.astype(np.complex128), m1, n)) < TOL[dtype]
    if n >= 2:
        for K in sorted({1, n - 1}):
            m2 = rand_nonunitary(rng, 4) / 3
            g = qc.Localised2SpinAdjGate(qc.Generic2SpinGate(m2.reshape(2, 2, 2, 2, order="F")), K)
            qc.apply_(qc.construct_apply_cache(src), dst, src, g)
            assert np.array_equal(src.to_numpy(), psi0)
            assert rel(dst.to_numpy(), onp.apply_2q(psi0.astype(np.complex128), m2, K)) < TOL[dtype]
            out = qc.apply(src, [g, g])  # device state in -> new device state out, input untouched
            assert np.array_equal(src.to_numpy(), psi0)
            assert rel(out.to_numpy(), onp.apply_2q(onp.apply_2q(psi0.astype(np.complex128), m2, K), m2, K)) < TOL[dtype]


# ---- src/apply.jl:106-115 error behaviour
def test_error_paths():
    psi = qc.B200State(qc.zero_state_tensor(4))
    out = psi.similar()
    for K in (0, 4, 7):
        with pytest.raises(AssertionError):
            qc.apply_(out, psi, qc.Localised2SpinAdjGate(qc.CNOTGate(), K))
    for K in (0, 5):
        with pytest.raises(AssertionError):
            qc.apply_(out, psi, qc.Localised1SpinGate(qc.HadamardGate(), K))
    with pytest.raises(ValueError, match="same size"):
        qc.apply_(qc.B200State(qc.zero_state_tensor(5)), psi, qc.Localised2SpinAdjGate(qc.CNOTGate(), 1))
    with pytest.raises(ValueError, match="same size"):
        qc.apply_(qc.B200State(qc.zero_state_tensor(np.complex64, 4)), psi, qc.Localised2SpinAdjGate(qc.CNOTGate(), 1))
    with pytest.raises(AssertionError, match="backend"):
        qc.apply_(np.zeros(16, dtype=np.complex128), psi, qc.Localised2SpinAdjGate(qc.CNOTGate(), 1))
    with pytest.raises(qc.QCBError):
        qc.apply(psi, qc.GenericBrickworkCircuit(5, 2))


# ---- src/circuits.jl:19-42 : brickwork circuits, fused tile path vs oracle
@pytest.mark.parametrize("n,layers,scale", [(4, 3, None), (5, 8, None), (8, 40, 0.1), (9, 7, None), (11, 12, None), (12, 20, None),
                                            (13, 9, 0.01), (14, 16, None), (16, 11, None), (17, 6, None), (20, 8, None), (22, 5, None)])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_brickwork_vs_oracle(oracle, n, layers, scale, dtype):
    rng = np.random.default_rng(1000 * n + layers)
    c = circuit(n, layers, rng, scale)
    psi0 = rand_state(rng, n, dtype) if n % 2 else qc.zero_state_vec(dtype, n)
    ref = oracle.apply_brickwork(np.asarray(psi0, dtype=np.complex128), n, layers, c.gate_angles)
    out = qc.apply(psi0, c)
    assert out.dtype == dtype
    assert rel(out, ref) < TOL[dtype]
    if dtype == np.complex64:  # and against the reference's own ComplexF32 semantics
        ref32 = oracle.apply_brickwork(np.asarray(psi0), n, layers, c.gate_angles)
        assert rel(out, ref32) < TOL[dtype]


@pytest.mark.parametrize("tb,lowb", [(9, 3), (10, 4), (12, 3), (12, 5), (13, 3), (13, 4)])
def test_brickwork_tile_shapes(oracle, tb, lowb):
    rng = np.random.default_rng(tb * 10 + lowb)
    n, layers = 16, 9
    c = circuit(n, layers, rng)
    psi0 = rand_state(rng, n)
    ref = oracle.apply_brickwork(psi0, n, layers, c.gate_angles)
    qc.set_option("tile_bits", tb)
    qc.set_option("low_bits", lowb)
    for dtype in (np.complex128, np.complex64):
        assert rel(qc.apply(psi0.astype(dtype), c), ref) < TOL[dtype]
    qc.set_option("max_gates_per_pass", 3)
    assert rel(qc.apply(psi0, c), ref) < 1e-12


def test_golden_fixtures():
    d = np.load(os.path.join(GOLD, "brickwork_12q_20l.npz"))
    for tag in ("uniform", "small"):
        c = qc.GenericBrickworkCircuit(12, 20, gate_angles=d[f"angles_{tag}"])
        assert rel(qc.apply(qc.zero_state_tensor(12), c), d[f"psi_{tag}"]) < 1e-12
        assert rel(qc.apply(qc.zero_state_tensor(np.complex64, 12), c), d[f"psi32_{tag}"]) < 1e-5
    g = np.load(os.path.join(GOLD, "generic_gates_9q.npz"))
    gates = [qc.Localised2SpinAdjGate(qc.Generic2SpinGate(m.reshape(2, 2, 2, 2, order="F")), int(k)) for k, m in zip(g["K"], g["mats"])]
    assert rel(qc.apply(g["psi0"], gates), g["psi"]) < 1e-12
    m = np.load(os.path.join(GOLD, "measure_10q.npz"))
    for p, e in zip([(1.0, 0.1, 0.5), (1.0, 0.9045, 1.4)], m["tfim"]):
        assert abs(qc.measure(qc.TFIMHamiltonian(*p), m["psi"]) - e) < 1e-12 * abs(e)
    assert abs(qc.measure(qc.EastModelHamiltonian(0.1, -0.1), m["psi"]) - m["east"][0]) < 1e-12
    gr = np.load(os.path.join(GOLD, "gradients_8q_4l.npz"))
    for tag in ("small", "uniform"):
        c = qc.GenericBrickworkCircuit(8, 4, gate_angles=gr[f"angles_{tag}"])
        for name, H in (("tfim", qc.TFIMHamiltonian(1.0, 0.1, 0.5)), ("east", qc.EastModelHamiltonian(0.1, -0.1))):
            E, grads = qc.gradients(H, qc.zero_state_tensor(8), c, calculate_energy=True)
            ref = gr[f"{name}_grads_{tag}"]
            assert abs(E - float(gr[f"{name}_E_{tag}"])) < 1e-12 * max(1, abs(E))
            assert np.max(np.abs(grads - ref)) < 1e-12 * max(np.max(np.abs(ref)), 1e-3)
    o = np.load(os.path.join(GOLD, "optimise_6q.npz"))
    c = qc.GenericBrickworkCircuit(6, 2, gate_angles=o["angles"])
    energies = qc.optimise_(c, qc.TFIMHamiltonian(1.0, 0.2, 0.5), qc.zero_state_tensor(6), 5, 0.01)
    np.testing.assert_allclose(energies, o["energies"], atol=1e-10)
    np.testing.assert_allclose(c.gate_angles, o["final"], atol=1e-10)


# ---- test/correctness_tests.jl:126-150, examples/test_gradients.jl : kernel Hamiltonians
@pytest.mark.parametrize("nbits", [1, 2, 5, 8, 10, 14, 20])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_measure_vs_oracle(oracle, nbits, dtype):
    rng = np.random.default_rng(126 + nbits)
    psi = rand_state(rng, nbits, dtype)
    tol = TOL[dtype]
    for p in [(1.0, 0.1, 0.5), (1.0, 0.9045, 1.4), (0.7, -0.3, 0.0)]:
        H = qc.TFIMHamiltonian(*p)
        ref = oracle.measure_tfim(psi.astype(np.complex128), *p).real
        assert abs(qc.measure(H, psi) - ref) < tol * max(1.0, abs(ref))
        if nbits <= 10:
            assert abs(ref - onp.measure_matrix(onp.dense_tfim(nbits, *p), psi.astype(np.complex128))) < 1e-11
    H = qc.EastModelHamiltonian(0.1, -0.1)
    ref = oracle.measure_east(psi.astype(np.complex128), H.c, H.A, H.B).real
    assert abs(qc.measure(H, psi) - ref) < tol * max(1.0, abs(ref))
    # closed forms
    # (ComplexF32 states: the TMA-fed kernel multiplies |psi|^2 by a single-precision diagonal -- 6e-8 relative)
    z = qc.B200State(qc.zero_state_tensor(dtype, nbits))
    cf_tol = 1e-12 if dtype == np.complex128 else 1e-6
    e_tfim, e_east = -((nbits - 1) + 0.1 * nbits), H.B * (1 + (nbits - 1) * (1 + H.c)) + H.c
    assert abs(qc.measure_(None, qc.TFIMHamiltonian(1.0, 0.1, 0.5), z) - e_tfim) < cf_tol * max(1.0, abs(e_tfim))
    assert abs(qc.measure_(None, H, z) - e_east) < cf_tol * max(1.0, abs(e_east))


# ---- the TMA-fed persistent expectation kernel (flip_tile.cuh, default from 14 / 15 qubits) against the oracle, every
# level structure: one window level (n <= 21 / 22), chunk-interleaved levels 0 + 1, a third level (n > 21 / 22)
@pytest.mark.parametrize("n", [14, 15, 16, 19, 21, 22, 23, 25])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_measure_tma_vs_oracle(oracle, n, dtype):
    rng = np.random.default_rng(4100 + n)
    psi = rand_state(rng, n, dtype)
    st = qc.B200State(psi)
    ref_psi = psi.astype(np.complex128)
    tol = TOL[dtype]
    for H, ref in [(qc.TFIMHamiltonian(1.0, 0.1, 0.5), oracle.measure_tfim(ref_psi, 1.0, 0.1, 0.5).real),
                   (qc.TFIMHamiltonian(0.7, -0.3, 1.4), oracle.measure_tfim(ref_psi, 0.7, -0.3, 1.4).real),
                   (qc.EastModelHamiltonian(0.1, -0.1), None)]:
        if ref is None:
            ref = oracle.measure_east(ref_psi, H.c, H.A, H.B).real
        qc.set_option("measure_tma", 1)
        E = qc.measure_(None, H, st)
        levels = qc.get_stat("passes")
        qc.set_option("measure_tma", 0)
        E_old = qc.measure_(None, H, st)
        qc.set_option("measure_tma", 1)
        assert abs(E - ref) < tol * max(1.0, abs(ref)), (type(H).__name__, E, ref)
        assert abs(E - E_old) < tol * max(1.0, abs(ref))
        tba = 12 if dtype == np.complex128 else 13
        assert levels == (1 + -(-(n - tba) // 9) if n >= tba + 2 else levels)


@pytest.mark.parametrize("n,J,g", [(11, 1.0, 0.5), (16, 1.0, 1.4), (18, 0.7, 1.0)])
def test_tfim_ground_state_against_free_fermion_solution(n, J, g):
    """A known answer from outside the code base (no oracle involved): the ground energy of the open chain
    -J (sum ZZ + g sum X) (src/hamiltonians/tfim.jl:17-54 with h = 0) is E0 = -sum of the singular values of Jg I - J S
    (Jordan-Wigner / Lieb-Schultz-Mattis, S = upper shift).  The ground vector comes from a Lanczos run on an index-arithmetic
    H psi; the GPU's expectation value of it (gather kernel at 11 q, TMA-fed kernel above) must be E0 in both dtypes."""
    from scipy.sparse.linalg import LinearOperator, eigsh

    E0 = -float(np.sum(np.linalg.svd(J * g * np.eye(n) - J * np.eye(n, k=1), compute_uv=False)))
    idx = np.arange(2 ** n, dtype=np.int64)
    z = [1 - 2 * ((idx >> i) & 1) for i in range(n)]
    diag = -J * sum(z[i] * z[i + 1] for i in range(n - 1)).astype(np.float64)

    def matvec(v):
        v = np.asarray(v).reshape(-1)
        return diag * v - J * g * sum(v[idx ^ (1 << i)] for i in range(n))

    w, v = eigsh(LinearOperator((2 ** n, 2 ** n), matvec=matvec, dtype=np.float64), k=1, which="SA", tol=1e-12)
    assert abs(w[0] - E0) < 1e-9 * n  # the Lanczos run found the free-fermion ground energy
    gs = (v[:, 0] / np.linalg.norm(v[:, 0])).astype(np.complex128)
    H = qc.TFIMHamiltonian(J, 0.0, g)
    assert abs(qc.measure_(None, H, qc.B200State(gs)) - E0) < 1e-10 * n
    assert abs(qc.measure_(None, H, qc.B200State(gs.astype(np.complex64))) - E0) < 1e-5 * n


@pytest.mark.parametrize("n,dtype", [(13, np.complex128), (17, np.complex64), (21, np.complex128)])
def test_measure_persistent_option(oracle, n, dtype):
    """Persistent form of the tiled expectation value (option measure_persistent) == the default one-tile-per-CTA form."""
    qc.set_option("measure_tma", 0)  # (the default TMA-fed kernel would take these sizes)
    rng = np.random.default_rng(5 * n)
    st = qc.B200State(rand_state(rng, n).astype(dtype))
    try:
        for H in (qc.TFIMHamiltonian(1.0, 0.1, 0.5), qc.EastModelHamiltonian(0.1, -0.1)):
            qc.set_option("measure_persistent", 0)
            E0 = qc.measure(H, st)
            qc.set_option("measure_persistent", 1)
            E1 = qc.measure(H, st)
            assert abs(E1 - E0) < (1e-12 if dtype == np.complex128 else 1e-6) * max(1.0, abs(E0))
    finally:
        qc.set_option("measure_persistent", 0)


def test_measure_with_circuit(oracle):
    rng = np.random.default_rng(17)
    n, layers = 12, 6
    c = circuit(n, layers, rng)
    psi0 = qc.zero_state_tensor(n)
    H = qc.TFIMHamiltonian(1.0, 0.2, 0.5)
    ref = oracle.measure_tfim(oracle.apply_brickwork(onp.zero_state(n), n, layers, c.gate_angles), 1.0, 0.2, 0.5).real
    assert abs(qc.measure(H, psi0, c) - ref) < 1e-12 * abs(ref)
    st = qc.B200State(psi0)
    assert abs(qc.measure(H, st, c) - ref) < 1e-12 * abs(ref)
    np.testing.assert_array_equal(st.to_numpy(), onp.zero_state(n))  # input not modified


# ---- src/circuits.jl:19-42 : apply!(cache, psi', psi, circuit) with distinct states -- psi' = circuit(psi), psi untouched;
# the first fused pass reads psi and writes psi' (no copy); small states and empty circuits fall back to a copy
@pytest.mark.parametrize("n,layers,dtype", [(4, 3, np.complex128), (9, 2, np.complex128), (13, 5, np.complex128), (16, 7, np.complex64), (12, 0, np.complex128)])
def test_out_of_place_circuit_apply(oracle, n, layers, dtype):
    rng = np.random.default_rng(19 * n + layers)
    c = circuit(n, max(layers, 1), rng)
    if layers == 0:
        c = qc.GenericBrickworkCircuit(n, 0)
    psi_host = rand_state(rng, n).astype(dtype)
    psi, out = qc.B200State(psi_host), qc.B200State(n, dtype)
    res = qc.apply_(qc.construct_apply_cache(psi), out, psi, c)
    assert res is out
    np.testing.assert_array_equal(psi.to_numpy(), psi_host)  # the input is untouched
    ref = oracle.apply_brickwork(psi_host.astype(np.complex128), n, c.nlayers, c.gate_angles) if c.ngates else psi_host.astype(np.complex128)
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    assert rel(out.to_numpy().astype(np.complex128), ref) < tol
    inplace = qc.B200State(psi_host)
    qc.apply_(inplace, inplace, c)
    np.testing.assert_array_equal(out.to_numpy(), inplace.to_numpy())  # same passes, same arithmetic
    st = qc.apply(psi, c)  # apply(psi, circuit) on a device state
    assert st is not psi
    np.testing.assert_array_equal(st.to_numpy(), out.to_numpy())
    np.testing.assert_array_equal(psi.to_numpy(), psi_host)


# ---- src/circuits.jl:49-62 : measure(H::AbstractMatrix, psi[, circuit]) on the device (CSR sweep)
@pytest.mark.parametrize("n", [1, 3, 5, 8, 11])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_measure_matrix_hamiltonian(oracle, n, dtype):
    rng = np.random.default_rng(49 * n)
    psi = rand_state(rng, n).astype(dtype)
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    Hs = qc.build_sparse_tfim_hamiltonian(n, 1.0, 0.1, 0.5)
    ref = onp.measure_matrix(Hs.toarray(), psi.astype(np.complex128))
    for H in (Hs, Hs.toarray()) if n <= 8 else (Hs,):  # sparse and dense forms
        assert abs(qc.measure(H, psi) - ref) < tol * max(1.0, abs(ref))
        assert abs(qc.measure_(None, H, qc.B200State(psi)) - ref) < tol * max(1.0, abs(ref))
    # examples/test_gradients.jl:10-28 : the kernel Hamiltonian and the sparse matrix give the same energy
    assert abs(qc.measure(qc.TFIMHamiltonian(1.0, 0.1, 0.5), psi) - qc.measure(Hs, psi)) < tol * max(1.0, abs(ref))
    if 2 <= n <= 8:  # test/correctness_tests.jl:126-150
        He = qc.EastModelHamiltonian(0.1, -0.1)
        assert abs(qc.measure(qc.mat(He, n), psi) - qc.measure(He, psi)) < tol * max(1.0, abs(ref))
        # a complex Hermitian matrix (not in the reference's examples; exercises the imaginary parts)
        A = rng.standard_normal((2 ** n, 2 ** n)) + 1j * rng.standard_normal((2 ** n, 2 ** n))
        A = A + A.conj().T
        refA = onp.measure_matrix(A, psi.astype(np.complex128))
        assert abs(qc.measure(A, psi) - refA) < tol * max(1.0, abs(refA)) * 2 ** n


def test_measure_matrix_with_circuit_and_errors(oracle):
    rng = np.random.default_rng(62)
    n, layers = 6, 4
    c = circuit(n, layers, rng, 0.01)
    Hs = qc.build_sparse_tfim_hamiltonian(n, 1.0, 0.1, 0.5)
    psi0 = onp.zero_state(n)
    ref = onp.measure_matrix(Hs.toarray(), oracle.apply_brickwork(psi0, n, layers, c.gate_angles))
    assert abs(qc.measure(Hs, psi0, c) - ref) < 1e-12 * abs(ref)  # src/circuits.jl:59-62
    with pytest.raises(qc.QCBError):  # wrong size
        qc.measure(Hs, onp.zero_state(n + 1))
    with pytest.raises(AssertionError):  # @assert imag(measurement) < 1e-12   src/circuits.jl:53
        qc.measure(1j * np.eye(2 ** n), rand_state(rng, n))


# ---- src/gradients.jl:88-152 : adjoint sweep == the reference's parameter-shift sweep
@pytest.mark.parametrize("n,layers,scale", [(2, 3, None), (4, 3, 0.1), (5, 4, None), (8, 4, 0.01), (9, 3, None), (10, 4, None), (12, 3, None)])
@pytest.mark.parametrize("ham", ["tfim", "east"])
def test_gradients_vs_reference_algorithm(oracle, n, layers, scale, ham):
    rng = np.random.default_rng(88 * n + layers)
    c = circuit(n, layers, rng, scale)
    if ham == "tfim":
        H, kind = qc.TFIMHamiltonian(1.0, 0.1, 0.5), 0
    else:
        H, kind = qc.EastModelHamiltonian(0.1, -0.1), 1
    psi0 = rand_state(rng, n) if n % 2 else onp.zero_state(n)
    E_ref, g_ref = oracle.gradients_brickwork(psi0, n, layers, c.gate_angles, kind, H.params())
    E, g = qc.gradients(H, psi0, c, calculate_energy=True)
    assert g.shape == (15, c.ngates)
    assert abs(E - E_ref) < 1e-12 * max(1.0, abs(E_ref))
    scale_g = max(np.max(np.abs(g_ref)), 1e-6)
    assert np.max(np.abs(g - g_ref)) < 1e-12 * max(scale_g, 1.0) + 1e-12 * scale_g
    g_only = qc.gradients(H, qc.B200State(psi0), c)
    np.testing.assert_allclose(g_only, g, atol=1e-15)
    # ComplexF32 state: compare with the ComplexF64 reference numbers at the c64 tolerance
    E32, g32 = qc.gradients(H, psi0.astype(np.complex64), c, calculate_energy=True)
    assert abs(E32 - E_ref) < 1e-5 * max(1.0, abs(E_ref))
    assert np.max(np.abs(g32 - g_ref)) < 1e-5 * max(scale_g, 1.0)


# ---- the same on the cluster-resident kernel (small_circuit.cuh: forward circuit, H psi, backward sweep and environments in one
# launch) == the reference's shift sweep == the pass-based adjoint sweep
@pytest.mark.parametrize("n,layers,scale", [(2, 3, None), (3, 5, None), (4, 3, 0.1), (6, 7, None), (8, 4, 0.01), (9, 6, None), (10, 4, None), (11, 5, None), (12, 3, None)])
@pytest.mark.parametrize("ham", ["tfim", "east"])
def test_cluster_gradient_kernel(oracle, n, layers, scale, ham):
    rng = np.random.default_rng(99 * n + layers)
    c = circuit(n, layers, rng, scale)
    if ham == "tfim":
        H, kind = qc.TFIMHamiltonian(1.0, 0.1, 0.5), 0
    else:
        H, kind = qc.EastModelHamiltonian(0.1, -0.1), 1
    psi0 = rand_state(rng, n) if n % 2 else onp.zero_state(n)
    E_ref, g_ref = oracle.gradients_brickwork(psi0, n, layers, c.gate_angles, kind, H.params())
    E_p, g_p = qc.gradients(H, psi0, c, calculate_energy=True)  # pass-based sweep (fixture: small_cluster = 0)
    qc.set_option("small_cluster", 12)
    try:
        k0 = qc.get_stat("cluster_circuits")
        E, g = qc.gradients(H, psi0, c, calculate_energy=True)
        assert qc.get_stat("cluster_circuits") == k0 + 1 and qc.get_stat("launches") == 3  # gradient kernel, finish, contraction
        g_only = qc.gradients(H, qc.B200State(psi0), c)
        E32, g32 = qc.gradients(H, psi0.astype(np.complex64), c, calculate_energy=True)
    finally:
        qc.set_option("small_cluster", 0)
    scale_g = max(np.max(np.abs(g_ref)), 1e-6)
    assert abs(E - E_ref) < 1e-12 * max(1.0, abs(E_ref)) and abs(E - E_p) < 1e-12 * max(1.0, abs(E_ref))
    assert np.max(np.abs(g - g_ref)) < 1e-12 * max(scale_g, 1.0) + 1e-12 * scale_g
    assert np.max(np.abs(g - g_p)) < 1e-12 * max(scale_g, 1.0) + 1e-12 * scale_g
    np.testing.assert_allclose(g_only, g, atol=1e-15)
    assert abs(E32 - E_ref) < 1e-5 * max(1.0, abs(E_ref))
    assert np.max(np.abs(g32 - g_ref)) < 1e-5 * max(scale_g, 1.0)


# (compute-sanitizer is closed on the GPU pool: a race in the cluster kernels' barrier / buffer-swap protocol would show up as run-to-run
# differences -- every sum is taken in a fixed order, so the results must repeat bit for bit)
@pytest.mark.parametrize("n,layers", [(12, 20), (11, 9), (10, 12)])
def test_cluster_kernels_repeat_bit_for_bit(n, layers):
    rng = np.random.default_rng(7 * n + layers)
    c = circuit(n, layers, rng)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    psi0 = qc.B200State(rand_state(rng, n))
    qc.set_option("small_cluster", 12)
    try:
        E0, g0 = qc.gradients(H, psi0, c, calculate_energy=True)
        out0 = qc.apply(psi0, c).to_numpy()
        for _ in range(40):
            E, g = qc.gradients(H, psi0, c, calculate_energy=True)
            assert E == E0 and np.array_equal(g, g0)
            assert np.array_equal(qc.apply(psi0, c).to_numpy(), out0)
    finally:
        qc.set_option("small_cluster", 0)


@pytest.mark.parametrize("n,layers,atb", [(9, 3, 0), (12, 4, 0), (13, 5, 9), (16, 6, 10), (18, 4, 11), (19, 7, 12), (22, 4, 0)])
def test_gradients_f32_tensor_core_pass(oracle, n, layers, atb):
    """ComplexF32 backward pass with both register windows resident, FFMA2 outer products and a warp transpose-reduction
    (adjoint_f32.cuh, default) vs the quad-per-thread pass (adjoint.cuh, option adjoint_f32 = 0), vs the ComplexF64 gradient of
    the same circuit, and at n <= 13 vs the oracle's restatement of the reference's shift sweep.  Tolerance: BASELINE.json's
    1e-5 relative to the gradient's max-norm; the two ComplexF32 passes differ only by FP32 summation order (2e-6)."""
    rng = np.random.default_rng(2000 * n + layers)
    c = circuit(n, layers, rng)
    H = qc.TFIMHamiltonian(1.0, 0.9045, 1.4) if n % 2 else qc.EastModelHamiltonian(0.1, -0.1)
    v = rand_state(rng, n)
    E64, g64 = qc.gradients(H, qc.B200State(v), c, calculate_energy=True)
    psi32 = qc.B200State(v.astype(np.complex64))
    qc.set_option("adjoint_tile_bits", atb)
    qc.set_option("adjoint_f32", 0)
    E0, g0 = qc.gradients(H, psi32, c, calculate_energy=True)
    qc.set_option("adjoint_f32", 1)
    E1, g1 = qc.gradients(H, psi32, c, calculate_energy=True)
    scale_g = max(np.max(np.abs(g64)), 1.0)
    assert E1 == E0  # the forward pass and H psi do not depend on the option
    assert np.max(np.abs(g1 - g0)) < 2e-6 * scale_g
    assert abs(E1 - E64) < 1e-5 * max(1.0, abs(E64))
    assert np.max(np.abs(g1 - g64)) < 1e-5 * scale_g
    if n <= 13:
        _, g_ref = oracle.gradients_brickwork(v, n, layers, c.gate_angles, 0 if n % 2 else 1, H.params())
        assert np.max(np.abs(g1 - g_ref)) < 1e-5 * scale_g
    qc.set_option("adjoint_f32", 1)


@pytest.mark.parametrize("n,layers,ham", [(11, 4, "tfim"), (16, 3, "east"), (20, 2, "tfim")])
def test_gradients_against_central_differences(n, layers, ham):
    """The definition of the derivative, no oracle involved: dE/dtheta of the adjoint sweep (cluster kernel at 11 q; fused
    tensor-core backward passes above) against (E(theta + eps) - E(theta - eps)) / 2 eps of the GPU's own forward circuit +
    expectation value, for angles of the first, a middle and the last gate (src/gradients.jl:88-152 computes the same
    numbers by the exact shift rule)."""
    rng = np.random.default_rng(9000 + n)
    c = circuit(n, layers, rng)
    H = qc.TFIMHamiltonian(1.0, 0.9045, 1.4) if ham == "tfim" else qc.EastModelHamiltonian(0.1, -0.1)
    psi0 = qc.B200State(rand_state(rng, n))
    g = qc.gradients(H, psi0, c)
    G = c.gate_angles.shape[1]
    eps = 1e-5
    for k, gate in [(0, 0), (7, G // 2), (14, G - 1), (3, 1), (11, G - 2)]:
        keep = c.gate_angles[k, gate]
        c.gate_angles[k, gate] = keep + eps
        ep = qc.measure(H, psi0, c)
        c.gate_angles[k, gate] = keep - eps
        em = qc.measure(H, psi0, c)
        c.gate_angles[k, gate] = keep
        assert abs((ep - em) / (2 * eps) - g[k, gate]) < 2e-8 * max(1.0, np.max(np.abs(g))), (k, gate)


@pytest.mark.parametrize("n,layers", [(13, 3), (14, 2), (17, 3), (21, 2), (24, 2)])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_gradients_tiled_hamiltonian_sweep(oracle, n, layers, dtype):
    """lambda = H psi with the low-bit flips served from shared memory (k_apply_ham_tiled; default for ComplexF32) vs the plain
    gather (option ham_tile = 0): the same sums in the same order, so the gradient is bit-identical and the energy agrees to the
    rounding of its partial sums; TFIM and East; at n <= 14 also against the oracle's restatement of the reference's sweep
    (src/gradients.jl:104 measure!, src/hamiltonians/tfim.jl:17-54, east_model.jl:30-52)."""
    rng = np.random.default_rng(4000 + 17 * n + layers)
    c = circuit(n, layers, rng)
    H = qc.EastModelHamiltonian(0.1, -0.1) if n % 2 else qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
    v = rand_state(rng, n, dtype)
    psi0 = qc.B200State(v)
    try:
        qc.set_option("ham_tile", 0)
        E0, g0 = qc.gradients(H, psi0, c, calculate_energy=True)
        for mode in (1, 2):  # tiles for ComplexF32 only (default); for both types
            qc.set_option("ham_tile", mode)
            E1, g1 = qc.gradients(H, psi0, c, calculate_energy=True)
            assert np.array_equal(g1, g0), mode
            assert abs(E1 - E0) < 1e-13 * max(1.0, abs(E0)), mode
    finally:
        qc.set_option("ham_tile", HAM_TILE_DEFAULT)
    if n <= 14:
        tol = 1e-12 if dtype == np.complex128 else 1e-5
        E_ref, g_ref = oracle.gradients_brickwork(v.astype(np.complex128), n, layers, c.gate_angles, 1 if n % 2 else 0, H.params())
        assert np.max(np.abs(g1 - g_ref)) < tol * max(1.0, np.max(np.abs(g_ref)))


@pytest.mark.parametrize("n,layers", [(9, 2), (12, 20), (14, 9), (17, 5), (20, 4), (23, 3)])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_gradients_deferred_reduction(n, layers, dtype):
    """The accumulator rows of all backward passes reduced by two launches after the sweep (default) == one reduction per
    pass (option adjoint_defer = 0); same fixed summation order, so the numbers agree to rounding of the final adds.  Every
    backward-pass kernel: tensor-core ComplexF64 (5 and 6 CTAs per SM), scalar ComplexF64, both ComplexF32 forms."""
    rng = np.random.default_rng(3000 + 31 * n + layers)
    c = circuit(n, layers, rng)
    H = qc.EastModelHamiltonian(0.1, -0.1) if n % 2 else qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
    psi0 = qc.B200State(rand_state(rng, n, dtype))
    variants = [("adjoint_mma", 1), ("adjoint_mma", 0)] if dtype == np.complex128 else [("adjoint_f32", 1), ("adjoint_f32", 0)]
    for key, val in variants:
        qc.set_option(key, val)
        for ctas in ((5, 6) if (key, val) == ("adjoint_mma", 1) else (0,)):
            qc.set_option("adjoint_ctas", ctas)
            qc.set_option("adjoint_defer", 0)
            E0, g0 = qc.gradients(H, psi0, c, calculate_energy=True)
            l0 = qc.get_stat("launches")
            qc.set_option("adjoint_defer", 1)
            E1, g1 = qc.gradients(H, psi0, c, calculate_energy=True)
            l1 = qc.get_stat("launches")
            assert E1 == E0
            assert np.max(np.abs(g1 - g0)) < 1e-13 * max(1.0, np.max(np.abs(g0)))
            if n >= 12 and val == 1:
                assert l1 < l0  # fewer launches per gradient (the default kernels note their gate ids themselves)
    qc.set_option("adjoint_mma", 1)
    qc.set_option("adjoint_f32", 1)
    qc.set_option("adjoint_ctas", 5)


@pytest.mark.parametrize("n,layers,atb", [(13, 5, 0), (16, 6, 9), (18, 4, 10), (18, 7, 11), (20, 3, 12), (22, 4, 0)])
def test_gradients_tensor_core_pass_equals_scalar_pass(oracle, n, layers, atb):
    """ComplexF64 backward pass on the FP64 tensor cores (adjoint_mma.cuh, default) vs the scalar-FMA pass (adjoint.cuh),
    every tile size; at n <= 13 also against the oracle's restatement of the reference's shift sweep."""
    rng = np.random.default_rng(1000 * n + layers)
    c = circuit(n, layers, rng)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5) if n % 2 else qc.EastModelHamiltonian(0.1, -0.1)
    psi0 = qc.B200State(rand_state(rng, n))
    qc.set_option("adjoint_tile_bits", atb)
    qc.set_option("adjoint_mma", 0)
    E0, g0 = qc.gradients(H, psi0, c, calculate_energy=True)
    qc.set_option("adjoint_mma", 1)
    E1, g1 = qc.gradients(H, psi0, c, calculate_energy=True)
    scale_g = max(np.max(np.abs(g0)), 1e-6)
    assert abs(E1 - E0) < 1e-13 * max(1.0, abs(E0))
    assert np.max(np.abs(g1 - g0)) < 1e-12 * max(scale_g, 1.0)
    if n <= 13:
        _, g_ref = oracle.gradients_brickwork(psi0.to_numpy(), n, layers, c.gate_angles, 0 if n % 2 else 1, H.params())
        assert np.max(np.abs(g1 - g_ref)) < 1e-12 * max(scale_g, 1.0)
    # the per-thread gate programs built on the device == the table the CPU replay builds on the host (uploaded): same kernel,
    # same inputs, so the numbers must be identical
    qc.set_option("adjoint_prog_host", 1)
    E2, g2 = qc.gradients(H, psi0, c, calculate_energy=True)
    assert E2 == E1 and np.array_equal(g2, g1)
    # the 16-lane cooperative contraction of the environments (default) == one thread per gate layer (gates_hd.cuh on
    # both sides: the same algebra the CPU tests check against gates_host.hpp)
    qc.set_option("contract_coop", 0)
    E3, g3 = qc.gradients(H, psi0, c, calculate_energy=True)
    assert E3 == E1 and np.max(np.abs(g3 - g1)) < 1e-13 * max(scale_g, 1.0)
    for dt in (np.complex64,):  # and for the scalar backward pass (environments, not post-gate matrices)
        psi32 = qc.B200State(psi0.to_numpy().astype(dt))
        ga = qc.gradients(H, psi32, c)
        qc.set_option("contract_coop", 1)
        gb = qc.gradients(H, psi32, c)
        assert np.max(np.abs(ga - gb)) < 1e-13 * max(scale_g, 1.0)


# ---- examples/test_gradients.jl:10-28 : gradients with the Hamiltonian as an explicit (sparse) matrix == kernel Hamiltonian
@pytest.mark.parametrize("n,layers,dtype", [(8, 4, np.complex128), (5, 3, np.complex128), (10, 3, np.complex128), (11, 2, np.complex64)])
def test_gradients_matrix_hamiltonian(oracle, n, layers, dtype):
    rng = np.random.default_rng(25 * n + layers)
    c = circuit(n, layers, rng, 0.01 if n == 8 else None)
    J, h, g = 1.0, 0.1, 0.5
    Hs, Heff = qc.build_sparse_tfim_hamiltonian(n, J, h, g), qc.TFIMHamiltonian(J, h, g)
    psi0 = qc.zero_state_tensor(n).astype(dtype) if n == 8 else rand_state(rng, n).astype(dtype)
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    E, grads = qc.gradients(Heff, psi0, c, calculate_energy=True)
    E_m, grads_m = qc.gradients(Hs, psi0, c, calculate_energy=True)
    assert abs(E - E_m) < tol * max(1.0, abs(E))  # @test E ≈ E_m
    assert np.max(np.abs(grads - grads_m)) < tol * max(1.0, np.max(np.abs(grads)))  # @test grads ≈ grads_m
    if n <= 8:  # dense form, and a Hermitian matrix that is no kernel Hamiltonian, against the shift rule on the oracle
        np.testing.assert_allclose(qc.gradients(Hs.toarray(), psi0, c), grads_m, atol=1e-13)
        A = rng.standard_normal((2 ** n, 2 ** n)) + 1j * rng.standard_normal((2 ** n, 2 ** n))
        A = (A + A.conj().T) / 2 ** n
        gA = qc.gradients(A, psi0, c)
        v0 = psi0.reshape(-1, order="F").astype(np.complex128)
        flat = c.gate_angles.reshape(-1, order="F")
        for idx in rng.choice(flat.size, 5, replace=False):
            es = []
            for sh in (np.pi / 2, -np.pi / 2):
                cc = qc.reconstruct(c, int(idx), sh)
                es.append(onp.measure_matrix(A, oracle.apply_brickwork(v0, n, layers, cc.gate_angles)))
            assert abs((es[0] - es[1]) / 2 - gA.reshape(-1, order="F")[idx]) < 1e-12


def test_gradients_c2_shape_against_shift_subset(oracle):
    """20-qubit TFIM config (BASELINE C2): adjoint gradient vs the shift rule evaluated on the GPU for a few angles,
    and vs the oracle's forward energy."""
    rng = np.random.default_rng(2020)
    n, layers = 20, 4
    c = circuit(n, layers, rng, 0.01)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    psi0 = qc.B200State(qc.zero_state_tensor(n))
    E, g = qc.gradients(H, psi0, c, calculate_energy=True)
    ref_psi = oracle.apply_brickwork(onp.zero_state(n), n, layers, c.gate_angles)
    assert abs(E - oracle.measure_tfim(ref_psi, 1.0, 0.1, 0.5).real) < 1e-12 * abs(E)
    flat = c.gate_angles.reshape(-1, order="F")
    for idx in rng.choice(flat.size, 6, replace=False):
        ep = qc.measure(H, psi0, qc.reconstruct(c, int(idx), np.pi / 2))
        em = qc.measure(H, psi0, qc.reconstruct(c, int(idx), -np.pi / 2))
        assert abs((ep - em) / 2 - g.reshape(-1, order="F")[idx]) < 1e-11


@pytest.mark.parametrize("small_cluster", [0, 12])  # pass-based sweep / cluster-resident kernels (small_circuit.cuh)
def test_optimise_trajectory(oracle, small_cluster):
    rng = np.random.default_rng(7)
    n, layers, epochs, lr = 8, 3, 6, 0.01
    c = circuit(n, layers, rng, 0.01)
    a0 = c.gate_angles.copy()
    H = qc.TFIMHamiltonian(1.0, 0.2, 0.5)
    e_ref, a_ref = oracle.optimise_brickwork(onp.zero_state(n), n, layers, a0, 0, H.params(), epochs, lr)
    qc.set_option("small_cluster", small_cluster)
    try:
        k0 = qc.get_stat("cluster_circuits")
        energies = qc.optimise_(c, H, qc.zero_state_tensor(n), epochs, lr)
        assert (qc.get_stat("cluster_circuits") > k0) == bool(small_cluster)
    finally:
        qc.set_option("small_cluster", 0)
    np.testing.assert_allclose(energies, e_ref, atol=1e-10)
    np.testing.assert_allclose(c.gate_angles, a_ref, atol=1e-10)


def test_propagate_forwards_and_adjoint_range(oracle):
    rng = np.random.default_rng(54)
    n, layers = 11, 5
    c = circuit(n, layers, rng)
    psi0 = rand_state(rng, n)
    st = qc.B200State(psi0)
    buf = st.similar()
    gate_idx = 9  # 1-based, as in Julia
    res, other = qc.propagate_forwards_(qc.construct_apply_cache(st), buf, st, c, 2, 4, gate_idx)
    ref = oracle.apply_brickwork(psi0, n, layers, c.gate_angles, g_begin=gate_idx - 1)
    assert rel(res.to_numpy(), ref) < 1e-12
    from qcb200.apply import _apply_circuit_inplace

    _apply_circuit_inplace(res, c, first=gate_idx - 1, adjoint=True)  # undo
    assert rel(res.to_numpy(), psi0) < 1e-12


# ---- src/apply.jl:144-182 : right_apply_gate!(M', M, gate) (operator evolution M' = M U) as one gate pass on the 2n-qubit tensor
@pytest.mark.parametrize("nbits", [2, 3, 6, 9])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_right_apply_gate(nbits, dtype):
    rng = np.random.default_rng(1440 + nbits)
    M = (rng.standard_normal((2 ** nbits, 2 ** nbits)) + 1j * rng.standard_normal((2 ** nbits, 2 ** nbits))).astype(dtype)
    for K in sorted({1, nbits - 1, max(1, nbits // 2)}):
        u4 = rand_nonunitary(rng, 4)
        gate = qc.Localised2SpinAdjGate(qc.Generic2SpinGate(u4.reshape(2, 2, 2, 2, order="F")), K)
        out = np.empty_like(M)
        res = qc.right_apply_gate_(out, M, gate)
        assert res is out
        ref = onp.right_apply_gate(M.astype(np.complex128), u4, K)
        assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < TOL[dtype]
    with pytest.raises(ValueError):
        qc.right_apply_gate_(np.empty((4, 2), dtype=dtype), M, gate)  # "M′ must be the same size as M"
    with pytest.raises(AssertionError):
        qc.right_apply_gate_(np.empty_like(M), M, qc.Localised2SpinAdjGate(qc.Generic2SpinGate(u4.reshape(2, 2, 2, 2, order="F")), nbits))


# ---- the TMA-fed persistent form of the gate pass (tile_tma.cuh, default for ComplexF64) == the LDG / STS form, bit for bit
# (same stage programs and arithmetic order; only the tile movement and the shared-memory swizzle differ), and == oracle
@pytest.mark.parametrize("n,layers,cap", [(11, 8, 0), (12, 20, 0), (14, 9, 3), (17, 12, 0), (20, 6, 2), (22, 10, 0), (24, 5, 4)])
def test_tma_gate_pass_equals_ldg_pass(oracle, n, layers, cap):
    rng = np.random.default_rng(1100 + 7 * n + layers)
    c = circuit(n, layers, rng)
    psi0 = rand_state(rng, n)
    if cap:
        qc.set_option("max_gates_per_pass", cap)
    qc.set_option("f64_3m_min_gates", 1 << 30)  # the TMA kernel runs the ordinary arithmetic: compare like with like
    outs = {}
    for mode in (1, 0):
        qc.set_option("pass_tma", mode)
        st = qc.B200State(psi0)
        qc.apply_(st, st, c)
        assert (qc.get_stat("tma_passes") == qc.get_stat("passes")) if mode else (qc.get_stat("tma_passes") == 0)
        outs[mode] = st.to_numpy()
        # out of place: the first pass reads psi and writes psi' through two tensor maps
        dst = st.similar()
        src = qc.B200State(psi0)
        qc.apply_(dst, src, c)
        assert np.array_equal(dst.to_numpy(), outs[mode]) and np.array_equal(src.to_numpy(), psi0)
    qc.set_option("pass_tma", 0)
    assert np.array_equal(outs[0], outs[1])
    if n <= 22:
        assert rel(outs[1], oracle.apply_brickwork(psi0, n, layers, c.gate_angles)) < 1e-12


# ---- small states: the whole gate list in one launch on a thread-block cluster (small_circuit.cuh) == tile passes == oracle
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,layers", [(2, 3), (3, 4), (5, 7), (8, 6), (9, 12), (10, 9), (11, 8), (12, 20), (13, 11), (14, 6), (15, 5), (16, 4)])
def test_cluster_circuit_kernel(oracle, n, layers, dtype):
    if n > (15 if dtype == np.complex128 else 16):
        pytest.skip("two 2^n-amplitude buffers do not fit the shared memory of 8 CTAs")
    rng = np.random.default_rng(4400 + 17 * n + layers)
    c = circuit(n, layers, rng)
    psi0 = rand_state(rng, n, dtype)
    ref = oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, c.gate_angles)
    qc.set_option("small_cluster", 1)
    try:
        k0 = qc.get_stat("cluster_circuits")
        st = qc.B200State(psi0)
        qc.apply_(st, st, c)
        assert qc.get_stat("cluster_circuits") == k0 + 1 and qc.get_stat("launches") == 1
        out = st.to_numpy()
        dst = st.similar()  # out of place: reads psi, writes psi'
        src = qc.B200State(psi0)
        qc.apply_(dst, src, c)
        assert np.array_equal(dst.to_numpy(), out) and np.array_equal(src.to_numpy(), psi0)
        # a gate list in random positions, some repeated (non-unitary matrices, test/correctness_tests.jl:86-124 style)
        ks = [int(k) for k in rng.integers(1, n, size=9)]
        us = [rand_nonunitary(rng, 4) for _ in ks]
        gl = [qc.Localised2SpinAdjGate(qc.Generic2SpinGate(u.reshape(2, 2, 2, 2, order="F")), k) for u, k in zip(us, ks)]
        k1 = qc.get_stat("cluster_circuits")
        got = qc.apply(psi0, gl)
        assert qc.get_stat("cluster_circuits") == k1 + 1
        want = psi0.astype(np.complex128)
        for u, k in zip(us, ks):
            want = oracle.apply_2q(want, u, k)
        assert rel(np.asarray(got).reshape(-1, order="F").astype(np.complex128), want) < 20 * TOL[dtype]
    finally:
        qc.set_option("small_cluster", 0)
    assert rel(out.astype(np.complex128), ref) < TOL[dtype]
    st2 = qc.B200State(psi0)
    qc.apply_(st2, st2, c)  # tile passes
    assert rel(out.astype(np.complex128), st2.to_numpy().astype(np.complex128)) < (1e-13 if dtype == np.complex128 else 1e-5)


# ---- persistent gate pass with asynchronous tile loads (tile_async.cuh) == one tile per CTA (tile.cuh), bit for bit
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,layers,cap", [(11, 8, 0), (13, 20, 0), (17, 12, 3), (22, 10, 0), (24, 5, 1)])
def test_async_gate_pass_equals_sync_pass(oracle, n, layers, cap, dtype):
    rng = np.random.default_rng(1200 + 7 * n + layers)
    c = circuit(n, layers, rng)
    psi0 = rand_state(rng, n, dtype)
    if cap:
        qc.set_option("max_gates_per_pass", cap)
    outs = {}
    for mode in (1, 0):
        qc.set_option("pass_async", mode)
        a0 = qc.get_stat("async_passes")
        st = qc.B200State(psi0)
        qc.apply_(st, st, c)
        assert (qc.get_stat("async_passes") - a0 == qc.get_stat("passes")) if mode else (qc.get_stat("async_passes") == a0)
        outs[mode] = st.to_numpy()
        dst = st.similar()  # out of place: the first pass reads psi and writes psi'
        src = qc.B200State(psi0)
        qc.apply_(dst, src, c)
        assert np.array_equal(dst.to_numpy(), outs[mode]) and np.array_equal(src.to_numpy(), psi0)
    qc.set_option("pass_async", PASS_ASYNC_DEFAULT)
    assert np.array_equal(outs[0], outs[1])
    if n <= 22:
        assert rel(outs[1].astype(np.complex128), oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, c.gate_angles)) < TOL[dtype]


# ---- ComplexF64 arithmetic: the 3M form of the deep passes (tile.cuh: apply_window_gate_3m) == the ordinary form == oracle
@pytest.mark.parametrize("n,layers", [(12, 20), (16, 12), (22, 10)])
def test_3m_form_equals_ordinary_form(oracle, n, layers):
    rng = np.random.default_rng(3300 + n)
    c = circuit(n, layers, rng)
    psi0 = rand_state(rng, n)
    outs = {}
    for min_gates in (0, 1 << 30):
        qc.set_option("f64_3m_min_gates", min_gates)
        st = qc.B200State(psi0)
        qc.apply_(st, st, c)
        outs[min_gates] = st.to_numpy()
    qc.set_option("f64_3m_min_gates", 4)
    ref = oracle.apply_brickwork(psi0, n, layers, c.gate_angles)
    assert not np.array_equal(outs[0], outs[1 << 30])
    assert rel(outs[0], ref) < 1e-12 and rel(outs[1 << 30], ref) < 1e-12 and rel(outs[0], outs[1 << 30]) < 1e-13


# ---- size-independent properties at sizes the oracle cannot check in seconds
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_fused_equals_unfused_26q(dtype):
    rng = np.random.default_rng(26)
    n, layers = 26, 6
    c = circuit(n, layers, rng)
    a = qc.B200State(n, dtype)
    b = qc.B200State(n, dtype)
    qc.apply_(a, a, c)
    qc.set_option("force_simple", 1)
    qc.apply_(b, b, c)
    qc.set_option("force_simple", 0)
    va, vb = a.to_numpy(), b.to_numpy()
    assert rel(va, vb) < (1e-13 if dtype == np.complex128 else 2e-6)
    assert abs(a.norm2() - 1) < (1e-12 if dtype == np.complex128 else 1e-5)
    # un-applying the circuit returns |0...0>
    from qcb200.apply import _apply_circuit_inplace

    _apply_circuit_inplace(a, c, adjoint=True)
    v = a.to_numpy()
    assert abs(v[0] - 1) < (1e-12 if dtype == np.complex128 else 1e-5)
    assert np.linalg.norm(v[1:]) < (1e-12 if dtype == np.complex128 else 1e-5)


def test_full_size_30q_properties(oracle):
    """BASELINE config C3 size (30 qubits, ComplexF64): shallow prefix vs the oracle is too slow here, so check the
    invariants the domain offers: norm, TFIM energy of a product state, circuit then inverse circuit = identity."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs 40 GiB of free HBM")
    rng = np.random.default_rng(30)
    n, layers = 30, 8
    c = circuit(n, layers, rng)
    st = qc.B200State(n, np.complex128)
    qc.apply_(st, st, c)
    assert abs(st.norm2() - 1) < 1e-12
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    E = qc.measure_(None, H, st)
    assert -2 * n < E < 2 * n
    from qcb200.apply import _apply_circuit_inplace

    _apply_circuit_inplace(st, c, adjoint=True)
    assert abs(qc.measure_(None, H, st) + ((n - 1) + 0.1 * n)) < 1e-10
    head = st.tensor[:4].cpu().numpy()
    assert abs(head[0] - 1) < 1e-12 and np.all(np.abs(head[1:]) < 1e-12)


def test_host_buffer_pipeline_equals_device_path():
    """apply(psi::Array, circuit) pipelines H2D / D2H chunks with the chunk-local head and tail of the circuit
    (n >= 25); the result must equal the resident-state path bit for bit up to FMA-order rounding."""
    rng = np.random.default_rng(2525)
    for n, layers, dtype in [(25, 9, np.complex128), (26, 30, np.complex64)]:
        c = circuit(n, layers, rng)
        psi0 = rand_state(rng, n, dtype)
        out_host = qc.apply(psi0, c)
        st = qc.B200State(psi0)
        qc.apply_(st, st, c)
        ref = st.to_numpy()
        assert rel(out_host, ref) < (1e-13 if dtype == np.complex128 else 2e-6)
        assert abs(np.linalg.norm(out_host.astype(np.complex128)) ** 2 - 1) < (1e-12 if dtype == np.complex128 else 1e-5)


def test_hamiltonian_layer_autograd(oracle):
    """examples/nns/circuit_layer.jl shape: Dense(1 => nangles) -> x*pi -> reshape(15, G) -> HamiltonianLayer -> E;
    the reverse rule calls gradients!(...; calculate_energy=true).  float32 network parameters, like Flux."""
    import torch

    from qcb200.nn import HamiltonianLayer

    torch.manual_seed(0)
    n, layers = 10, 4
    G = qc.brickwork_num_gates(n, layers)
    H = qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
    layer = HamiltonianLayer(n, layers, qc.zero_state_tensor(n), H)
    dense = torch.nn.Linear(1, 15 * G, bias=False)
    x = torch.ones(1)
    angles = (dense(x) * np.pi).reshape(G, 15).T  # column g = gate g
    E = layer(angles)
    E.backward()
    a64 = angles.detach().double().numpy()
    E_ref, g_ref = oracle.gradients_brickwork(onp.zero_state(n), n, layers, a64, 0, H.params())
    assert abs(float(E) - E_ref) < 1e-5 * abs(E_ref)
    expect = (g_ref.T.reshape(-1) * np.pi)  # dE/dW = dE/dangle * pi * x
    got = dense.weight.grad.reshape(-1).double().numpy()
    assert np.max(np.abs(got - expect)) < 1e-5 * max(1.0, np.max(np.abs(expect)))
    with torch.no_grad():  # inference path: plain measure
        assert abs(float(layer(angles.detach())) - E_ref) < 1e-5 * abs(E_ref)


@pytest.mark.parametrize("N,nl", [(6, 1), (8, 2), (10, 2)])
def test_polar_optimise_statevector(N, nl):
    """SURVEY 8f rank 1 (examples/test_optimise.jl shape: identity start, TFIM ground state target): the GPU path
    (environment kernel + host SVD) follows the NumPy restatement of ext/MPSExt/polar_optimise.jl."""
    from qcb200.polar import polar_optimise

    hp = (1.0, 0.9045, 1.4)
    Hd = onp.dense_tfim(N, *hp)
    w, v = np.linalg.eigh(Hd)
    gs = v[:, 0].astype(np.complex128)
    ident = [[np.eye(4, dtype=np.complex128) for _ in range(N - 1)] for _ in range(nl)]
    its = 6
    c_ref, ov_ref, en_ref = onp.polar_optimise(ident, gs, lambda s: onp.measure_tfim(s, *hp).real, N, its)
    c_gpu, ov, en = polar_optimise(ident, gs, qc.TFIMHamiltonian(*hp), N, its)
    np.testing.assert_allclose(ov, ov_ref, atol=1e-9)
    np.testing.assert_allclose(en, en_ref, atol=1e-8)
    assert ov[-1] > ov[0] - 1e-12 and en[-1] < en[0] + 1e-9  # the sweep improves overlap and energy
    for lg, lr in zip(c_gpu, c_ref):
        for a, b in zip(lg, lr):
            m = a.reshape(4, 4, order="F")
            np.testing.assert_allclose(m @ m.conj().T, np.eye(4), atol=1e-12)


def test_polar_optimise_invariants():
    """The GPU path on an exactly representable target (the reference's algorithm guarantees these whatever the data): the
    exact circuit is a fixed point with overlap 1, and from the identity start the overlap never decreases."""
    from qcb200.polar import polar_optimise

    rng = np.random.default_rng(92)
    N, nl = 8, 2
    sites = [2 * idx - 1 for idx in range(1, N // 2 + 1)] + [2 * idx - N for idx in range(N // 2 + 1, N)]

    def rand_u():
        q, r = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))
        return q * (np.diag(r) / np.abs(np.diag(r)))
    exact = [[rand_u() for _ in range(N - 1)] for _ in range(nl)]
    target = onp.zero_state(N)
    for layer in exact:
        for g, site in zip(layer, sites):
            target = onp.apply_2q(target, g, site)
    H = qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
    _, ov, en = polar_optimise(exact, target, H, N, 2)
    np.testing.assert_allclose(ov, 1.0, atol=1e-12)
    np.testing.assert_allclose(en, onp.measure_tfim(target, 1.0, 0.9045, 1.4).real, atol=1e-11)
    ident = [[np.eye(4, dtype=np.complex128) for _ in range(N - 1)] for _ in range(nl)]
    _, ov, _ = polar_optimise(ident, target, H, N, 10)
    _, ov_ref, _ = onp.polar_optimise(ident, target, lambda s: 0.0, N, 10)
    assert np.all(np.diff(ov) > -1e-12) and ov[-1] > ov[0]
    np.testing.assert_allclose(ov, ov_ref, atol=1e-9)


def test_trim_releases_and_reallocates_work_buffers(oracle):
    import torch

    rng = np.random.default_rng(3)
    n, layers = 22, 3
    c = circuit(n, layers, rng, 0.1)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    psi0 = qc.B200State(qc.zero_state_tensor(n))
    g1 = qc.gradients(H, psi0, c)
    torch.cuda.synchronize()
    free_before, _ = torch.cuda.mem_get_info()
    qc.trim()
    free_after, _ = torch.cuda.mem_get_info()
    assert free_after >= free_before + (16 << n)  # the gradient work states are gone
    g2 = qc.gradients(H, psi0, c)  # and come back on demand
    np.testing.assert_array_equal(g1, g2)
