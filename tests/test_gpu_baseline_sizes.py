"""GPU parity at the sizes BASELINE.json names (configs C1, C3, C5), against the CPU oracle.

Everything goes through the C ABI (Python mirror of the reference API).  Tolerances are BASELINE.json's: 1e-12 relative
(ComplexF64), 1e-5 (ComplexF32).  The oracle is the C + OpenMP restatement of src/apply.jl:62-84, src/circuits.jl:19-31,
src/hamiltonians/tfim.jl:17-54 and src/gradients.jl:88-152; these tests are sized so that it finishes in about a minute
on the GPU box's host cores."""
import numpy as np
import pytest

import qcb200 as qc
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def circuit(n, layers, rng, scale=None):
    c = qc.GenericBrickworkCircuit(n, layers)
    c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape) if scale is None else rng.standard_normal(c.gate_angles.shape) * scale
    return c


@pytest.fixture(autouse=True)
def _defaults():
    qc.init()
    for k, v in (("force_simple", 0), ("tile_bits", 11), ("low_bits", 3), ("max_gates_per_pass", 1 << 30), ("adjoint_mma", 1),
                 ("adjoint_tile_bits", 0)):
        qc.set_option(k, v)
    yield
    qc.trim()


def _avail_host_bytes():
    try:
        import psutil

        return psutil.virtual_memory().available
    except Exception:
        return 0


# ---- C3: 30 qubits, ComplexF64 -- the first two half-layers (29 gates, including the gates on the top qubit pairs
# (29,30) and (28,29)) against the oracle, full vector.  src/circuits.jl:19-31 + src/apply.jl:62-84.
def test_c3_30q_prefix_vs_oracle(oracle):
    import torch

    n, layers = 30, 100
    free, _ = torch.cuda.mem_get_info()
    if free < 20 * 2 ** 30:
        pytest.skip("needs 20 GiB of free HBM")
    if _avail_host_bytes() < 36 * 2 ** 30:
        pytest.skip("the oracle needs 2 x 16 GiB of host memory at 30 qubits")
    rng = np.random.default_rng(3030)
    c = circuit(n, layers, rng)
    last = 29  # two half-layers: 15 + 14 gates
    pos = [j - 1 for l in (1, 2) for j in range(1 + (l - 1) % 2, n, 2)]
    assert len(pos) == last and (n - 2) in pos and (n - 3) in pos  # the gates on (29,30) and (28,29), 1-based
    st = qc.B200State(n, np.complex128)
    from qcb200.apply import _apply_circuit_inplace

    _apply_circuit_inplace(st, c, first=0, last=last)
    assert qc.get_stat("passes") < last  # fused passes, not one launch per gate
    assert abs(st.norm2() - 1) < 1e-12
    oracle.set_threads(0)
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[0] = 1
    oracle.apply_brickwork(ref, n, layers, c.gate_angles, 0, last, inplace=True)
    # compare chunk by chunk (1 GiB of host memory at a time)
    num = den = 0.0
    step = 1 << 26
    for a in range(0, 1 << n, step):
        got = st.tensor[a:a + step].cpu().numpy()
        d = got - ref[a:a + step]
        num += float(np.vdot(d, d).real)
        den += float(np.vdot(ref[a:a + step], ref[a:a + step]).real)
    assert abs(den - 1) < 1e-10
    assert np.sqrt(num / den) < 1e-12
    # and the expectation value of the same state, kernel vs oracle (src/hamiltonians/tfim.jl:17-54)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    E = qc.measure_(None, H, st)
    E_ref = oracle.measure_tfim(ref, 1.0, 0.1, 0.5).real
    assert abs(E - E_ref) < 1e-12 * max(1.0, abs(E_ref))


# ---- C3 size, a known answer that involves neither the oracle nor the device code: with the entangling angles at zero every
# gate is a product gate after a swap (tests/product_circuit.py), so circuit(|0..0>) is a product state whose amplitude at any
# index is a product of 30 numbers.  30 qubits x 12 half-layers (174 gates, fused passes incl. the top qubit pairs), 2^20
# random indices + the corners + the norm; ComplexF32 at 28 qubits.
@pytest.mark.parametrize("n,layers,dtype", [(30, 12, np.complex128), (28, 9, np.complex64), (17, 30, np.complex128)])
def test_non_entangling_circuit_gives_the_analytic_product_state(n, layers, dtype):
    import torch

    from product_circuit import product_amplitudes, product_circuit_vectors

    free, _ = torch.cuda.mem_get_info()
    if free < (20 if n == 30 else 6) * 2 ** 30:
        pytest.skip("not enough free HBM")
    rng = np.random.default_rng(77 + n)
    c = circuit(n, layers, rng)
    c.gate_angles[12:15] = 0
    st = qc.B200State(n, dtype)
    from qcb200.apply import _apply_circuit_inplace

    _apply_circuit_inplace(st, c)
    assert qc.get_stat("passes") < c.gate_angles.shape[1] / 3  # fused passes
    tol = 1e-12 if dtype == np.complex128 else 2e-5
    assert abs(st.norm2() - 1) < tol
    idx = np.concatenate([rng.integers(0, 1 << n, size=1 << 20), [0, 1, (1 << n) - 1, 1 << (n - 1), (1 << n) - 2]])
    want = product_amplitudes(product_circuit_vectors(n, layers, c.gate_angles), idx)
    got = st.tensor[torch.from_numpy(idx).to(st.tensor.device)].cpu().numpy()
    # amplitudes of a 30-qubit product state are ~ 2^-15: compare relative to the largest sampled magnitude
    assert np.max(np.abs(got - want)) < tol * np.max(np.abs(want))
    big = np.abs(want) > 0.01 * np.max(np.abs(want))
    assert np.count_nonzero(big) > 100
    assert np.max(np.abs(got[big] - want[big]) / np.abs(want[big])) < (1e-10 if dtype == np.complex128 else 2e-3)


# ---- C5: 28 qubits x 4 half-layers (54 gates, 810 angles), TFIM(J=1, h=0.9045, g=1.4), ComplexF64 and ComplexF32
# (examples/test_optimise.jl:9-16 shape; src/gradients.jl:1-17,88-152)
def test_c5_28q_gradient_and_optimise(oracle):
    import torch

    n, layers = 28, 4
    free, _ = torch.cuda.mem_get_info()
    if free < 24 * 2 ** 30:
        pytest.skip("needs 24 GiB of free HBM")
    if _avail_host_bytes() < 10 * 2 ** 30:
        pytest.skip("the oracle needs 2 x 4 GiB of host memory at 28 qubits")
    rng = np.random.default_rng(2828)
    c = circuit(n, layers, rng, 0.01)
    assert c.ngates == 54
    H = qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
    psi0 = qc.B200State(n, np.complex128)
    E, g = qc.gradients(H, psi0, c, calculate_energy=True)
    assert g.shape == (15, 54)
    # (i) energy against the oracle's forward pass + kernel measure
    oracle.set_threads(0)
    ref = onp.zero_state(n)
    oracle.apply_brickwork(ref, n, layers, c.gate_angles, inplace=True)
    E_ref = oracle.measure_tfim(ref, H.J, H.h, H.g).real
    assert abs(E - E_ref) < 1e-12 * max(1.0, abs(E_ref))
    # ... and the forward state itself (so that the shift rule below runs on an oracle-checked forward path)
    fwd = qc.apply(psi0, c)
    got = fwd.to_numpy()
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-12
    del got, ref, fwd
    # (ii) 12 random angles against the parameter-shift rule (src/gradients.jl:124-135) evaluated with that forward path
    flat = g.reshape(-1, order="F")
    scale_g = max(float(np.max(np.abs(g))), 1e-6)
    for idx in rng.choice(flat.size, 12, replace=False):
        ep = qc.measure(H, psi0, qc.reconstruct(c, int(idx), np.pi / 2))
        em = qc.measure(H, psi0, qc.reconstruct(c, int(idx), -np.pi / 2))
        assert abs((ep - em) / 2 - flat[idx]) < 1e-12 * max(1.0, scale_g) + 1e-11
    # (iii) ComplexF32 state: energy and gradient at the c64 tolerance against the ComplexF64 numbers
    psi32 = qc.B200State(n, np.complex64)
    E32, g32 = qc.gradients(H, psi32, c, calculate_energy=True)
    assert abs(E32 - E) < 1e-5 * max(1.0, abs(E))
    assert np.max(np.abs(g32 - g)) < 1e-5 * max(1.0, scale_g)
    # (iv) optimise!: 3 epochs, energies decrease, equal across dtypes to 1e-5, energies[1] is the energy checked above
    c64, c32 = qc.GenericBrickworkCircuit(n, layers, gate_angles=c.gate_angles.copy()), qc.GenericBrickworkCircuit(n, layers, gate_angles=c.gate_angles.copy())
    e64 = qc.optimise_(c64, H, psi0, 3, 0.01)
    e32 = qc.optimise_(c32, H, psi32, 3, 0.01)
    assert e64[0] == 0.0 and abs(e64[1] - E) < 1e-12 * abs(E)
    assert e64[2] < e64[1] and e64[3] < e64[2]
    assert np.max(np.abs(e64 - e32)) < 1e-5 * max(1.0, abs(E))
    assert np.max(np.abs(c64.gate_angles - c32.gate_angles)) < 1e-5
    # one epoch = theta - lr * grad with the gradient checked above (src/gradients.jl:13)
    c1 = qc.GenericBrickworkCircuit(n, layers, gate_angles=c.gate_angles.copy())
    qc.optimise_(c1, H, psi0, 1, 0.01)
    np.testing.assert_allclose(c1.gate_angles, c.gate_angles - 0.01 * g, atol=1e-15)


# ---- C1 shape: 12 qubits x 20 half-layers (G = 110, 1650 angles), ComplexF64, against the oracle's literal O(G^2)
# restatement of the reference's shift sweep (src/gradients.jl:88-152) -- TFIM and East, two angle distributions
@pytest.mark.parametrize("ham", ["tfim", "east"])
@pytest.mark.parametrize("scale", [None, 0.01])
def test_c1_12q_depth20_gradient_vs_reference_sweep(oracle, ham, scale):
    n, layers = 12, 20
    rng = np.random.default_rng(1220 + (0 if scale is None else 1))
    c = circuit(n, layers, rng, scale)
    assert c.ngates == 110
    H, kind = (qc.TFIMHamiltonian(1.0, 0.1, 0.5), 0) if ham == "tfim" else (qc.EastModelHamiltonian(0.1, -0.1), 1)
    oracle.set_threads(0)
    E_ref, g_ref = oracle.gradients_brickwork(onp.zero_state(n), n, layers, c.gate_angles, kind, H.params())
    E, g = qc.gradients(H, qc.zero_state_tensor(n), c, calculate_energy=True)
    scale_g = max(float(np.max(np.abs(g_ref))), 1e-6)
    assert abs(E - E_ref) < 1e-12 * max(1.0, abs(E_ref))
    assert np.max(np.abs(g - g_ref)) < 1e-12 * max(1.0, scale_g)
    # the forward circuit of the same config (BASELINE configs[0])
    out = qc.apply(qc.zero_state_tensor(n), c).reshape(-1, order="F")
    ref = oracle.apply_brickwork(onp.zero_state(n), n, layers, c.gate_angles)
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-12
    E32, g32 = qc.gradients(H, qc.zero_state_tensor(n).astype(np.complex64), c, calculate_energy=True)
    assert abs(E32 - E_ref) < 1e-5 * max(1.0, abs(E_ref))
    assert np.max(np.abs(g32 - g_ref)) < 1e-5 * max(1.0, scale_g)


# ---- C2 shape: 20 qubits x 4 half-layers, TFIM(1, 0.1, 0.5): shift rule on the ORACLE (not the GPU) for a subset of the
# angles, energy from the oracle (examples/test_gradients.jl:10-28)
def test_c2_20q_gradient_vs_oracle_shift(oracle):
    n, layers = 20, 4
    rng = np.random.default_rng(2004)
    c = circuit(n, layers, rng, 0.01)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    oracle.set_threads(0)
    E, g = qc.gradients(H, qc.B200State(qc.zero_state_tensor(n)), c, calculate_energy=True)
    ref = oracle.apply_brickwork(onp.zero_state(n), n, layers, c.gate_angles)
    E_ref = oracle.measure_tfim(ref, 1.0, 0.1, 0.5).real
    assert abs(E - E_ref) < 1e-12 * max(1.0, abs(E_ref))
    flat = g.reshape(-1, order="F")
    for idx in rng.choice(flat.size, 16, replace=False):
        es = []
        for sh in (np.pi / 2, -np.pi / 2):
            cc = qc.reconstruct(c, int(idx), sh)
            es.append(oracle.measure_tfim(oracle.apply_brickwork(onp.zero_state(n), n, layers, cc.gate_angles), 1.0, 0.1, 0.5).real)
        assert abs((es[0] - es[1]) / 2 - flat[idx]) < 1e-12


# ---- ComplexF32 fused backward pass (adjoint.cuh, FFMA2) against the ComplexF64 gradient, every tile size the backward
# pass supports, and against the oracle where it is fast enough
@pytest.mark.parametrize("n,layers,atb", [(13, 5, 0), (16, 6, 9), (18, 4, 10), (18, 7, 11), (20, 3, 12), (22, 4, 0)])
def test_c64_backward_pass_vs_c128(oracle, n, layers, atb):
    rng = np.random.default_rng(6400 + 10 * n + layers)
    c = circuit(n, layers, rng)
    H = qc.TFIMHamiltonian(1.0, 0.1, 0.5) if n % 2 else qc.EastModelHamiltonian(0.1, -0.1)
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    v /= np.linalg.norm(v)
    qc.set_option("adjoint_tile_bits", atb)
    E64, g64 = qc.gradients(H, qc.B200State(v), c, calculate_energy=True)
    E32, g32 = qc.gradients(H, qc.B200State(v.astype(np.complex64)), c, calculate_energy=True)
    scale_g = max(float(np.max(np.abs(g64))), 1e-6)
    assert abs(E32 - E64) < 1e-5 * max(1.0, abs(E64))
    assert np.max(np.abs(g32 - g64)) < 1e-5 * max(1.0, scale_g)
    if n <= 13:
        E_ref, g_ref = oracle.gradients_brickwork(v, n, layers, c.gate_angles, 0 if n % 2 else 1, H.params())
        assert np.max(np.abs(g32 - g_ref)) < 1e-5 * max(1.0, scale_g)
        assert np.max(np.abs(g64 - g_ref)) < 1e-12 * max(1.0, scale_g)


def test_optimise_matrix_hamiltonian(oracle):
    """optimise! is generic over H (src/gradients.jl:1-17): sparse-matrix TFIM == kernel TFIM trajectory."""
    n, layers, epochs, lr = 8, 3, 4, 0.01
    rng = np.random.default_rng(81)
    c = circuit(n, layers, rng, 0.01)
    a0 = c.gate_angles.copy()
    e_k = qc.optimise_(c, qc.TFIMHamiltonian(1.0, 0.2, 0.5), qc.zero_state_tensor(n), epochs, lr)
    c2 = qc.GenericBrickworkCircuit(n, layers, gate_angles=a0.copy())
    e_m = qc.optimise_(c2, qc.build_sparse_tfim_hamiltonian(n, 1.0, 0.2, 0.5), qc.zero_state_tensor(n), epochs, lr)
    np.testing.assert_allclose(e_m, e_k, atol=1e-11)
    np.testing.assert_allclose(c2.gate_angles, c.gate_angles, atol=1e-11)
    e_ref, a_ref = oracle.optimise_brickwork(onp.zero_state(n), n, layers, a0, 0, np.array([1.0, 0.2, 0.5]), epochs, lr)
    np.testing.assert_allclose(e_m, e_ref, atol=1e-10)
