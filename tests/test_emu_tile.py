"""CPU replay of the fused tile pass + scheduler (the same __host__ __device__ code the GPU runs)
against the oracle.  Needs nvcc (host compile only), no GPU."""
import shutil

import numpy as np
import pytest

from oracle import oracle_np as onp

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")


@pytest.fixture(scope="module")
def emu():
    import emu_helper

    emu_helper.build()
    return emu_helper


def rand_state(rng, n, dtype=np.complex128):
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    return (v / np.linalg.norm(v)).astype(dtype)


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("n,layers,tb,c", [(9, 6, 9, 3), (12, 20, 12, 3), (13, 9, 9, 3), (14, 12, 11, 3), (14, 7, 10, 4),
                                           (15, 10, 11, 3), (13, 8, 13, 3), (16, 5, 12, 4)])
def test_emu_brickwork_vs_oracle(emu, oracle, n, layers, tb, c):
    rng = np.random.default_rng(n * 100 + layers)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = rand_state(rng, n)
    ref = oracle.apply_brickwork(psi0, n, layers, angles)
    out, (npass, nst) = emu.apply_gates(psi0, q0, mats, tb, c)
    assert rel(out, ref) < 1e-13
    assert npass < G and nst < G  # fused: fewer sweeps than gates


@pytest.mark.parametrize("n,layers,tb,c", [(12, 7, 11, 3), (15, 9, 11, 3), (16, 6, 12, 4)])
def test_emu_non_entangling_circuit_gives_the_analytic_product_state(emu, n, layers, tb, c):
    """No oracle involved: the device code of the fused gate pass (tile.cuh) and the pass scheduler, replayed on the CPU with
    matrices from the library's host gate builder (qcb_general_unitary_gate), against the analytic product state of a
    brickwork circuit without entangling angles (tests/product_circuit.py)."""
    import qcb200
    from product_circuit import product_amplitudes, product_circuit_vectors

    rng = np.random.default_rng(n * 7 + layers)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    angles[12:15] = 0
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [qcb200.build_general_unitary_gate(angles[:, g]).mat().reshape(4, 4, order="F") for g in range(G)]
    out, (npass, _) = emu.apply_gates(onp.zero_state(n).reshape(-1), q0, mats, tb, c)
    want = product_amplitudes(product_circuit_vectors(n, layers, angles), np.arange(2 ** n))
    assert np.max(np.abs(out - want)) < 1e-14
    assert npass < G


def _product_case(n, layers, seed):
    import qcb200
    from product_circuit import product_amplitudes, product_circuit_vectors

    rng = np.random.default_rng(seed)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    angles[12:15] = 0
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [qcb200.build_general_unitary_gate(angles[:, g]).mat().reshape(4, 4, order="F") for g in range(G)]
    want = product_amplitudes(product_circuit_vectors(n, layers, angles), np.arange(2 ** n))
    return q0, mats, want


@pytest.mark.parametrize("n,g,layers,tb,fuse", [(13, 1, 14, 9, False), (14, 2, 21, 9, True), (15, 3, 18, 9, False), (16, 3, 17, 11, True)])
def test_emu_sharded_non_entangling_circuit_gives_the_analytic_product_state(emu, n, g, layers, tb, fuse):
    """No oracle involved: sharded planner (dist_plan.hpp: qubit swaps, bit-permuting passes, swaps fused into gate passes)
    with 2^g simulated ranks against the analytic product state (tests/product_circuit.py)."""
    q0, mats, want = _product_case(n, layers, 50 * n + g)
    out, phys, st = emu.sharded_apply_gates(onp.zero_state(n).reshape(-1), g, q0, mats, tb, 3, fuse_swaps=fuse)
    assert list(phys) == list(range(n))
    assert st["swaps"] >= 1
    assert np.max(np.abs(out - want)) < 1e-14


@pytest.mark.parametrize("n,layers,gbits,dtype", [(8, 9, -1, np.complex128), (11, 12, 2, np.complex128), (12, 20, 3, np.complex128), (12, 7, 3, np.complex64)])
def test_emu_cluster_kernel_non_entangling_circuit_gives_the_analytic_product_state(emu, n, layers, gbits, dtype):
    """No oracle involved: the cluster-resident small-circuit kernel (small_circuit.cuh, state in the distributed shared
    memory of 2^gbits CTAs) replayed on the CPU against the analytic product state."""
    q0, mats, want = _product_case(n, layers, 90 * n + layers)
    out = emu.small_circuit(onp.zero_state(n).reshape(-1).astype(dtype), q0, mats, gbits)
    assert np.max(np.abs(out - want)) < (1e-14 if dtype == np.complex128 else 2e-6)


def test_emu_3m_form_equals_ordinary_form(emu, oracle):
    """The 3M arithmetic of the deep ComplexF64 passes (tile.cuh: apply_window_gate_3m, 48 products + 4 sums per amplitude
    quad) against the ordinary 64-product form and the oracle (src/apply.jl:62-84): same result up to rounding, and the
    error of either form stays at the level of the other (normwise bound) over a deep circuit."""
    rng = np.random.default_rng(3)
    n, layers = 13, 24
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = rand_state(rng, n)
    ref = oracle.apply_brickwork(psi0, n, layers, angles)
    try:
        emu.set_m3(False)
        out0, _ = emu.apply_gates(psi0, q0, mats, 11, 3)
        emu.set_m3(True)
        out3, _ = emu.apply_gates(psi0, q0, mats, 11, 3)
    finally:
        emu.set_m3(True)
    assert not np.array_equal(out0, out3)  # two different instruction sequences ...
    assert rel(out0, ref) < 1e-13 and rel(out3, ref) < 1e-13  # ... with the same accuracy
    assert rel(out3, out0) < 1e-13
    assert abs(np.linalg.norm(out3) - 1) < 1e-13


@pytest.mark.parametrize("n,layers,tb,grid,dt", [(13, 9, 9, 3, np.complex128), (14, 8, 11, 5, np.complex128), (15, 6, 11, 1, np.complex64), (12, 10, 11, 7, np.complex128)])
def test_emu_async_gate_pass(emu, oracle, n, layers, tb, grid, dt):
    """Persistent gate pass with asynchronous tile loads (tile_async.cuh): the same numbers as the one-tile-per-CTA form,
    bit for bit, on any grid size (a grid that does not divide the tile count, more CTAs than tiles)."""
    rng = np.random.default_rng(7700 + n)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = rand_state(rng, n, dt)
    ref, _ = emu.apply_gates(psi0, q0, mats, tb, 3)
    try:
        emu.set_async(grid)
        out, _ = emu.apply_gates(psi0, q0, mats, tb, 3)
    finally:
        emu.set_async(0)
    assert np.array_equal(out, ref)
    assert rel(out.astype(np.complex128), oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, angles)) < (1e-13 if dt == np.complex128 else 1e-5)


@pytest.mark.parametrize("n,layers,gbits,dt", [(4, 6, 0, np.complex128), (9, 8, 0, np.complex128), (10, 7, 1, np.complex128), (11, 6, 2, np.complex64),
                                               (12, 20, 3, np.complex128), (12, 5, -1, np.complex64), (8, 9, 3, np.complex128), (6, 5, 2, np.complex128)])
def test_emu_small_circuit(emu, oracle, n, layers, gbits, dt):
    """Cluster-resident small-circuit kernel (small_circuit.cuh), replayed with simulated CTAs: local gates in place, gates on the
    CTA-rank bits owner-computes through the other CTAs' buffers; vs the oracle (src/apply.jl:62-84), and U^dagger undoes U."""
    rng = np.random.default_rng(5100 + 13 * n + layers)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = rand_state(rng, n, dt)
    tol = 1e-13 if dt == np.complex128 else 1e-5
    out = emu.small_circuit(psi0, q0, mats, gbits)
    assert rel(out.astype(np.complex128), oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, angles)) < tol
    back = emu.small_circuit(out, q0[::-1], mats[::-1], gbits, adjoint=True)
    assert rel(back.astype(np.complex128), psi0.astype(np.complex128)) < 10 * tol


def test_emu_c64(emu, oracle):
    rng = np.random.default_rng(64)
    n, layers = 13, 10
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = rand_state(rng, n, np.complex64)
    ref = oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, angles)
    out, _ = emu.apply_gates(psi0, q0, mats, 11, 4)
    assert out.dtype == np.complex64
    assert rel(out.astype(np.complex128), ref) < 1e-5


def test_emu_random_nonunitary_gate_list(emu, oracle):
    # test/correctness_tests.jl:86-124 style matrices, random (non-brickwork) positions
    rng = np.random.default_rng(7)
    n, G = 13, 60
    q0 = rng.integers(0, n - 1, size=G)
    mats = []
    for _ in range(G):
        u = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        mats.append(u / np.linalg.det(u) ** 0.25)
    psi0 = rand_state(rng, n)
    ref = psi0
    for q, M in zip(q0, mats):
        ref = oracle.apply_2q(ref, M, int(q) + 1)
    for tb, c in [(9, 3), (11, 3), (13, 4)]:
        out, _ = emu.apply_gates(psi0, q0, mats, tb, c)
        assert rel(out, ref) < 1e-12


def test_emu_adjoint_undoes_circuit(emu, oracle):
    rng = np.random.default_rng(5)
    n, layers = 12, 6
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = rand_state(rng, n)
    fwd, _ = emu.apply_gates(psi0, q0, mats, 10, 3)
    back, _ = emu.apply_gates(fwd, q0[::-1], mats[::-1], 10, 3, adjoint=True)
    assert rel(back, psi0) < 1e-13


def test_emu_permuted_layout(emu, oracle):
    """logical -> physical bit permutation (what the sharded path uses after a qubit swap)."""
    rng = np.random.default_rng(11)
    n, layers = 12, 5
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = rand_state(rng, n)
    ref = oracle.apply_brickwork(psi0, n, layers, angles)
    for trial in range(3):
        phys = rng.permutation(n)
        # physical-layout copy of the logical state: bit phys[q] of the physical index = bit q of the logical index
        idx = np.arange(2 ** n)
        pidx = np.zeros_like(idx)
        for q in range(n):
            pidx |= ((idx >> q) & 1) << phys[q]
        psi_phys = np.empty_like(psi0)
        psi_phys[pidx] = psi0
        out_phys, _ = emu.apply_gates(psi_phys, q0, mats, 9, 3, phys=phys)
        assert rel(out_phys[pidx], ref) < 1e-13


def test_schedule_statistics(emu):
    # 30 qubits x 100 half-layers (BASELINE config C3): the pass count bounds HBM traffic
    st = emu.plan_brickwork(30, 100, 11, 3)
    assert st["gates"] == 1450
    assert st["passes"] <= 100 and st["stages"] <= 520
    st4 = emu.plan_brickwork(30, 100, 11, 3, max_gates=4)
    assert st4["gates"] == 1450 and st4["passes"] >= 1450 // 4


# ---- sharded path with simulated ranks: scheduler, bit-permuting passes, swap bookkeeping ----------------
@pytest.mark.parametrize("n,g,layers,tb,c", [(13, 1, 12, 9, 3), (14, 2, 20, 9, 3), (15, 3, 30, 9, 3), (14, 2, 9, 10, 4), (16, 3, 25, 11, 3)])
def test_emu_sharded_brickwork(emu, oracle, n, g, layers, tb, c):
    rng = np.random.default_rng(n * 1000 + g * 100 + layers)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, i]) for i in range(G)]
    psi0 = rand_state(rng, n)
    ref = oracle.apply_brickwork(psi0, n, layers, angles)
    out, phys, st = emu.sharded_apply_gates(psi0, g, q0, mats, tb, c)
    assert list(phys) == list(range(n))
    assert rel(out, ref) < 1e-13
    assert st["swaps"] >= 1  # the top qubits had to be exchanged at least once
    # without canonicalisation the result is the same state in a permuted physical layout
    out2, phys2, _ = emu.sharded_apply_gates(psi0, g, q0, mats, tb, c, canonical_out=False)
    idx = np.arange(2 ** n)
    pidx = np.zeros_like(idx)
    for q in range(n):
        pidx |= ((idx >> q) & 1) << int(phys2[q])
    assert rel(out2[pidx], ref) < 1e-13


def test_emu_sharded_random_gate_list(emu, oracle):
    rng = np.random.default_rng(99)
    n, g, G = 14, 2, 80
    q0 = rng.integers(0, n - 1, size=G)
    mats = []
    for _ in range(G):
        u = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        mats.append(u / np.linalg.det(u) ** 0.25)
    psi0 = rand_state(rng, n)
    ref = psi0
    for q, M in zip(q0, mats):
        ref = oracle.apply_2q(ref, M, int(q) + 1)
    out, phys, st = emu.sharded_apply_gates(psi0, g, q0, mats, 9, 3)
    assert rel(out, ref) < 1e-12


def test_sharded_schedule_34q(emu):
    # BASELINE config C4: 34 qubits x 100 half-layers on 8 ranks: a handful of all-to-all swaps
    st = emu.plan_sharded_brickwork(34, 100, 3, 11, 3)
    assert st["swaps"] <= 6 and st["passes"] <= 110


# ---- fused backward pass of the adjoint gradient sweep ----------------------------------------------------
@pytest.mark.parametrize("n,layers,tb,dtype", [(9, 4, 9, np.complex128), (12, 6, 10, np.complex128), (13, 5, 11, np.complex128), (12, 4, 11, np.complex64)])
def test_emu_adjoint_sweep_gradients(emu, oracle, n, layers, tb, dtype):
    """env from the fused backward pass, contracted with the analytic dU/dtheta, equals the reference's
    parameter-shift gradients (oracle restatement of src/gradients.jl:88-152)."""
    import ctypes as C

    from qcb200 import _lib

    rng = np.random.default_rng(n * 7 + layers)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psi0 = onp.zero_state(n)
    hp = (1.0, 0.1, 0.5)
    psiG = oracle.apply_brickwork(psi0, n, layers, angles)
    H = onp.dense_tfim(n, *hp) if n <= 10 else None
    if H is not None:
        lamG = H @ psiG
    else:  # matrix-free H psi for the TFIM
        C_ = np.arange(2 ** n)
        zz = sum(np.where(((C_ >> i) & 3 == 0) | ((C_ >> i) & 3 == 3), 1.0, -1.0) for i in range(n - 1))
        z = n - 2 * sum((C_ >> i) & 1 for i in range(n))
        lamG = -hp[0] * ((zz + hp[1] * z) * psiG + hp[2] * sum(psiG[C_ ^ (1 << i)] for i in range(n)))
    back_psi, back_lam, env = emu.adjoint_sweep(psiG.astype(dtype), lamG.astype(dtype), q0, mats, tb, 3)
    tol = 1e-12 if dtype == np.complex128 else 2e-5
    assert rel(back_psi.astype(np.complex128), psi0) < tol  # every gate was un-applied
    grads = np.zeros((15, G))
    for g in range(G):
        a = np.ascontiguousarray(angles[:, g])
        for k in range(15):
            d = np.empty(16, dtype=np.complex128)
            _lib.check(_lib.lib().qcb_general_unitary_gate_derivative(a.ctypes.data_as(C.c_void_p), C.c_int(k), d.ctypes.data_as(C.c_void_p)))
            grads[k, g] = 2 * np.real(np.sum(env[g] * d.reshape(4, 4, order="F")))
    _, g_ref = oracle.gradients_brickwork(psi0, n, layers, angles, 0, hp)
    assert np.max(np.abs(grads - g_ref)) < tol * max(1.0, np.max(np.abs(g_ref)))


@pytest.mark.parametrize("n,layers,tb,nw", [(9, 4, 9, 4), (12, 6, 10, 4), (12, 5, 10, 8), (13, 5, 11, 8), (12, 3, 12, 16), (10, 4, 9, 2)])
def test_emu_adjoint_sweep_mma(emu, oracle, n, layers, tb, nw):
    """The tensor-core backward pass (adjoint_mma.cuh, software model of the m8n8k4 fragments) gives the same un-applied
    states and -- after env = conj(U) M on the host -- the same environments as the scalar pass, and its stage
    loads/stores are shared-memory bank-conflict free."""
    rng = np.random.default_rng(n * 11 + layers + nw)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psiG, lamG = rand_state(rng, n), rand_state(rng, n)
    p_ref, l_ref, env_ref = emu.adjoint_sweep(psiG, lamG, q0, mats, tb, 3)
    p, l, env, (wf, wf_min) = emu.adjoint_sweep_mma(psiG, lamG, q0, mats, tb, 3, nw)
    assert rel(p, p_ref) < 1e-13 and rel(l, l_ref) < 1e-13
    assert np.max(np.abs(env - env_ref)) < 1e-13
    # independent definition of the environment: env[g][r, c] = sum conj(lam_g[r]) psi_{g-1}[c]
    psi, lam = psiG.copy(), lamG.copy()
    for g in range(G - 1, -1, -1):
        Ud = np.conj(mats[g].reshape(4, 4)).T
        psi = oracle.apply_2q(psi, Ud, int(q0[g]) + 1)
        k = int(q0[g])
        L4 = np.moveaxis(lam.reshape(2 ** (n - k - 2), 4, 2 ** k), 1, 0).reshape(4, -1)
        P4 = np.moveaxis(psi.reshape(2 ** (n - k - 2), 4, 2 ** k), 1, 0).reshape(4, -1)
        assert np.max(np.abs(env[g] - np.conj(L4) @ P4.T)) < 1e-13
        lam = oracle.apply_2q(lam, Ud, int(q0[g]) + 1)
    assert wf == wf_min, f"{wf - wf_min} bank-conflict wavefronts out of {wf_min}"


@pytest.mark.parametrize("n,layers,tb", [(9, 4, 9), (11, 5, 10), (12, 6, 11), (13, 4, 11), (12, 3, 12), (14, 5, 10)])
def test_emu_adjoint_sweep_f32(emu, oracle, n, layers, tb):
    """ComplexF32 backward pass with both windows in registers, per-thread outer products and the warp transpose-reduction
    (adjoint_f32.cuh, all 32 lanes replayed): un-applied states bit-identical to the quad-per-thread pass (same FMA order),
    environments equal to the ComplexF64 definition at single precision."""
    rng = np.random.default_rng(n * 17 + layers + tb)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, g]) for g in range(G)]
    psiG, lamG = rand_state(rng, n), 3.0 * rand_state(rng, n)
    p_ref, l_ref, env_ref = emu.adjoint_sweep(psiG.astype(np.complex64), lamG.astype(np.complex64), q0, mats, tb, 3)
    p, l, env = emu.adjoint_sweep_f32(psiG, lamG, q0, mats, tb, 3)
    assert np.array_equal(p, p_ref) and np.array_equal(l, l_ref)
    # independent ComplexF64 definition: env[g][r, c] = sum conj(lam_g[r]) psi_{g-1}[c]
    psi, lam = psiG.copy(), lamG.copy()
    scale = np.linalg.norm(psiG) * np.linalg.norm(lamG)
    worst = 0.0
    for g in range(G - 1, -1, -1):
        Ud = np.conj(mats[g].reshape(4, 4)).T
        psi = oracle.apply_2q(psi, Ud, int(q0[g]) + 1)
        k = int(q0[g])
        L4 = np.moveaxis(lam.reshape(2 ** (n - k - 2), 4, 2 ** k), 1, 0).reshape(4, -1)
        P4 = np.moveaxis(psi.reshape(2 ** (n - k - 2), 4, 2 ** k), 1, 0).reshape(4, -1)
        worst = max(worst, np.max(np.abs(env[g] - np.conj(L4) @ P4.T)))
        lam = oracle.apply_2q(lam, Ud, int(q0[g]) + 1)
    assert worst < 2e-6 * scale, worst / scale
    # FP32 accumulation over one warp and tile only: within rounding of the pass that accumulates every product in double
    assert np.max(np.abs(env - env_ref)) < 2e-6 * scale


@pytest.mark.parametrize("n,layers,g,tb,ham", [(12, 5, 1, 9, "tfim"), (13, 6, 2, 9, "east"), (13, 4, 3, 10, "tfim"), (12, 6, 2, 10, "east")])
def test_emu_sharded_gradients(emu, oracle, n, layers, g, tb, ham):
    """Sharded adjoint gradients on 2^g simulated ranks: lambda = H psi in two sweeps around one qubit swap, backward sweep on
    the (psi, lambda) pair through the sharded planner -> the reference's parameter-shift gradients and energy."""
    import ctypes as C

    from qcb200 import _lib

    rng = np.random.default_rng(n * 13 + layers + g)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, k]) for k in range(G)]
    psi0 = rand_state(rng, n)
    if ham == "tfim":
        kind, hp = 0, (1.0, 0.1, 0.5)
    else:
        A, B = onp.east_params(0.1, -0.1)
        kind, hp = 1, (0.1, A, B)
    env, E, st = emu.sharded_gradient_env(psi0, g, kind, hp, q0, mats, tb, 3)
    assert not st["canonical_after_forward"]  # lambda = H psi is built in whatever layout the forward circuit ends in
    E_ref, g_ref = oracle.gradients_brickwork(psi0, n, layers, angles, kind, hp)
    assert abs(E.real - E_ref) < 1e-12 * max(1.0, abs(E_ref)) and abs(E.imag) < 1e-12
    grads = np.zeros((15, G))
    for k in range(G):
        a = np.ascontiguousarray(angles[:, k])
        for j in range(15):
            d = np.empty(16, dtype=np.complex128)
            _lib.check(_lib.lib().qcb_general_unitary_gate_derivative(a.ctypes.data_as(C.c_void_p), C.c_int(j), d.ctypes.data_as(C.c_void_p)))
            grads[j, k] = 2 * np.real(np.sum(env[k] * d.reshape(4, 4, order="F")))
    assert np.max(np.abs(grads - g_ref)) < 1e-12 * max(1.0, np.max(np.abs(g_ref)))
    assert st["swaps"] >= 1  # the top qubits' gates needed at least one more exchange after the Hamiltonian's


@pytest.mark.parametrize("n,layers,g,ham", [(12, 3, 1, "tfim"), (13, 2, 2, "east")])
def test_emu_gradients_against_central_differences(emu, n, layers, g, ham):
    """No oracle involved: the adjoint sweep replayed on the CPU (sharded planner on 2^g simulated ranks, environments
    contracted with the library's dU/dtheta) against central differences of the replayed forward circuit + expectation value,
    gate matrices from the library's host builder (src/gradients.jl:88-152 computes the same numbers by the shift rule)."""
    import ctypes as C

    import qcb200
    from qcb200 import _lib

    rng = np.random.default_rng(n * 17 + layers)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    build = lambda a: [qcb200.build_general_unitary_gate(a[:, k]).mat().reshape(4, 4, order="F") for k in range(G)]
    psi0 = rand_state(rng, n)
    kind, hp = (0, (1.0, 0.9045, 1.4)) if ham == "tfim" else (1, (0.1,) + tuple(onp.east_params(0.1, -0.1)))
    env, E, _ = emu.sharded_gradient_env(psi0, g, kind, hp, q0, build(angles), 9, 3)

    def energy(a):
        out, _ = emu.apply_gates(psi0, q0, build(a), 10, 3)
        return emu.measure(out, kind, hp, 9, 3)[0]

    assert abs(E.real - energy(angles)) < 1e-12
    eps = 1e-5
    for j, k in [(0, 0), (7, G // 2), (14, G - 1), (12, 1), (13, G - 2), (5, 2)]:
        a = np.ascontiguousarray(angles[:, k])
        d = np.empty(16, dtype=np.complex128)
        _lib.check(_lib.lib().qcb_general_unitary_gate_derivative(a.ctypes.data_as(C.c_void_p), C.c_int(j), d.ctypes.data_as(C.c_void_p)))
        grad = 2 * np.real(np.sum(env[k] * d.reshape(4, 4, order="F")))
        ap, am = angles.copy(), angles.copy()
        ap[j, k] += eps
        am[j, k] -= eps
        assert abs((energy(ap) - energy(am)) / (2 * eps) - grad) < 2e-8, (j, k)


# ---- tiled expectation value -----------------------------------------------------------------------------
@pytest.mark.parametrize("n,tb,c", [(10, 8, 3), (13, 9, 3), (14, 12, 3), (15, 13, 3), (16, 10, 4), (17, 12, 3)])
def test_emu_tiled_measure(emu, oracle, n, tb, c):
    rng = np.random.default_rng(n * 31 + tb)
    psi = rand_state(rng, n)
    A, B = onp.east_params(0.1, -0.1)
    for kind, hp, ref in [(0, (1.0, 0.1, 0.5), oracle.measure_tfim(psi, 1.0, 0.1, 0.5).real),
                          (0, (0.7, -0.3, 1.4), oracle.measure_tfim(psi, 0.7, -0.3, 1.4).real),
                          (1, (0.1, A, B), oracle.measure_east(psi, 0.1, A, B).real)]:
        E, npass = emu.measure(psi, kind, hp, tb, c)
        assert abs(E - ref) < 1e-12 * max(1.0, abs(ref))
        assert npass == 1 + max(0, -(-(n - tb) // (tb - c)))
    E32, _ = emu.measure(psi.astype(np.complex64), 0, (1.0, 0.1, 0.5), tb, c)
    assert abs(E32 - oracle.measure_tfim(psi, 1.0, 0.1, 0.5).real) < 1e-5 * n


# ---- lambda = H psi with the low-bit flips from a staged tile -----------------------------------------------
def _ham_apply_np(psi, kind, hp):
    """H psi by index arithmetic: TFIM -J (sum ZZ + h sum Z + g sum X) (src/hamiltonians/tfim.jl:17-54), East model
    (east_model.jl:30-52); qubit i + 1 is index bit i.  Checked against the oracle's dense matrices in the test below."""
    v = np.asarray(psi, dtype=np.complex128).reshape(-1)
    n = int(v.size).bit_length() - 1
    idx = np.arange(v.size, dtype=np.int64)
    bits = [(idx >> i) & 1 for i in range(n)]
    if kind == 0:
        J, h, g = hp
        z = [1 - 2 * b for b in bits]
        diag = -J * (sum(z[i] * z[i + 1] for i in range(n - 1)) + h * sum(z))
        return diag * v - J * g * sum(v[idx ^ (1 << i)] for i in range(n))
    c, A, B = hp
    nn = [1 - b for b in bits]  # the projector N = |0><0|
    diag = B * (nn[0] + sum(nn[i - 1] * nn[i] + c * nn[i - 1] for i in range(1, n))) + c
    x = v[idx ^ 1] + sum(nn[i - 1] * v[idx ^ (1 << i)] for i in range(1, n))
    return diag * v + A * x


def test_ham_apply_np_matches_dense_hamiltonians():
    rng = np.random.default_rng(77)
    n = 8
    psi = rand_state(rng, n)
    A, B = onp.east_params(0.1, -0.1)
    assert np.max(np.abs(_ham_apply_np(psi, 0, (0.7, -0.3, 1.4)) - onp.dense_tfim(n, 0.7, -0.3, 1.4) @ psi)) < 1e-13
    assert np.max(np.abs(_ham_apply_np(psi, 1, (0.1, A, B)) - onp.dense_east(n, 0.1, -0.1) @ psi)) < 1e-13


@pytest.mark.parametrize("n,grid", [(12, 1), (13, 3), (15, 5)])
def test_emu_apply_ham_tiled(emu, oracle, n, grid):
    """k_apply_ham_tiled replayed thread by thread (tile load, LDS flips of the low bits, gathers of the rest): lambda against
    the index-arithmetic H psi, E = <psi|lambda> against the oracle's expectation kernels; both dtypes, TFIM and East."""
    rng = np.random.default_rng(500 + n)
    psi = rand_state(rng, n)
    A, B = onp.east_params(0.1, -0.1)
    for kind, hp, ref in [(0, (0.7, -0.3, 1.4), oracle.measure_tfim(psi, 0.7, -0.3, 1.4)),
                          (1, (0.1, A, B), oracle.measure_east(psi, 0.1, A, B))]:
        want = _ham_apply_np(psi, kind, hp)
        lam, E = emu.apply_ham_tiled(psi, kind, hp, grid)
        assert np.max(np.abs(lam - want)) < 1e-14 * n
        assert abs(E - ref) < 1e-12 * max(1.0, abs(ref))
        p32 = psi.astype(np.complex64)
        lam32, E32 = emu.apply_ham_tiled(p32, kind, hp, grid)
        assert lam32.dtype == np.complex64
        assert np.max(np.abs(lam32 - _ham_apply_np(p32, kind, hp))) < 2e-7 * np.max(np.abs(want))  # one FP32 rounding
        assert abs(E32 - ref) < 1e-5 * max(1.0, abs(ref))


# ---- scheduler / tile-shape corner cases ------------------------------------------------------------------
@pytest.mark.parametrize("n,tb,c,maxg", [(9, 13, 3, 0), (10, 10, 0, 0), (11, 9, 5, 0), (12, 12, 3, 1), (12, 9, 3, 2), (13, 11, 4, 5), (10, 9, 3, 3)])
def test_emu_tile_shape_corner_cases(emu, oracle, n, tb, c, maxg):
    """tile larger than the state (clamped), no coalescing bits, more low bits than usual, and fusion caps of 1-5 gates."""
    rng = np.random.default_rng(n * 17 + tb * 3 + c + maxg)
    G = 45
    q0 = rng.integers(0, n - 1, size=G)
    mats = []
    for _ in range(G):
        u = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        mats.append(u / np.linalg.det(u) ** 0.25)
    psi0 = rand_state(rng, n)
    ref = psi0
    for q, M in zip(q0, mats):
        ref = oracle.apply_2q(ref, M, int(q) + 1)
    out, (npass, nst) = emu.apply_gates(psi0, q0, mats, tb, c, max_gates=maxg)
    assert rel(out, ref) < 1e-12
    if maxg:
        assert npass >= -(-G // maxg)


def test_emu_repeated_gate_on_same_pair(emu, oracle):
    """many consecutive gates on one qubit pair (no brickwork structure at all) still schedule and stay exact."""
    rng = np.random.default_rng(5)
    n, G = 11, 30
    q0 = np.full(G, 4)
    mats = [oracle.build_general_unitary_gate(rng.uniform(0, 2 * np.pi, 15)) for _ in range(G)]
    psi0 = rand_state(rng, n)
    ref = psi0
    for M in mats:
        ref = oracle.apply_2q(ref, M, 5)
    out, (npass, nst) = emu.apply_gates(psi0, q0, mats, 9, 3)
    assert rel(out, ref) < 1e-12
    assert npass <= 4  # up to 12 stages x 1-2 gates per pass


@pytest.mark.parametrize("n,g,layers,tb,c", [(13, 1, 14, 9, 3), (14, 2, 20, 9, 3), (15, 3, 30, 9, 3), (16, 3, 25, 11, 3)])
def test_emu_sharded_fused_swaps(emu, oracle, n, g, layers, tb, c):
    """qubit swaps fused into the following pass: every rank loads its tiles from the peers' old shards (the address
    algebra of tile_load_peer / PeerSrc) instead of performing the all-to-all first."""
    rng = np.random.default_rng(n * 1000 + g * 100 + layers + 7)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.array(onp.brickwork_gate_positions(n, layers)) - 1
    mats = [oracle.build_general_unitary_gate(angles[:, i]) for i in range(G)]
    psi0 = rand_state(rng, n)
    ref = oracle.apply_brickwork(psi0, n, layers, angles)
    out, phys, st = emu.sharded_apply_gates(psi0, g, q0, mats, tb, c, fuse_swaps=True)
    assert list(phys) == list(range(n))
    assert rel(out, ref) < 1e-13
    assert st["fused"] >= 1 and st["fused"] == st["swaps"]  # every exchange rode on a pass


@pytest.mark.parametrize("n,g,tb,c", [(13, 1, 9, 3), (14, 2, 9, 3), (15, 3, 10, 3), (16, 3, 12, 3)])
def test_emu_sharded_measure(emu, oracle, n, g, tb, c):
    """sharded expectation value: rank bits enter the diagonal / East controls through c_hi, the flips of the qubits in the
    rank bits are taken after one qubit swap."""
    rng = np.random.default_rng(n * 5 + g)
    psi = rand_state(rng, n)
    A, B = onp.east_params(0.1, -0.1)
    for kind, hp, ref in [(0, (1.0, 0.1, 0.5), oracle.measure_tfim(psi, 1.0, 0.1, 0.5).real),
                          (1, (0.1, A, B), oracle.measure_east(psi, 0.1, A, B).real)]:
        E = emu.sharded_measure(psi, g, kind, hp, tb, c)
        assert abs(E - ref) < 1e-12 * max(1.0, abs(ref))


def test_device_gate_algebra_matches_host(emu, oracle):
    """gates_hd.cuh (runs on the GPU in k_contract_env; compiled for the host here) == gates_host.hpp through the ABI: the
    gate itself, the 15-angle contraction, and the env = conj(U) M step of the tensor-core backward pass."""
    import ctypes as C

    from qcb200 import _lib

    rng = np.random.default_rng(2718)
    L = _lib.lib()
    for _ in range(10):
        a = np.ascontiguousarray(rng.uniform(-2 * np.pi, 2 * np.pi, 15))
        env = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        out, U = emu.contract_env_gate(a, env, False)
        U_ref = np.asarray(oracle.build_general_unitary_gate(a)).reshape(4, 4)
        assert np.max(np.abs(U - U_ref)) < 1e-14
        ref = np.zeros(15)
        e = np.ascontiguousarray(env.reshape(-1, order="F"))
        _lib.check(L.qcb_general_unitary_gate_gradient(a.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p)))
        assert np.max(np.abs(out - ref)) < 1e-13 * max(1.0, np.max(np.abs(ref)))
        out_post, _ = emu.contract_env_gate(a, env, True)
        e2 = np.ascontiguousarray((np.conj(U_ref) @ env).reshape(-1, order="F"))
        _lib.check(L.qcb_general_unitary_gate_gradient(a.ctypes.data_as(C.c_void_p), e2.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p)))
        assert np.max(np.abs(out_post - ref)) < 1e-13 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("n,tb,c,grid", [(13, 9, 3, 5), (14, 12, 3, 3), (15, 10, 4, 7), (16, 12, 3, 1)])
def test_emu_tiled_measure_persistent(emu, oracle, n, tb, c, grid):
    """Persistent form of the tiled expectation value (tile-independent thread constants hoisted out of the tile loop; option
    measure_persistent, off by default) == the one-tile-per-CTA form == the oracle."""
    rng = np.random.default_rng(n * 37 + tb)
    A, B = onp.east_params(0.1, -0.1)
    for dt in (np.complex128, np.complex64):
        psi = rand_state(rng, n).astype(dt)
        for kind, hp, ref in [(0, (1.0, 0.1, 0.5), oracle.measure_tfim(psi.astype(np.complex128), 1.0, 0.1, 0.5).real),
                              (1, (0.1, A, B), oracle.measure_east(psi.astype(np.complex128), 0.1, A, B).real)]:
            E = emu.measure_persistent(psi, kind, hp, tb, c, grid)
            E_tile, _ = emu.measure(psi, kind, hp, tb, c)
            tol = 1e-12 if dt == np.complex128 else 1e-5
            assert abs(E - ref) < tol * max(1.0, abs(ref))
            assert abs(E - E_tile) < 1e-12 * max(1.0, abs(ref))


def test_emu_scheduler_property_random_gate_lists(emu, oracle):
    """Property test of the pass scheduler + tile kernels on the CPU replay: ANY program-ordered list of adjacent 2-qubit
    gates, any tile shape, with or without a cap on the gates per pass, gives the state the gate-by-gate oracle gives."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
    @given(n=st.integers(9, 12), tb=st.integers(9, 12), low=st.integers(0, 4), cap=st.sampled_from([0, 1, 3, 7]),
           seed=st.integers(0, 2 ** 31 - 1), ngates=st.integers(1, 30))
    def run(n, tb, low, cap, seed, ngates):
        tb = min(tb, n)
        rng = np.random.default_rng(seed)
        q0 = rng.integers(0, n - 1, size=ngates)
        mats = []
        for _ in range(ngates):
            u = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
            mats.append(u / np.linalg.norm(u) * 2.0)
        psi0 = rand_state(rng, n)
        ref = psi0
        for q, M in zip(q0, mats):
            ref = oracle.apply_2q(ref, M, int(q) + 1)
        out, (npass, _) = emu.apply_gates(psi0, q0, mats, tb, min(low, tb - 4), max_gates=cap)
        assert rel(out, ref) < 1e-12
        if cap:
            assert npass >= -(-ngates // cap)

    run()


def test_emu_adjoint_mma_property_random_gate_lists(emu, oracle):
    """Property test of the tensor-core backward pass replay: any list of adjacent unitary 2-qubit gates (not only brickwork
    layers), any supported tile / warp shape: same un-applied states and environments as the scalar pass, and no
    shared-memory bank conflict in any stage (the choice of `sel` in fill_adj_mma must work for every window position)."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    shapes = [(9, 4), (9, 2), (10, 4), (10, 8), (11, 8)]

    @settings(max_examples=20, deadline=None, suppress_health_check=list(HealthCheck))
    @given(n=st.integers(11, 12), shape=st.sampled_from(shapes), seed=st.integers(0, 2 ** 31 - 1), ngates=st.integers(1, 24))
    def run(n, shape, seed, ngates):
        tb, nw = shape
        rng = np.random.default_rng(seed)
        q0 = rng.integers(0, n - 1, size=ngates)
        mats = [oracle.build_general_unitary_gate(rng.uniform(0, 2 * np.pi, 15)) for _ in range(ngates)]
        psi, lam = rand_state(rng, n), rand_state(rng, n)
        p_ref, l_ref, env_ref = emu.adjoint_sweep(psi, lam, q0, mats, tb, 3)
        p, l, env, (wf, wf_min) = emu.adjoint_sweep_mma(psi, lam, q0, mats, tb, 3, nw)
        assert rel(p, p_ref) < 1e-12 and rel(l, l_ref) < 1e-12
        assert np.max(np.abs(env - env_ref)) < 1e-12
        assert wf == wf_min

    run()


# ---- TMA-fed persistent expectation value (flip_tile.cuh): consumer code replayed on emulated tile loads --------------------
@pytest.mark.parametrize("n,dt", [(14, np.complex128), (15, np.complex128), (17, np.complex128), (21, np.complex128), (22, np.complex128),
                                  (15, np.complex64), (16, np.complex64), (19, np.complex64), (22, np.complex64), (23, np.complex64)])
def test_emu_flip_measure(emu, oracle, n, dt):
    rng = np.random.default_rng(700 + n)
    psi = rand_state(rng, n).astype(dt)
    ref_psi = psi.astype(np.complex128)
    A, B = -np.exp(0.1) * np.sqrt(0.1 * 0.9), 0.8
    tol = 1e-12 if dt == np.complex128 else 1e-5
    for kind, hp, ref in [(0, (1.0, 0.1, 0.5), oracle.measure_tfim(ref_psi, 1.0, 0.1, 0.5).real),
                          (0, (0.7, -0.3, 1.4), oracle.measure_tfim(ref_psi, 0.7, -0.3, 1.4).real),
                          (1, (0.1, A, B), oracle.measure_east(ref_psi, 0.1, A, B).real)]:
        d, f, nl = emu.flip_measure(psi, kind, hp, grid=5)
        coef = -hp[0] * hp[2] if kind == 0 else hp[1]
        E = d + coef * f
        assert abs(E - ref) < tol * max(1.0, abs(ref)), (kind, E, ref)
        tba = 12 if dt == np.complex128 else 13
        assert nl == 1 + -(-(n - tba) // 9)


def _tfim_ground_state(n, J, g):
    """(E0, ground vector) of -J (sum ZZ + g sum X) on an open chain: E0 from the free-fermion solution (minus the sum of the
    singular values of Jg I - J S), the vector from a Lanczos run on an index-arithmetic H psi -- no oracle involved."""
    from scipy.sparse.linalg import LinearOperator, eigsh

    E0 = -float(np.sum(np.linalg.svd(J * g * np.eye(n) - J * np.eye(n, k=1), compute_uv=False)))
    idx = np.arange(2 ** n, dtype=np.int64)
    z = [1 - 2 * ((idx >> i) & 1) for i in range(n)]
    diag = -J * sum(z[i] * z[i + 1] for i in range(n - 1)).astype(np.float64)
    mv = lambda v: diag * np.asarray(v).reshape(-1) - J * g * sum(np.asarray(v).reshape(-1)[idx ^ (1 << i)] for i in range(n))
    w, v = eigsh(LinearOperator((2 ** n, 2 ** n), matvec=mv, dtype=np.float64), k=1, which="SA", tol=1e-12)
    assert abs(w[0] - E0) < 1e-9 * n
    return E0, (v[:, 0] / np.linalg.norm(v[:, 0])).astype(np.complex128)


@pytest.mark.parametrize("n,J,g", [(15, 1.0, 1.4), (16, 0.7, 1.0)])
def test_emu_expectation_kernels_reproduce_the_free_fermion_ground_energy(emu, n, J, g):
    """No oracle involved: every expectation-value kernel replayed on the CPU (round-1 tiles, their persistent form, the
    TMA-fed flip tiles, the sharded form with 4 simulated ranks, and <psi|H psi> of the tiled H psi sweep) returns the
    free-fermion ground energy of the open transverse-field chain for its ground vector."""
    E0, gs = _tfim_ground_state(n, J, g)
    hp = (J, 0.0, g)
    tol = 1e-10 * n
    assert abs(emu.measure(gs, 0, hp, 11, 3)[0] - E0) < tol
    assert abs(emu.measure_persistent(gs, 0, hp, 10, 3, 5) - E0) < tol
    d, f, _ = emu.flip_measure(gs, 0, hp, grid=3)
    assert abs(d - J * g * f - E0) < tol
    assert abs(emu.sharded_measure(gs, 2, 0, hp, 9, 3) - E0) < tol
    lam, E = emu.apply_ham_tiled(gs, 0, hp, 3)
    assert abs(E.real - E0) < tol and abs(E.imag) < tol
    assert np.max(np.abs(lam - E0 * gs)) < 1e-7  # H psi = E0 psi up to the Lanczos residual
    E32 = emu.flip_measure(gs.astype(np.complex64), 0, hp, grid=2)
    assert abs(E32[0] - J * g * E32[1] - E0) < 1e-5 * n


def test_emu_flip_measure_simulated_shards(emu, oracle):
    """local part of a sharded expectation value: every simulated rank sweeps its shard with c_hi = rank << nloc; the sum of
    the diagonal parts and of the local-qubit flips equals the oracle's value minus the flips of the rank-bit qubits."""
    n, g = 16, 2
    nloc = n - g
    rng = np.random.default_rng(1602)
    psi = rand_state(rng, n)
    for kind, hp in [(0, (1.0, 0.1, 0.5)), (1, (0.1, -0.33, 0.8))]:
        coef = -hp[0] * hp[2] if kind == 0 else hp[1]
        tot = 0.0
        for r in range(1 << g):
            d, f, _ = emu.flip_measure(psi[r << nloc:(r + 1) << nloc], kind, hp, grid=2, n_total=n, c_hi=r << nloc)
            tot += d + coef * f
        # flips of the qubits nloc..n-1, from the definition (src/hamiltonians/tfim.jl:39-44, east_model.jl:37-44)
        top = 0.0
        idx = np.arange(1 << n)
        for i in range(nloc, n):
            w = np.ones(1 << n) if kind == 0 else ((idx >> (i - 1)) & 1 == 0).astype(float)
            top += float(np.sum(w * (np.conj(psi[idx ^ (1 << i)]) * psi).real))
        ref = (oracle.measure_tfim(psi, *hp) if kind == 0 else oracle.measure_east(psi, *hp)).real
        assert abs(tot + coef * top - ref) < 1e-12 * max(1.0, abs(ref))


# ---- TMA-fed gate pass (tile_tma.cuh): stage programs with the hardware swizzle, emulated box loads / stores ------------------
@pytest.mark.parametrize("n,layers,maxg", [(11, 6, 0), (12, 9, 0), (13, 12, 3), (14, 7, 0), (15, 5, 2), (16, 4, 0)])
def test_emu_tma_gate_pass(emu, oracle, n, layers, maxg):
    rng = np.random.default_rng(9000 + 31 * n + layers)
    q0 = [j - 1 for l in range(1, layers + 1) for j in range(1 + (l - 1) % 2, n, 2)]
    mats = []
    for _ in q0:
        u = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        mats.append(u / np.linalg.det(u) ** 0.25)  # non-unitary "unitaries", as in test/correctness_tests.jl:90-98
    psi0 = rand_state(rng, n)
    ref = psi0
    for q, M in zip(q0, mats):
        ref = oracle.apply_2q(ref, M, int(q) + 1)
    out, (npass, ntma) = emu.apply_gates_tma(psi0, q0, mats, max_gates=maxg)
    assert ntma == npass and npass >= 1
    assert rel(out, ref) < 1e-12
    # the LDG form of the same plan gives the same numbers bit for bit (same stage programs, same arithmetic order; the
    # TMA kernel runs the ordinary ComplexF64 arithmetic, so the LDG form is replayed with it too)
    try:
        emu.set_m3(False)
        out2, _ = emu.apply_gates(psi0, q0, mats, 11, 3, max_gates=maxg)
    finally:
        emu.set_m3(True)
    assert np.array_equal(out, out2)
