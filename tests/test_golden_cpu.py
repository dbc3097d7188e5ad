"""CPU: the committed golden fixtures still equal what the oracle produces, and the CPU replay of the
device tile-pass code (tests/emu) reproduces the fixture states."""
import os
import shutil

import numpy as np
import pytest

from oracle import oracle_np as onp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_brickwork_fixture_matches_oracle(oracle):
    d = np.load(os.path.join(GOLD, "brickwork_12q_20l.npz"))
    for tag in ("uniform", "small"):
        psi = oracle.apply_brickwork(onp.zero_state(12), 12, 20, d[f"angles_{tag}"])
        np.testing.assert_allclose(psi, d[f"psi_{tag}"], atol=1e-14)
        assert np.linalg.norm(d[f"psi32_{tag}"].astype(np.complex128) - psi) < 1e-5


def test_gradient_fixture_matches_oracle(oracle):
    d = np.load(os.path.join(GOLD, "gradients_8q_4l.npz"))
    E, g = oracle.gradients_brickwork(onp.zero_state(8), 8, 4, d["angles_small"], 0, (1.0, 0.1, 0.5))
    assert abs(E - float(d["tfim_E_small"])) < 1e-13
    np.testing.assert_allclose(g, d["tfim_grads_small"], atol=1e-14)


def test_measure_fixture_matches_dense():
    d = np.load(os.path.join(GOLD, "measure_10q.npz"))
    psi = d["psi"]
    assert abs(onp.measure_matrix(onp.dense_tfim(10, 1.0, 0.1, 0.5), psi) - d["tfim"][0]) < 1e-12
    assert abs(onp.measure_matrix(onp.dense_east(10, 0.1, -0.1), psi) - d["east"][0]) < 1e-12


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_emulated_device_code_reproduces_fixture(oracle):
    import emu_helper

    d = np.load(os.path.join(GOLD, "brickwork_12q_20l.npz"))
    ang = d["angles_uniform"]
    q0 = np.array(onp.brickwork_gate_positions(12, 20)) - 1
    mats = [oracle.build_general_unitary_gate(ang[:, g]) for g in range(ang.shape[1])]
    out, (npass, nst) = emu_helper.apply_gates(onp.zero_state(12), q0, mats, 11, 3)
    assert np.linalg.norm(out - d["psi_uniform"]) < 1e-13
    g = np.load(os.path.join(GOLD, "generic_gates_9q.npz"))
    out, _ = emu_helper.apply_gates(g["psi0"], g["K"] - 1, list(g["mats"]), 9, 3)
    assert np.linalg.norm(out - g["psi"]) / np.linalg.norm(g["psi"]) < 1e-12
