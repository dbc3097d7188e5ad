"""Generates tests/golden/*.npz from the CPU oracle (oracle/qc_oracle.c) with fixed Philox seeds.

The reference (pure Julia) cannot run in this image and stores no vectors of its own, so these
fixtures are oracle outputs; the oracle itself is pinned by tests/test_oracle.py.  Re-run with
`python tests/golden/make_golden.py` -- the files are small and committed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402


def rng_for(seed):
    return np.random.Generator(np.random.Philox(seed))


def rand_state(rng, n):
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    return v / np.linalg.norm(v)


def main():
    oracle.build()
    # 1) brickwork apply, C1 shape (12 qubits x 20 half-layers), angles U[0, 2pi) and N(0,1)*0.01
    out = {}
    for tag, seed, scale in [("uniform", 1201, None), ("small", 1202, 0.01)]:
        rng = rng_for(seed)
        G = onp.brickwork_num_gates(12, 20)
        ang = rng.uniform(0, 2 * np.pi, (15, G)) if scale is None else rng.standard_normal((15, G)) * scale
        psi = oracle.apply_brickwork(onp.zero_state(12), 12, 20, ang)
        out[f"angles_{tag}"] = ang
        out[f"psi_{tag}"] = psi
        out[f"psi32_{tag}"] = oracle.apply_brickwork(onp.zero_state(12, np.complex64), 12, 20, ang)
    np.savez_compressed(os.path.join(HERE, "brickwork_12q_20l.npz"), **out)

    # 2) generic (non-unitary) gate list, n = 9
    rng = rng_for(901)
    n, G = 9, 40
    K = rng.integers(1, n, size=G)
    mats = []
    for _ in range(G):
        u = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
        mats.append(u / np.linalg.det(u) ** 0.25)
    psi0 = rand_state(rng, n)
    psi = psi0
    for k, M in zip(K, mats):
        psi = oracle.apply_2q(psi, M, int(k))
    np.savez_compressed(os.path.join(HERE, "generic_gates_9q.npz"), K=K, mats=np.array(mats), psi0=psi0, psi=psi)

    # 3) energies + gradients (reference O(G^2) shift sweep), n = 8, 4 half-layers (examples/test_gradients.jl shape)
    rng = rng_for(801)
    n, L = 8, 4
    G = onp.brickwork_num_gates(n, L)
    res = {}
    for tag, ang in [("small", rng.standard_normal((15, G)) * 0.01), ("uniform", rng.uniform(0, 2 * np.pi, (15, G)))]:
        res[f"angles_{tag}"] = ang
        E, g = oracle.gradients_brickwork(onp.zero_state(n), n, L, ang, 0, (1.0, 0.1, 0.5))
        res[f"tfim_E_{tag}"], res[f"tfim_grads_{tag}"] = E, g
        A, B = onp.east_params(0.1, -0.1)
        E, g = oracle.gradients_brickwork(onp.zero_state(n), n, L, ang, 1, (0.1, A, B))
        res[f"east_E_{tag}"], res[f"east_grads_{tag}"] = E, g
    np.savez_compressed(os.path.join(HERE, "gradients_8q_4l.npz"), **res)

    # 4) measure on random states n = 10
    rng = rng_for(1001)
    psi = rand_state(rng, 10)
    A, B = onp.east_params(0.1, -0.1)
    np.savez_compressed(os.path.join(HERE, "measure_10q.npz"), psi=psi,
                        tfim=np.array([oracle.measure_tfim(psi, *p).real for p in [(1.0, 0.1, 0.5), (1.0, 0.9045, 1.4)]]),
                        east=np.array([oracle.measure_east(psi, 0.1, A, B).real]))

    # 5) optimise! trajectory, n = 6, 2 half-layers, 5 epochs
    rng = rng_for(601)
    n, L = 6, 2
    G = onp.brickwork_num_gates(n, L)
    ang = rng.standard_normal((15, G)) * 0.01
    energies, final = oracle.optimise_brickwork(onp.zero_state(n), n, L, ang, 0, (1.0, 0.2, 0.5), 5, 0.01)
    np.savez_compressed(os.path.join(HERE, "optimise_6q.npz"), angles=ang, energies=energies, final=final)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
