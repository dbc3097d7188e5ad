#!/usr/bin/env python
"""Timing of the gradient / optimisation configs of BASELINE.json (C2: 20 q TFIM energy + gradients, C5: 28 q variational
loop) next to the reference algorithm's cost on the CPU (per-apply / per-measure oracle timings x the reference's exact
operation counts 2G + 15G(G+1) applies, 30G measures).  Prints one JSON line per config."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import oracle
    import qcb200 as qc
    from oracle import oracle_np as onp

    qc.init(0)
    cfgs = [("C2", 20, 4, (1.0, 0.1, 0.5), np.complex128), ("C5", 28, 4, (1.0, 0.9045, 1.4), np.complex128),
            ("C5", 28, 4, (1.0, 0.9045, 1.4), np.complex64), ("C1", 12, 20, (1.0, 0.1, 0.5), np.complex128)]
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if len(sys.argv) > 2:
        qc.set_option("adjoint_tile_bits", int(sys.argv[2]))
    # backward-pass variants: ADJ_MMA=0 selects the scalar-FMA pass (adjoint.cuh), ADJ_WARPS=4|8 the CTA size of the tensor-core pass
    adj_mma = int(os.environ.get("ADJ_MMA", "1"))
    qc.set_option("adjoint_mma", adj_mma)
    qc.set_option("adjoint_warps", int(os.environ.get("ADJ_WARPS", "0")))
    for tag, n, layers, hp, dt in cfgs:
        if only and (tag != only or dt != np.complex128):
            continue
        rng = np.random.default_rng(n)
        c = qc.GenericBrickworkCircuit(n, layers)
        c.gate_angles[...] = rng.standard_normal(c.gate_angles.shape) * 0.01
        H = qc.TFIMHamiltonian(*hp)
        psi0 = qc.B200State(n, dt)
        qc.gradients(H, psi0, c, calculate_energy=True)  # warm-up
        torch.cuda.synchronize()
        reps = int(os.environ.get("REPS", "1" if only else "3"))
        t0 = time.perf_counter()
        for _ in range(reps):
            E, g = qc.gradients(H, psi0, c, calculate_energy=True)
        torch.cuda.synchronize()
        dt_gpu = (time.perf_counter() - t0) / reps
        launches = qc.get_stat("launches")
        # forward only (circuits/s for the small config)
        t0 = time.perf_counter()
        for _ in range(reps):
            qc.apply_(psi0.similar(), psi0, c)
        torch.cuda.synchronize()
        dt_fwd = (time.perf_counter() - t0) / reps
        # standalone expectation value (tiled sweep-and-reduce)
        st = qc.B200State(n, dt)
        qc.apply_(st, st, c)
        qc.measure_(None, H, st)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            qc.measure_(None, H, st)
        dt_meas = (time.perf_counter() - t0) / reps
        meas_passes = qc.get_stat("passes")
        del st
        # CPU: one apply + one measure of the oracle at this size, scaled by the reference's operation counts
        G = c.ngates
        npdt = dt
        n_cpu = min(n, 26)
        psi = np.zeros(1 << n_cpu, dtype=npdt)
        psi[0] = 1
        ang = rng.standard_normal((15, onp.brickwork_num_gates(n_cpu, layers))) * 0.01
        t0 = time.perf_counter()
        oracle.apply_brickwork(psi, n_cpu, layers, ang, 0, 2, inplace=True)
        t_apply = (time.perf_counter() - t0) / 2 * 2 ** (n - n_cpu)
        t0 = time.perf_counter()
        oracle.measure_tfim(psi, *hp)
        t_meas = (time.perf_counter() - t0) * 2 ** (n - n_cpu)
        napply, nmeas = 2 * G + 15 * G * (G + 1), 30 * G
        cpu_est = napply * t_apply + nmeas * t_meas
        print(json.dumps({"config": tag, "nqubits": n, "half_layers": layers, "gates": G, "dtype": np.dtype(dt).name, "adjoint_mma": adj_mma, "adjoint_warps": int(os.environ.get("ADJ_WARPS", "0")),
                          "gpu_s_per_gradient": dt_gpu, "gpu_s_forward": dt_fwd, "gpu_s_measure": dt_meas, "measure_passes": meas_passes, "gpu_launches_per_gradient": launches, "energy": E,
                          "cpu_reference_algorithm_s_estimate": cpu_est, "cpu_threads": oracle.max_threads(),
                          "cpu_note": f"oracle per-apply {t_apply:.4f} s, per-measure {t_meas:.4f} s (measured at {n_cpu} q, scaled 2^{n - n_cpu}) x "
                                      f"{napply} applies + {nmeas} measures of src/gradients.jl:88-152",
                          "speedup_vs_cpu_estimate": cpu_est / dt_gpu}), flush=True)


if __name__ == "__main__":
    main()
