#!/usr/bin/env python
"""Timing of the expectation-value kernels: tiled sweep-and-reduce (measure_tile.cuh) vs the gather form (k_measure),
TFIM and East, both dtypes.  usage: bench_measure.py [nqubits=30]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import qcb200 as qc

    qc.init(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    if os.environ.get("MEASURE_PERSISTENT"):
        qc.set_option("measure_persistent", int(os.environ["MEASURE_PERSISTENT"]))
        print("measure_persistent", os.environ["MEASURE_PERSISTENT"])
    if os.environ.get("MEASURE_CTAS"):
        qc.set_option("measure_ctas", int(os.environ["MEASURE_CTAS"]))
        print("measure_ctas", os.environ["MEASURE_CTAS"])
    c = qc.GenericBrickworkCircuit(n, 2)
    c.gate_angles[...] = np.random.default_rng(n).uniform(0, 2 * np.pi, c.gate_angles.shape)
    print(f"== measure {n}q")
    for dt in (np.complex128, np.complex64):
        st = qc.B200State(n, dt)
        qc.apply_(st, st, c)
        for H in (qc.TFIMHamiltonian(1.0, 0.1, 0.5), qc.EastModelHamiltonian(0.1, -0.1)):
            for simple in (0, 1):
                qc.set_option("force_simple", simple)
                E = qc.measure_(None, H, st)
                torch.cuda.synchronize()
                reps = 5
                t0 = time.perf_counter()
                for _ in range(reps):
                    E = qc.measure_(None, H, st)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) / reps * 1e3
                passes = qc.get_stat("passes") if not simple else 1
                gbs = passes * (np.dtype(dt).itemsize << n) / ms / 1e6  # algorithmic: one read of the state per pass
                print(f"{np.dtype(dt).name} {type(H).__name__} {'gather' if simple else 'tiled'} E={E:.12f} ms={ms:.2f} passes={int(passes)} "
                      f"GB/s={gbs:.0f}", flush=True)
        qc.set_option("force_simple", 0)
        del st


if __name__ == "__main__":
    main()
