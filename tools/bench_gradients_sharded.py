#!/usr/bin/env python
"""Timing of the adjoint gradient on a sharded state (run under torch.distributed.run, one rank per GPU): the C5 shape
(4 half-layers, TFIM J=1, h=0.9045, g=1.4) at a size with 2^28 amplitudes per rank, next to the single-GPU 28-qubit
gradient.  Rank 0 prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    import qcb200 as qc
    from qcb200 import dist as qdist

    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    qdist.init_from_torch()
    nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    n = nloc + int(np.log2(world))
    layers, hp = 4, (1.0, 0.9045, 1.4)
    rng = np.random.default_rng(n)
    c = qc.GenericBrickworkCircuit(n, layers)
    c.gate_angles[...] = rng.standard_normal(c.gate_angles.shape) * 0.01
    H = qc.TFIMHamiltonian(*hp)
    out = {"config": "C5 shape, sharded", "nqubits": n, "ranks": world, "local_qubits": nloc, "half_layers": layers, "gates": c.ngates}
    for dt in (np.complex128, np.complex64):
        st = qdist.ShardedState(n, dt)
        E, g = st.gradients(H, c, calculate_energy=True)  # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            E, g = st.gradients(H, c, calculate_energy=True)
        torch.cuda.synchronize()
        dt_s = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt_s, op=dist.ReduceOp.MAX)
        out[np.dtype(dt).name] = {"s_per_gradient": float(dt_s.item()), "energy": E, "swaps": qc.get_stat("swaps"), "grad_max": float(np.max(np.abs(g)))}
        st.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    qdist.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
