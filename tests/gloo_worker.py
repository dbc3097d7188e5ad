"""CPU multi-process check of the sharded protocol (world_size 2, gloo): every process owns one shard, runs the
device tile-pass code through the CPU replay library, and performs the qubit swap as a real all-to-all between
processes.  Exercises block indexing of the exchange + layout bookkeeping across ranks without a GPU."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import emu_helper
    import oracle
    from oracle import oracle_np as onp

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = world.bit_length() - 1
    n, layers = 13, 14
    rng = np.random.default_rng(2026)
    G = onp.brickwork_num_gates(n, layers)
    angles = rng.uniform(0, 2 * np.pi, (15, G))
    q0 = np.ascontiguousarray(np.array(onp.brickwork_gate_positions(n, layers)) - 1, dtype=np.int32)
    mats = np.ascontiguousarray(np.stack([oracle.build_general_unitary_gate(angles[:, i]).reshape(-1, order="F") for i in range(G)]))
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    psi0 = v / np.linalg.norm(v)
    nloc = n - g
    per = 1 << nloc
    shard = psi0[rank * per:(rank + 1) * per].copy()
    nswaps = [0]

    @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int)
    def swap_cb(ptr, nloc_, g_):
        # block j of this rank <-> block `rank` of rank j
        buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2 << nloc_,))
        t_in = torch.from_numpy(buf.copy()).view(world, -1)
        t_out = torch.empty_like(t_in)
        try:
            dist.all_to_all_single(t_out.view(-1), t_in.view(-1))
        except Exception:
            reqs = []
            for j in range(world):
                if j == rank:
                    t_out[j].copy_(t_in[j])
                else:
                    reqs.append(dist.isend(t_in[j].contiguous(), j))
                    reqs.append(dist.irecv(t_out[j], j))
            for r in reqs:
                r.wait()
        buf[:] = t_out.view(-1).numpy()
        nswaps[0] += 1
        if os.environ.get("EMU_DEBUG"): print(rank, "swap", nloc_, g_, float(np.linalg.norm(buf)), flush=True)
        return 0

    phys = np.zeros(n, dtype=np.int32)
    rc = emu_helper.lib().emu_rank_apply_gates(C.c_int(n), C.c_int(g), shard.ctypes.data_as(C.c_void_p), C.c_int(G),
                                               q0.ctypes.data_as(C.c_void_p), mats.ctypes.data_as(C.c_void_p), C.c_int(9), C.c_int(3),
                                               C.c_int(1), swap_cb, phys.ctypes.data_as(C.c_void_p))
    ref = oracle.apply_brickwork(psi0, n, layers, angles)[rank * per:(rank + 1) * per]
    err = float(np.linalg.norm(shard - ref))
    ok = rc == 0 and err < 1e-12 and nswaps[0] >= 1 and list(phys) == list(range(n))
    # ---- adjoint gradients on the sharded state: forward circuit, lambda = H psi around one swap of both states, backward
    # sweep on the pair; the per-rank environments and energy parts are all-reduced here like NCCL does on the GPUs ----------
    n2, layers2 = 12, 5
    G2 = onp.brickwork_num_gates(n2, layers2)
    angles2 = rng.uniform(0, 2 * np.pi, (15, G2))
    q2 = np.ascontiguousarray(np.array(onp.brickwork_gate_positions(n2, layers2)) - 1, dtype=np.int32)
    mats2 = np.ascontiguousarray(np.stack([oracle.build_general_unitary_gate(angles2[:, i]).reshape(-1, order="F") for i in range(G2)]))
    v2 = rng.standard_normal(2 ** n2) + 1j * rng.standard_normal(2 ** n2)
    psi2 = v2 / np.linalg.norm(v2)
    per2 = 1 << (n2 - g)
    shard2 = np.ascontiguousarray(psi2[rank * per2:(rank + 1) * per2])
    hp = (1.0, 0.1, 0.5)
    env = np.zeros((G2, 32))
    E2 = np.zeros(2)
    swaps_before = nswaps[0]
    rc2 = emu_helper.lib().emu_rank_gradient_env(C.c_int(n2), C.c_int(g), C.c_int(rank), shard2.ctypes.data_as(C.c_void_p), C.c_int(0),
                                                 C.c_double(hp[0]), C.c_double(hp[1]), C.c_double(hp[2]), C.c_int(G2), q2.ctypes.data_as(C.c_void_p),
                                                 mats2.ctypes.data_as(C.c_void_p), C.c_int(9), C.c_int(3), swap_cb, env.ctypes.data_as(C.c_void_p),
                                                 E2.ctypes.data_as(C.c_void_p))
    t = torch.from_numpy(np.concatenate([env.reshape(-1), E2]))
    dist.all_reduce(t)
    tot = t.numpy()
    envc = (tot[:-2].reshape(G2, 32)[:, 0::2] + 1j * tot[:-2].reshape(G2, 32)[:, 1::2]).reshape(G2, 4, 4).transpose(0, 2, 1)
    from qcb200 import _lib
    grads = np.zeros((15, G2))
    for k in range(G2):
        a = np.ascontiguousarray(angles2[:, k])
        e = np.ascontiguousarray(envc[k].reshape(-1, order="F"))
        out = np.zeros(15)
        _lib.check(_lib.lib().qcb_general_unitary_gate_gradient(a.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        grads[:, k] = out
    E_ref, g_ref = oracle.gradients_brickwork(psi2, n2, layers2, angles2, 0, hp)
    gerr = float(np.max(np.abs(grads - g_ref)))
    ok2 = rc2 == 0 and abs(tot[-2] - E_ref) < 1e-12 * max(1.0, abs(E_ref)) and abs(tot[-1]) < 1e-12 and gerr < 1e-12 and nswaps[0] - swaps_before >= 3
    ok = ok and ok2
    # ---- a known answer without the oracle: a circuit without entangling angles maps |0..0> to an analytic product state
    # (tests/product_circuit.py); matrices from the library's host gate builder ------------------------------------------
    import qcb200
    from product_circuit import product_amplitudes, product_circuit_vectors

    angles3 = rng.uniform(0, 2 * np.pi, (15, G))
    angles3[12:15] = 0
    mats3 = np.ascontiguousarray(np.stack([qcb200.build_general_unitary_gate(angles3[:, i]).mat().reshape(-1, order="F") for i in range(G)]))
    shard3 = np.zeros(per, dtype=np.complex128)
    if rank == 0:
        shard3[0] = 1
    phys3 = np.zeros(n, dtype=np.int32)
    rc3 = emu_helper.lib().emu_rank_apply_gates(C.c_int(n), C.c_int(g), shard3.ctypes.data_as(C.c_void_p), C.c_int(G),
                                                q0.ctypes.data_as(C.c_void_p), mats3.ctypes.data_as(C.c_void_p), C.c_int(9), C.c_int(3),
                                                C.c_int(1), swap_cb, phys3.ctypes.data_as(C.c_void_p))
    want3 = product_amplitudes(product_circuit_vectors(n, layers, angles3), np.arange(rank * per, (rank + 1) * per))
    perr = float(np.max(np.abs(shard3 - want3)))
    ok = ok and rc3 == 0 and perr < 1e-14 and list(phys3) == list(range(n))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print(("PASS" if all(flags) else "FAIL"), f"gloo sharded world={world} rc={rc} err={err:.2e} swaps={nswaps[0]}; "
              f"gradients rc={rc2} dE={abs(tot[-2] - E_ref):.2e} dgrad={gerr:.2e}; product state rc={rc3} err={perr:.2e}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
