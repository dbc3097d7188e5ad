"""Worker for the multi-GPU parity tests (launched with torch.distributed.run, one rank per GPU).
Compares the sharded path against the oracle on the same seeded inputs; rank 0 prints PASS/FAIL lines."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import oracle
    import qcb200 as qc
    from qcb200 import dist as qdist
    from oracle import oracle_np as onp

    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    qdist.init_from_torch()
    ok = True
    cases = [(20, 12, np.complex128), (21, 30, np.complex128), (22, 9, np.complex64), (23, 16, np.complex128)]
    for n, layers, dt in cases:
        rng = np.random.default_rng(n * 100 + layers)
        c = qc.GenericBrickworkCircuit(n, layers)
        c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
        v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        psi0 = (v / np.linalg.norm(v)).astype(dt)
        st = qdist.ShardedState(n, dt)
        per = 1 << st.local_qubits
        st.upload_shard(psi0[rank * per:(rank + 1) * per])
        st.apply_circuit(c)
        swaps, fused = qc.get_stat("swaps"), qc.get_stat("fused_swaps")
        n2 = st.norm2()
        shard = st.download_shard()
        if rank == 0:
            ref = oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, c.gate_angles)
        else:
            ref = None
        box = [ref]
        dist.broadcast_object_list(box, src=0)
        ref = box[0][rank * per:(rank + 1) * per]
        err2 = torch.tensor([np.linalg.norm(shard - ref) ** 2, np.linalg.norm(ref) ** 2], dtype=torch.float64, device="cuda")
        dist.all_reduce(err2)
        relerr = float(torch.sqrt(err2[0] / err2[1]).item())
        tol = 1e-12 if dt == np.complex128 else 1e-5
        good = relerr < tol and abs(n2 - 1) < (1e-12 if dt == np.complex128 else 1e-5) and swaps >= 1
        good = good and (not st.peers_mapped or fused >= 1)  # with peers mapped the exchanges ride on gate passes
        ok = ok and good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} sharded n={n} layers={layers} {np.dtype(dt).name} P={world} rel={relerr:.2e} "
                  f"norm2-1={n2 - 1:.1e} swaps={int(swaps)} fused={int(fused)} peers={st.peers_mapped}", flush=True)
        # the NCCL-only path must give the same state
        qc.set_option("fuse_swaps", 0)
        st.upload_shard(psi0[rank * per:(rank + 1) * per])
        st.apply_circuit(c)
        shard2 = st.download_shard()
        qc.set_option("fuse_swaps", 1)
        d2 = torch.tensor([np.linalg.norm(shard2 - shard) ** 2], dtype=torch.float64, device="cuda")
        dist.all_reduce(d2)
        good = float(torch.sqrt(d2[0]).item()) < (1e-13 if dt == np.complex128 else 1e-5)
        ok = ok and good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} fused == nccl-only n={n} diff={float(torch.sqrt(d2[0]).item()):.2e}", flush=True)
        # expectation values on the sharded state (TFIM and East) against the oracle on the full state
        if dt == np.complex128:
            Ht, He = qc.TFIMHamiltonian(1.0, 0.1, 0.5), qc.EastModelHamiltonian(0.1, -0.1)
            Et, Ee = st.measure(Ht), st.measure(He)
            if rank == 0:
                full = oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, c.gate_angles)
                refs = [oracle.measure_tfim(full, 1.0, 0.1, 0.5).real, oracle.measure_east(full, He.c, He.A, He.B).real]
            else:
                refs = None
            box = [refs]
            dist.broadcast_object_list(box, src=0)
            good = abs(Et - box[0][0]) < 1e-11 * max(1, abs(box[0][0])) and abs(Ee - box[0][1]) < 1e-11 * max(1, abs(box[0][1]))
            ok = ok and good
            if rank == 0:
                print(f"{'PASS' if good else 'FAIL'} sharded measure n={n}: TFIM {Et:.12f} vs {box[0][0]:.12f}, East {Ee:.12f} vs {box[0][1]:.12f}", flush=True)
        # round trip: circuit then inverse circuit on the sharded state
        st.apply_circuit(c, adjoint=True)
        back = st.download_shard()
        e2 = torch.tensor([np.linalg.norm(back - psi0[rank * per:(rank + 1) * per]) ** 2], dtype=torch.float64, device="cuda")
        dist.all_reduce(e2)
        rt = float(torch.sqrt(e2[0]).item())
        good = rt < (1e-12 if dt == np.complex128 else 2e-5)
        ok = ok and good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} round-trip n={n} err={rt:.2e}", flush=True)
        st.close()
    # ---- shards without a second buffer (what 34 qubits on 2 GPUs look like): in-place qubit swaps over NVLink peer memory
    # (k_swap_peer) == the staged NCCL exchange == oracle
    qc.set_option("dist_second_buffer", 0)
    for n, layers, dt in [(20, 12, np.complex128), (22, 9, np.complex64)]:
        rng = np.random.default_rng(n * 100 + layers + 7)
        c = qc.GenericBrickworkCircuit(n, layers)
        c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
        v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        psi0 = (v / np.linalg.norm(v)).astype(dt)
        st = qdist.ShardedState(n, dt)
        per = 1 << st.local_qubits
        outs = {}
        for mode in (1, 0):
            qc.set_option("peer_swap", mode)
            p0 = qc.get_stat("peer_swaps")
            st.upload_shard(psi0[rank * per:(rank + 1) * per])
            st.apply_circuit(c)
            swaps, pswaps = qc.get_stat("swaps"), qc.get_stat("peer_swaps") - p0
            outs[mode] = st.download_shard()
            want = swaps >= 1 and (pswaps >= 1 if (mode and st.peer_swap_mapped) else pswaps == 0)
            ok = ok and want
            if rank == 0 and not want:
                print(f"FAIL no-second-buffer n={n}: swaps={swaps} peer_swaps={pswaps} mode={mode} mapped={st.peer_swap_mapped}", flush=True)
        qc.set_option("peer_swap", 1)
        ref = oracle.apply_brickwork(psi0.astype(np.complex128), n, layers, c.gate_angles)[rank * per:(rank + 1) * per]
        err2 = torch.tensor([np.linalg.norm(outs[1] - ref) ** 2, np.linalg.norm(ref) ** 2, np.linalg.norm(outs[1] - outs[0]) ** 2], dtype=torch.float64, device="cuda")
        dist.all_reduce(err2)
        relerr, same = float(torch.sqrt(err2[0] / err2[1]).item()), float(torch.sqrt(err2[2]).item())
        good = relerr < (1e-12 if dt == np.complex128 else 1e-5) and same == 0.0 and not st.peers_mapped
        ok = ok and good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} no second buffer n={n} {np.dtype(dt).name} P={world}: peer swap rel={relerr:.2e}, "
                  f"peer swap == staged NCCL diff={same:.1e}, peer_swap_mapped={st.peer_swap_mapped}", flush=True)
        st.close()
    qc.set_option("dist_second_buffer", 1)
    # ---- adjoint gradients on the sharded state vs the single-GPU gradient of the same circuit (both Hamiltonians, both
    # dtypes; the ComplexF64 backward pass runs on the tensor cores, the scalar pass is checked with adjoint_mma = 0) and,
    # at a size the oracle handles, vs the oracle's restatement of the reference's shift sweep ----------------------------
    gcases = [(20, 4, qc.TFIMHamiltonian(1.0, 0.1, 0.5), 1, np.complex128), (21, 3, qc.EastModelHamiltonian(0.1, -0.1), 1, np.complex128),
              (20, 3, qc.TFIMHamiltonian(1.0, 0.9045, 1.4), 0, np.complex128), (21, 4, qc.TFIMHamiltonian(1.0, 0.1, 0.5), 1, np.complex64),
              (15, 3, qc.TFIMHamiltonian(1.0, 0.1, 0.5), 1, np.complex128)]
    for n, layers, Hq, mma, dt in gcases:
        if n - int(np.log2(world)) < 14:  # qcb_state_create_sharded: at least 14 local qubits
            continue
        qc.set_option("adjoint_mma", mma)
        rng = np.random.default_rng(400 + n + layers)
        c = qc.GenericBrickworkCircuit(n, layers)
        c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
        v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        psi0 = (v / np.linalg.norm(v)).astype(dt)
        E_ref, g_ref = qc.gradients(Hq, qc.B200State(psi0), c, calculate_energy=True)  # every rank, on its own GPU
        if n <= 15:
            E_o, g_o = oracle.gradients_brickwork(psi0.astype(np.complex128), n, layers, c.gate_angles, Hq.kind, Hq.params())
            E_ref, g_ref = E_o, g_o
        st = qdist.ShardedState(n, dt)
        per = 1 << st.local_qubits
        st.upload_shard(psi0[rank * per:(rank + 1) * per])
        E_sh, g_sh = st.gradients(Hq, c, calculate_energy=True)
        tol = 1e-12 if dt == np.complex128 else 1e-5
        scale = max(1.0, float(np.max(np.abs(g_ref))))
        dE, dg = abs(E_sh - E_ref), float(np.max(np.abs(g_sh - g_ref)))
        good = dE < tol * max(1.0, abs(E_ref)) and dg < tol * scale
        g_again = st.gradients(Hq, c)  # the input state is untouched: same numbers again
        good = good and float(np.max(np.abs(g_again - g_sh))) == 0.0
        back = st.download_shard()
        good = good and float(np.max(np.abs(back - psi0[rank * per:(rank + 1) * per]))) == 0.0
        t = torch.tensor([0 if good else 1], device="cuda")
        dist.all_reduce(t)
        good = t.item() == 0
        ok = ok and good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} sharded gradients n={n} layers={layers} {type(Hq).__name__} {np.dtype(dt).name} mma={mma} "
                  f"dE={dE:.2e} dgrad={dg:.2e}", flush=True)
        st.close()
    qc.set_option("adjoint_mma", 1)
    # ---- optimise! on a sharded state (its final energy runs circuit + measure on a work copy that must not inherit
    # psi0's peer views): same trajectory as on one GPU, with the swaps fused into gate passes and on NCCL only ----------
    for fuse in (1, 0):
        n, layers, epochs, lr = 20, 3, 3, 0.01
        qc.set_option("fuse_swaps", fuse)
        rng = np.random.default_rng(77)
        a0 = rng.standard_normal((15, qc.brickwork_num_gates(n, layers))) * 0.1
        Hq = qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
        c_ref = qc.GenericBrickworkCircuit(n, layers, gate_angles=a0.copy())
        e_ref = qc.optimise_(c_ref, Hq, qc.B200State(n, np.complex128), epochs, lr)
        st = qdist.ShardedState(n, np.complex128)
        c_sh = qc.GenericBrickworkCircuit(n, layers, gate_angles=a0.copy())
        e_sh = qc.optimise_(c_sh, Hq, st, epochs, lr)
        de, da = float(np.max(np.abs(e_sh - e_ref))), float(np.max(np.abs(c_sh.gate_angles - c_ref.gate_angles)))
        good = de < 1e-11 and da < 1e-12
        t = torch.tensor([0 if good else 1], device="cuda")
        dist.all_reduce(t)
        good = t.item() == 0
        ok = ok and good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} sharded optimise fuse_swaps={fuse} peers={st.peers_mapped} dE={de:.2e} dangles={da:.2e}", flush=True)
        st.close()
    qc.set_option("fuse_swaps", 1)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.barrier()
    qdist.finalize()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
