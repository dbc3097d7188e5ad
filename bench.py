#!/usr/bin/env python
"""bench.py -- gate-applications/s of the dense-statevector hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W          our arm, 30-qubit brickwork depth 100, ComplexF64
  torchrun ... bench.py --gpus N ...                      our arm, 34-qubit circuit sharded over N GPUs
  python bench.py --impl reference ...                    the reference's CPU path (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" is one application of the whole circuit to the resident state.
`value` is measured with inputs resident in HBM; `e2e` goes through the public host-buffer call
(`apply(psi::Array, circuit)` -> qcb_apply_brickwork_host) with pinned host buffers, H2D and D2H inside
the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
UNIT_MULTI = "gate-apps/s in 2^30-amplitude units (n-qubit gate-app counts 2^(n-30))"
NOMINAL_FP64_TFLOPS = 37.0  # 148 SMs x 64 DFMA/clk x 1.965 GHz x 2 (datasheet-level figure; measured one preferred)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--nqubits", type=int, default=None)
    p.add_argument("--layers", type=int, default=100)
    p.add_argument("--dtype", default="c128", choices=["c128", "c64"])
    p.add_argument("--tile-bits", type=int, default=None)
    p.add_argument("--low-bits", type=int, default=None)
    p.add_argument("--max-gates-per-pass", type=int, default=None)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the C1/C2/C5/measure timings appended at --gpus 1")
    p.add_argument("--no-parity-check", action="store_true", help="skip the sharded-vs-single-GPU check at --gpus N > 1")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the bounded CPU sample")
    p.add_argument("--seed", type=int, default=20261019)
    return p.parse_args()


def workload(args):
    n = args.nqubits if args.nqubits is not None else (30 if args.gpus == 1 else 34)
    return n, args.layers


def brickwork_num_gates(n, layers):
    """src/circuits.jl:10-12: odd half-layers hold n // 2 gates, even ones (n - 1) // 2."""
    return sum(len(range(1 + (l - 1) % 2, n, 2)) for l in range(1, layers + 1))


def config_of(n, layers, G, dtype, world):
    """The workload description -- identical in both arms (ours / --impl reference)."""
    cname = "ComplexF64" if dtype == "c128" else "ComplexF32"
    amp = 16 if dtype == "c128" else 8
    gbits = max(0, world.bit_length() - 1)
    return {"workload": f"{n}-qubit brickwork circuit, {layers} half-layers ({G} gates), {cname}, psi0=|0...0>, angles U[0,2pi)",
            "nqubits": n, "half_layers": layers, "gates": G, "state_dtype": cname,
            "state_bytes_per_gpu": int(amp * 2 ** (n - gbits)),
            "l2_policy": "inputs larger than L2: the state (>= 16 GiB per GPU) is far larger than the 126 MB L2",
            "sharding": "none" if world == 1 else f"top {gbits} qubits over {world} ranks",
            "scaling_note": ("N=1 runs the 30-qubit config (a 34-qubit ComplexF64 state does not fit one GPU); N=2,4,8 run the same "
                             "34-qubit circuit (fixed total work); values are quoted in 2^30-amplitude units so that the driver's "
                             "v_N / (N v_1) compares per-GPU throughput")}


def make_angles(n, layers, seed):
    """angles ~ U[0, 2pi) (examples/hpc/barren_plateau_trial.jl:14-15), Philox counter RNG, 15 x G column-major."""
    G = brickwork_num_gates(n, layers)
    rng = np.random.Generator(np.random.Philox(seed))
    return np.asfortranarray(rng.uniform(0, 2 * np.pi, (15, G))), G


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's execution shape (oracle/qc_oracle.c), bounded sample
# ------------------------------------------------------------------------------------------------
def host_cores():
    """Threads the CPU arm may use: every core this process is allowed on.  (torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers; the arm ignores that and asks OpenMP for this many threads explicitly.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class CpuArm:
    """The reference's CPU path (oracle/qc_oracle.c: zero fill + one work-item per output amplitude, src/apply.jl:62-84,135,
    two ping-pong buffers like src/circuits.jl:19-31) on a bounded sample of the workload's gates.

    The state is at most 30 qubits (2 x 16 GiB ComplexF64 on the host); for the 34-qubit sharded config the per-amplitude
    rate at n_run qubits is scaled by 2^(n_run - n) -- the reference's cost per gate is proportional to 2^n.  A sample is
    three gates of the circuit's first half-layer: lowest, middle and top qubit pair (cheap, typical and cache-hostile
    strides), each timed on its own after an untimed first-touch gate."""

    def __init__(self, n, layers, angles, dtype, threads=None):
        import oracle

        oracle.build()
        self.o = oracle
        self.n, self.layers = n, layers
        self.npdt = np.complex128 if dtype == "c128" else np.complex64
        self.cores = threads or host_cores()
        n_run = min(n, 30)
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = 48 << 30
        while 2 * (np.dtype(self.npdt).itemsize << n_run) > 0.6 * avail and n_run > 20:
            n_run -= 1
        self.n_run = n_run
        self.angles = angles if n_run == n else make_angles(n_run, layers, 1)[0]
        self.a = np.zeros(1 << n_run, dtype=self.npdt)
        self.a[0] = 1
        self.b = np.empty_like(self.a)
        G1 = n_run // 2  # gates of the first half-layer: gate g acts on qubits (2g+1, 2g+2), 1-based
        self.sample_gates = sorted({0, G1 // 2, G1 - 1})
        self.mats = {g: oracle.build_general_unitary_gate(self.angles[:, g]) for g in range(G1)}
        oracle.set_threads(self.cores)
        self._gate(1 if G1 > 1 else 0)  # untimed: first touch of both buffers
        self._gate(1 if G1 > 1 else 0)

    def _gate(self, g):
        self.o.apply_2q(self.a, self.mats[g], 2 * g + 1, out=self.b)
        self.a, self.b = self.b, self.a

    def step(self, threads=None):
        """one bounded sample; returns (gate-apps/s at n qubits, seconds, gates)"""
        self.o.set_threads(threads or self.cores)
        t = 0.0
        for g in self.sample_gates:
            t0 = time.perf_counter()
            self._gate(g)
            t += time.perf_counter() - t0
        self.o.set_threads(self.cores)
        rate = len(self.sample_gates) / t
        return rate * 2.0 ** (self.n_run - self.n), t, len(self.sample_gates)

    def describe(self, secs, reps):
        Ks = ", ".join(str(2 * g + 1) for g in self.sample_gates)
        d = (f"{reps} x gates of the first half-layer at K = {Ks} (lowest, middle and top qubit pair) of the {self.n_run}-qubit circuit, "
             f"{secs:.2f} s, {self.cores} OpenMP threads (C restatement of the reference CPU path, not Julia; two persistent ping-pong buffers)")
        if self.n_run != self.n:
            d += f"; the host cannot hold {self.n} qubits: per-amplitude rate at {self.n_run} qubits scaled by 2^{self.n_run - self.n}"
        return d

    def one_thread(self):
        """the reference's default (`julia` starts with 1 thread): one middle gate at <= 26 qubits, scaled per amplitude"""
        m = min(self.n_run, 26)
        a = np.zeros(1 << m, dtype=self.npdt)
        a[0] = 1
        b = np.empty_like(a)
        u = self.o.build_general_unitary_gate(self.angles[:, 0])
        K = 2 * (m // 4) + 1
        self.o.set_threads(1)
        self.o.apply_2q(a, u, K, out=b)
        t0 = time.perf_counter()
        self.o.apply_2q(b, u, K, out=a)
        dt = time.perf_counter() - t0
        self.o.set_threads(self.cores)
        return (1.0 / dt) * 2.0 ** (m - self.n), f"1 gate at K = {K} of a {m}-qubit state, 1 thread, scaled by 2^{m - self.n}"


def cpu_sample(n, layers, angles, dtype, budget_s):
    arm = CpuArm(n, layers, angles, dtype)
    rates, secs, reps = [], 0.0, 0
    while secs < budget_s and reps < 8:
        r, t, _ = arm.step()
        rates.append(r)
        secs += t
        reps += 1
    one, one_desc = arm.one_thread()
    return float(np.mean(rates)), arm.cores, arm.describe(secs, reps), one, one_desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, layers = workload(args)
    angles, G = make_angles(n, layers, args.seed)
    t_start = time.perf_counter()
    arm = CpuArm(n, layers, angles, args.dtype)
    budget = 150.0  # the whole arm is bounded: steps stop being repeated once the budget is used up
    for _ in range(args.warmup):
        if time.perf_counter() - t_start > 0.3 * budget:
            break
        arm.step()
    rates, secs = [], 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, t, _ = arm.step()
        rates.append(r)
        secs += t
        if time.perf_counter() - t_start > budget:
            break
    wall = time.perf_counter() - t0
    done = len(rates)
    value = float(np.mean(rates))
    one, one_desc = arm.one_thread()
    unit = "gate-apps/s"
    if args.gpus > 1:  # same normalisation as our arm: 2^30-amplitude gate applications per second
        value *= 2.0 ** (n - 30)
        one *= 2.0 ** (n - 30)
        unit = UNIT_MULTI
    desc = arm.describe(secs, done) + (f"; {done} of {args.steps} steps ran inside the {budget:.0f} s bound" if done < args.steps else "")
    line = {
        "impl": "reference", "metric": "gate_apps_per_sec", "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, done), "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64" if args.dtype == "c128" else "f32", "data": "synthetic",
        "config": config_of(n, layers, G, args.dtype, args.gpus),
        "details": {"step": "one bounded sample of the workload: 3 gates (lowest, middle, top qubit pair); gate-apps/s = gates / time",
                    "steps_done": done},
        "cpu_baseline": {"value": value, "unit": unit, "cores": arm.cores, "kind": "port", "sample": desc,
                         "value_1_thread": one, "sample_1_thread": one_desc},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def profile_note(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f)
        except Exception:
            return None
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    import qcb200 as qc
    from qcb200 import _lib
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    qc.init(local_rank)
    # Libraries (NCCL's version banner, torch warnings) may write to stdout; rank 0's stdout must carry exactly one JSON
    # line, so file descriptor 1 is pointed at stderr for the duration of the run and the line goes to the saved fd.
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.tile_bits is not None:
        qc.set_option("tile_bits", args.tile_bits)
    if args.low_bits is not None:
        qc.set_option("low_bits", args.low_bits)
    if args.max_gates_per_pass is not None:
        qc.set_option("max_gates_per_pass", args.max_gates_per_pass)

    n, layers = workload(args)
    angles, G = make_angles(n, layers, args.seed)
    circuit = qc.GenericBrickworkCircuit(n, layers, gate_angles=angles)
    npdt = np.complex128 if args.dtype == "c128" else np.complex64
    amp_bytes = np.dtype(npdt).itemsize

    if world == 1:
        state = qc.B200State(n, npdt)  # |0...0> resident in HBM
        step = lambda: qc.apply_(state, state, circuit)  # noqa: E731
        nloc = n
    else:
        from qcb200 import dist as qdist

        qdist.init_from_torch()
        if os.environ.get("QCB_FUSE_SWAPS") == "0":
            qc.set_option("fuse_swaps", 0)
        if os.environ.get("QCB_PEER_SWAP") == "0":  # A/B: staged NCCL exchange instead of the in-place peer swap
            qc.set_option("peer_swap", 0)
        state = qdist.ShardedState(n, npdt)
        step = lambda: state.apply_circuit(circuit)  # noqa: E731
        nloc = state.local_qubits

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    passes = qc.get_stat("passes")
    stages = qc.get_stat("stages")
    launches0 = qc.get_stat("launches_total")
    gates_3m0 = qc.get_stat("gates_3m")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qc.set_option("time_passes", 1)  # CUDA events around every gate-pass launch, on the stream the kernel is launched on
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    pass_count = qc.get_stat("pass_count")
    pass_ms = qc.get_stat("pass_ms")  # sum over the timed region's gate-pass launches (this rank)
    qc.set_option("time_passes", 0)
    clocks = sampler.stop() if rank == 0 else None
    launches = qc.get_stat("launches_total") - launches0
    gates_3m = (qc.get_stat("gates_3m") - gates_3m0) / max(1, args.steps)  # gates per step that ran the 3M arithmetic (this rank)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = float(lt.item())
    ms_step = ms_total / args.steps
    gate_apps = G / (ms_step * 1e-3)
    # whole-job value in 2^30-amplitude gate applications per second (== gate-apps/s for the 30-qubit config)
    value = gate_apps * (2.0 ** (n - 30)) if world > 1 else gate_apps

    # ---- roofline of the dominant kernel (tile_pass_kernel): HBM bytes per launch / average launch duration
    bytes_pass = 2.0 * amp_bytes * (2.0 ** nloc)
    # average duration of one gate-pass launch, from the per-launch events of the timed region (swaps, barriers and
    # bit-permuting copies of the sharded path are NOT in it; permuting passes are launches of the same kernel and are)
    launch_ms = pass_ms / max(pass_count, 1.0)
    peak, peak_src = measured_peaks()
    fp64 = profile_note("r01_fp64_peak.json")
    roofline = None
    if launch_ms:
        achieved = bytes_pass / (launch_ms * 1e-3) / 1e9
        flops = 32.0 * (2.0 ** n) * G / (ms_step * 1e-3) / 1e12 / world  # 16 FMA per amplitude per gate; per GPU
        # FP64-pipe instructions actually executed: a ComplexF64 gate of a pass of >= f64_3m_min_gates gates runs in the 3M form, 48 DFMA/DMUL +
        # 4 DADD per amplitude quad instead of 64 DFMA (tile.cuh); the pipe roof is in instructions x 2 flops
        pipe_frac = 1.0 - (12.0 / 64.0) * (gates_3m / G if args.dtype == "c128" else 0.0)
        # ComplexF32 runs on packed FFMA2, whose measured issue roof is below the scalar-FFMA one
        peak_fl = (fp64 or {}).get("dfma_tflops" if args.dtype == "c128" else "ffma2_reg_operand_tflops") or (fp64 or {}).get("ffma_tflops")
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "binding_roof": "fp_fma_issue (see fma_roof); HBM-bound regime of the same kernel in hbm_bound_regime",
            "traffic": None, "traffic_note": "dram__bytes_read + dram__bytes_write per launch from the ncu --set full capture of this kernel: "
                                             "profiles/README.md (not re-measured inside bench.py)",
            "peak_source": peak_src, "timed_launches": pass_count, "share_of_step": pass_ms / max(ms_total, 1e-9),
            "kernel": f"tile_pass_kernel<{'double' if args.dtype == 'c128' else 'float'}, TB={int(min(qc_opt('tile_bits', 11), nloc))}>",
            "launches_per_step": pass_count / max(1, args.steps), "gate_passes_per_step": passes, "stages_per_step": stages, "gates_per_launch": G / max(passes, 1.0),
            "algorithmic_bytes_per_launch": bytes_pass, "avg_launch_ms": launch_ms,
            "fma_roof": {"achieved_tflops": flops, "executed_pipe_tflops": flops * pipe_frac, "gates_3m_per_step": gates_3m,
                         "pipe_frac_of_peak": flops * pipe_frac / (peak_fl if peak_fl else NOMINAL_FP64_TFLOPS * (1 if args.dtype == "c128" else 2)),
                         "peak_tflops": peak_fl if peak_fl else NOMINAL_FP64_TFLOPS * (1 if args.dtype == "c128" else 2),
                         "peak_source": "measured tools/fp64_peak (profiles/r01_fp64_peak.json)" if peak_fl else "nominal",
                         "note": "achieved_tflops counts the ordinary 64-FMA form (comparable across rounds); executed_pipe_tflops is what the FP64 pipe "
                                 "issued (3M form in the deep passes) and is the one to hold against peak_tflops. "
                                 "With this many gates fused per HBM pass the kernel is bound by FP FMA issue, not HBM; "
                                 "--max-gates-per-pass 4 gives the HBM-bound regime"},
        }

    # ---- the same kernel in its HBM-bound regime (north-star: ">= 70 % of HBM roofline for fused 2-qubit gate passes"):
    # cap the fusion at 3 gates per pass, one warm-up + one timed circuit, then restore the default
    if roofline and world == 1 and args.max_gates_per_pass is None:
        qc.set_option("max_gates_per_pass", 3)
        step()
        torch.cuda.synchronize()
        p3 = qc.get_stat("passes")
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        qc.set_option("time_passes", 1)
        h0.record()
        step()
        h1.record()
        torch.cuda.synchronize()
        ms3 = h0.elapsed_time(h1)
        k3, kms3 = qc.get_stat("pass_count"), qc.get_stat("pass_ms")
        qc.set_option("time_passes", 0)
        launches += 2 * p3
        ach3 = bytes_pass / (kms3 / max(k3, 1.0) * 1e-3) / 1e9
        roofline["hbm_bound_regime"] = {"max_gates_per_pass": 3, "launches_per_circuit": p3, "avg_launch_ms": kms3 / max(k3, 1.0), "achieved": ach3,
                                        "peak": peak, "unit": "GB/s", "frac": ach3 / peak, "gate_apps_per_sec": G / (ms3 * 1e-3),
                                        "circuit_ms": ms3, "share_of_circuit": kms3 / ms3}
        # ... and with one gate per pass (one register stage: what the reference's one-launch-per-gate structure would cost
        # on this kernel), one timed circuit
        qc.set_option("max_gates_per_pass", 1)
        qc.set_option("time_passes", 1)
        h0.record()
        step()
        h1.record()
        torch.cuda.synchronize()
        ms1 = h0.elapsed_time(h1)
        k1, kms1 = qc.get_stat("pass_count"), qc.get_stat("pass_ms")
        qc.set_option("time_passes", 0)
        launches += k1
        ach1 = bytes_pass / (kms1 / max(k1, 1.0) * 1e-3) / 1e9
        roofline["one_gate_per_pass"] = {"launches_per_circuit": k1, "avg_launch_ms": kms1 / max(k1, 1.0), "achieved": ach1, "peak": peak, "unit": "GB/s",
                                         "frac": ach1 / peak, "gate_apps_per_sec": G / (ms1 * 1e-3), "circuit_ms": ms1}
        qc.set_option("max_gates_per_pass", 1 << 30)

    # ---- end to end through the host-buffer API
    e2e = None
    if not args.no_e2e and world == 1:
        nelem = 1 << n
        tdt = torch.complex128 if args.dtype == "c128" else torch.complex64
        h_in = torch.zeros(nelem, dtype=tdt, pin_memory=True)
        h_out = torch.empty(nelem, dtype=tdt, pin_memory=True)
        h_in[0] = 1
        a_in, a_out = h_in.numpy(), h_out.numpy()
        qc.apply_circuit_host(a_in, circuit, out=a_out)  # warm-up (allocates the device staging state)
        torch.cuda.synchronize()
        k = max(1, min(args.steps, 3))
        l0 = qc.get_stat("launches_total")
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            qc.apply_circuit_host(a_in, circuit, out=a_out)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), wall * 1e3) / k
        launches += qc.get_stat("launches_total") - l0
        nrm = 0.0  # the whole result vector is looked at: ||psi_out||^2 must be 1 (unitary circuit on |0...0>)
        for a0 in range(0, nelem, 1 << 26):
            blk = a_out[a0:a0 + (1 << 26)]
            nrm += float(np.vdot(blk, blk).real)
        if abs(nrm - 1.0) > (1e-10 if args.dtype == "c128" else 1e-4):
            raise SystemExit(f"e2e result is not normalised: ||psi||^2 = {nrm!r}")
        e2e = {"value": G / (ms * 1e-3), "unit": "gate-apps/s", "h2d_bytes_per_step": int(nelem * amp_bytes + angles.nbytes),
               "d2h_bytes_per_step": int(nelem * amp_bytes), "ms_per_step": ms, "steps": k,
               "api": "apply(psi::Array, circuit) -> qcb_apply_brickwork_host, pinned host buffers",
               "check_norm2": nrm}
        del h_in, h_out

    if not args.no_e2e and world > 1:
        # sharded end-to-end: every rank moves its canonical shard host -> device, runs the circuit, and reads the
        # canonical result shard back (download includes the qubit swaps that restore the reference's index order).
        # If the host cannot pin the shards (34 q = 256 GiB in total), the step is |0...0> built in HBM + circuit +
        # an all-reduced norm read back to the host, and the line says so.
        per_rank = amp_bytes << nloc
        try:
            import psutil

            avail = psutil.virtual_memory().available / max(1, world)
        except Exception:
            avail = 0
        use_host = avail > 1.6 * per_rank and per_rank <= (72 << 30)  # pinning > 72 GiB per rank takes minutes
        flag = torch.tensor([1 if use_host else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use_host = bool(flag.item())
        k = max(1, min(args.steps, 2))
        hn = None
        if use_host:
            try:
                tdt = torch.complex128 if args.dtype == "c128" else torch.complex64
                h = torch.zeros(1 << nloc, dtype=tdt, pin_memory=True)
                if rank == 0:
                    h[0] = 1
                hn = h.numpy()
            except Exception:  # pinning refused (memlock / cgroup limit): fall back to the device-built input
                hn = None
            ok = torch.tensor([1 if hn is not None else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            use_host = bool(ok.item())
        if use_host:

            def e2e_step():
                state.upload_shard(hn)
                state.apply_circuit(circuit)
                state.download_shard(out=hn)
        else:
            def e2e_step():
                state.set_zero_state()
                state.apply_circuit(circuit)
                return state.norm2()
        e2e_step()
        barrier()
        l0 = qc.get_stat("launches_total")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            e2e_step()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), wall * 1e3) / k
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        launches += (qc.get_stat("launches_total") - l0) * world
        e2e = {"value": G / (ms * 1e-3) * (2.0 ** (n - 30)), "unit": UNIT_MULTI,
               "h2d_bytes_per_step": int(world * per_rank + angles.nbytes) if use_host else int(angles.nbytes * world),
               "d2h_bytes_per_step": int(world * per_rank) if use_host else 8 * world, "ms_per_step": ms, "steps": k,
               "api": ("ShardedState.upload_shard + apply_circuit + download_shard (pinned host shards on every rank)" if use_host else
                       "ShardedState.set_zero_state + apply_circuit + norm2 read-back (host RAM cannot hold the 2^n-amplitude state; "
                       "inputs = gate angles, result = all-reduced norm)")}

    if e2e is not None and world > 1:
        # the result is looked at: squared norm of what came back (host shards) / of the resident result, summed over ranks
        loc = float(np.vdot(hn, hn).real) if use_host else float(state.norm2()) / world
        t = torch.tensor([loc], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        e2e["check_norm2"] = float(t.item())
        if abs(e2e["check_norm2"] - 1.0) > (1e-10 if args.dtype == "c128" else 1e-4):
            raise SystemExit(f"sharded e2e result is not normalised: {e2e['check_norm2']!r}")

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = sharded_parity_check(qc, world, rank)
        launches += parity.pop("_launches", 0)

    extra = None
    if world == 1 and not args.no_extra:
        del state
        qc.trim()
        torch.cuda.empty_cache()
        extra = extra_configs(qc, peak, args)
        launches += sum(e.pop("_launches", 0) for e in extra)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, desc, one, one_desc = cpu_sample(n, layers, angles, args.dtype, args.cpu_seconds)
        cpu = {"value": v, "unit": "gate-apps/s", "cores": cores, "kind": "port", "sample": desc,
               "value_1_thread": one, "sample_1_thread": one_desc}

    if rank == 0:
        line = {
            "metric": "gate_apps_per_sec", "value": value,
            "unit": "gate-apps/s" if world == 1 else UNIT_MULTI,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64" if args.dtype == "c128" else "f32", "data": "synthetic",
            "config": config_of(n, layers, G, args.dtype, world),
            "details": {"swaps": None if world == 1 else ("qubit swaps fused into gate passes as NVLink peer loads (CUDA IPC)"
                                                          if getattr(state, "peers_mapped", False) else
                                                          "in-place qubit swaps over NVLink peer memory (k_swap_peer: no second buffer fits)"
                                                          if getattr(state, "peer_swap_mapped", False) else "qubit swaps via NCCL all-to-all"),
                        "peer_swaps_total": None if world == 1 else qc.get_stat("peer_swaps"),
                        "swaps_per_circuit": None if world == 1 else qc.get_stat("swaps"),
                        "gate_apps_per_sec_raw": gate_apps, "tile_bits": int(qc_opt("tile_bits", 11)), "low_bits": int(qc_opt("low_bits", 3))},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if roofline:
            line["roofline"] = roofline
        if cpu:
            line["cpu_baseline"] = cpu
        if e2e:
            line["e2e"] = e2e
        if parity:
            line["parity_check"] = parity
        if extra:
            line["extra"] = extra
        sys.stdout.flush()
        os.write(_json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        state.close()
        dist.barrier()
        dist.destroy_process_group()


def sharded_parity_check(qc, world, rank):
    """Correctness of the sharded path inside the bench run (the driver's pytest box has one GPU): a 24-qubit circuit,
    TFIM expectation value and adjoint gradient on the sharded state against the single-GPU path on this rank's own GPU
    (which tests/test_gpu_parity.py checks against the oracle).  Every rank compares its own shard; errors are all-reduced."""
    import torch
    import torch.distributed as dist
    from qcb200 import dist as qdist

    n, layers = 24, 12
    rng = np.random.default_rng(2412)
    c = qc.GenericBrickworkCircuit(n, layers)
    c.gate_angles[...] = rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
    H = qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
    l0 = qc.get_stat("launches_total")
    single = qc.B200State(n, np.complex128)
    qc.apply_(single, single, c)
    ref = single.to_numpy()
    E1 = qc.measure_(None, H, single)
    cg = qc.GenericBrickworkCircuit(n, 4)
    cg.gate_angles[...] = rng.standard_normal(cg.gate_angles.shape) * 0.1
    Eg1, g1 = qc.gradients(H, qc.B200State(n, np.complex128), cg, calculate_energy=True)
    st = qdist.ShardedState(n, np.complex128)
    per = 1 << st.local_qubits
    st.apply_circuit(c)
    Es = st.measure(H)
    n2 = st.norm2()
    shard = st.download_shard()
    d = shard - ref[rank * per:(rank + 1) * per]
    t = torch.tensor([float(np.vdot(d, d).real)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    st.set_zero_state()
    Egs, gs = st.gradients(H, cg, calculate_energy=True)
    st.close()
    out = {"config": f"{n}-qubit, {layers} half-layers (apply, TFIM measure) and 4 half-layers (gradient), ComplexF64, {world} ranks vs 1 GPU",
           "rel_apply": float(np.sqrt(t.item())), "dE": abs(Es - E1), "dE_gradient_call": abs(Egs - Eg1),
           "dgrad": float(np.max(np.abs(gs - g1))), "grad_max": float(np.max(np.abs(g1))), "norm2": n2,
           "_launches": (qc.get_stat("launches_total") - l0) * world}
    ok = out["rel_apply"] < 1e-12 and out["dE"] < 1e-11 and out["dgrad"] < 1e-12 * max(1.0, out["grad_max"]) and abs(n2 - 1) < 1e-12
    out["ok"] = bool(ok)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    if flag.item() != 0:
        raise SystemExit(f"sharded parity check failed on some rank: {out}")
    return out


def extra_configs(qc, peak, args):
    """The other BASELINE.json configs, timed in the same driver-run process (CUDA events, >= 5 repetitions after warm-up):
    C5 (28 q x 4 half-layers: TFIM energy + all 810 gradients, ComplexF64 and ComplexF32), C2 (20 q x 4), C1 (12 q x 20:
    forward circuit, and energy + gradient) and the 30-qubit expectation value.  cpu_estimate_s: the oracle's measured cost
    per apply / per measure at that size times the reference algorithm's exact operation counts (2G + 15G(G+1) applies and
    30G measures per gradient, src/gradients.jl:88-152), on the host cores of this box."""
    import torch

    out = []

    def timeit(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(min(ts))

    def cpu_costs(n, dtype, H):
        """seconds per oracle apply (middle gate) and per oracle measure at n qubits"""
        if args.no_cpu_baseline:
            return None, None
        import oracle

        oracle.build()
        oracle.set_threads(host_cores())
        m = min(n, 26)
        a = np.zeros(1 << m, dtype=dtype)
        a[0] = 1
        b = np.empty_like(a)
        u = oracle.build_general_unitary_gate(np.linspace(0.1, 1.5, 15))
        K = 2 * (m // 4) + 1
        oracle.apply_2q(a, u, K, out=b)
        t0 = time.perf_counter()
        oracle.apply_2q(b, u, K, out=a)
        ta = (time.perf_counter() - t0) * 2.0 ** (n - m)
        oracle.measure_tfim(a, H.J, H.h, H.g, scratch=b)
        t0 = time.perf_counter()
        oracle.measure_tfim(a, H.J, H.h, H.g, scratch=b)
        tm = (time.perf_counter() - t0) * 2.0 ** (n - m)
        return ta, tm

    def grad_entry(name, n, layers, dtype, H, scale):
        rng = np.random.default_rng(n * 1000 + layers)
        c = qc.GenericBrickworkCircuit(n, layers)
        c.gate_angles[...] = rng.standard_normal(c.gate_angles.shape) * scale if scale else rng.uniform(0, 2 * np.pi, c.gate_angles.shape)
        G = c.ngates
        psi0 = qc.B200State(n, dtype)
        l0 = qc.get_stat("launches_total")
        res = {}

        def fn():
            res["E"], res["g"] = qc.gradients(H, psi0, c, calculate_energy=True)
        ms, ms_min = timeit(fn)
        per_call = (qc.get_stat("launches_total") - l0) / 7.0
        fw, bw = qc.get_stat("fw_passes"), qc.get_stat("bw_passes")
        amp = np.dtype(dtype).itemsize
        S = float(amp) * 2.0 ** n
        # algorithmic bytes: forward passes read + write the state, lambda = H psi reads psi and writes lambda,
        # a backward pass reads and writes both states
        alg = (2 * fw + 2 + 4 * bw) * S
        ta, tm = cpu_costs(n, dtype, H)
        one_launch = bw == 0 and G > 0  # states of <= 12 qubits: circuit, H psi, backward sweep and environments in one cluster launch
        if one_launch:
            alg = 2 * S  # the state is read once and stays in shared memory; what leaves the kernel is 32 doubles per gate
        e = {"config": name, "ms": ms, "ms_min": ms_min,
             "kernel": ("cluster_gradient_kernel (state in distributed shared memory, one launch) + finish + contraction" if one_launch else
                        "forward tile_pass x%d + H psi + adjoint_pass x%d + contraction" % (fw, bw)),
             "launches_per_call": per_call, "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak,
             "energy": res["E"], "grad_max": float(np.max(np.abs(res["g"]))),
             "cpu_estimate_s": None if ta is None else (2 * G + 15 * G * (G + 1)) * ta + 30 * G * tm,
             "cpu_estimate_note": "oracle seconds per apply / per measure at this size (<= 26 q measured, scaled per amplitude) x the reference "
                                  "algorithm's operation counts", "_launches": qc.get_stat("launches_total") - l0}
        del psi0
        return e

    H5 = qc.TFIMHamiltonian(1.0, 0.9045, 1.4)
    H2 = qc.TFIMHamiltonian(1.0, 0.1, 0.5)
    out.append(grad_entry("C5: 28 q x 4 half-layers (54 gates, 810 angles), TFIM(1, 0.9045, 1.4), ComplexF64, E + gradient", 28, 4, np.complex128, H5, 0.01))
    out.append(grad_entry("C5: 28 q x 4 half-layers, TFIM(1, 0.9045, 1.4), ComplexF32, E + gradient", 28, 4, np.complex64, H5, 0.01))
    qc.trim()
    out.append(grad_entry("C2: 20 q x 4 half-layers (38 gates, 570 angles), TFIM(1, 0.1, 0.5), ComplexF64, E + gradient", 20, 4, np.complex128, H2, 0.01))
    out.append(grad_entry("C1: 12 q x 20 half-layers (110 gates, 1650 angles), TFIM(1, 0.1, 0.5), ComplexF64, E + gradient", 12, 20, np.complex128, H2, None))
    # C1 forward circuit alone (BASELINE configs[0])
    rng = np.random.default_rng(12020)
    c1 = qc.GenericBrickworkCircuit(12, 20)
    c1.gate_angles[...] = rng.uniform(0, 2 * np.pi, c1.gate_angles.shape)
    st = qc.B200State(12, np.complex128)
    l0 = qc.get_stat("launches_total")
    ms, ms_min = timeit(lambda: qc.apply_(st, st, c1), reps=20, warm=3)
    S = 16.0 * 2 ** 12
    out.append({"config": "C1: 12 q x 20 half-layers forward circuit (110 gates), ComplexF64, resident state", "ms": ms, "ms_min": ms_min,
                "kernel": "small-state circuit kernel" if qc.get_stat("passes") <= 1 else "tile_pass_kernel", "launches_per_call": (qc.get_stat("launches_total") - l0) / 23.0,
                "algorithmic_bytes": 2 * S, "achieved_gbs": 2 * S / (ms * 1e-3) / 1e9, "frac": 2 * S / (ms * 1e-3) / 1e9 / peak,
                "gate_apps_per_sec": 110 / (ms * 1e-3), "note": "a 64 KiB state: one cluster launch, the state stays in distributed shared memory; bound by per-gate latency (barriers, DSMEM reads), ms includes the host's table build and the launch", "_launches": qc.get_stat("launches_total") - l0})
    # expectation value at 30 qubits (one read of the state is the algorithmic traffic)
    for dtype, nm in ((np.complex128, "ComplexF64"), (np.complex64, "ComplexF32")):
        st = qc.B200State(30, dtype)
        cs = qc.GenericBrickworkCircuit(30, 2)
        cs.gate_angles[...] = rng.uniform(0, 2 * np.pi, cs.gate_angles.shape)
        qc.apply_(st, st, cs)
        l0 = qc.get_stat("launches_total")
        val = {}

        def fn():
            val["E"] = qc.measure_(None, H5, st)
        ms, ms_min = timeit(fn)
        S = float(np.dtype(dtype).itemsize) * 2.0 ** 30
        out.append({"config": f"measure(TFIM, psi) at 30 q, {nm} (dense product state)", "ms": ms, "ms_min": ms_min, "kernel": "flip-tile expectation kernel",
                    "passes": qc.get_stat("passes"), "launches_per_call": (qc.get_stat("launches_total") - l0) / 7.0, "algorithmic_bytes": S,
                    "achieved_gbs": S / (ms * 1e-3) / 1e9, "frac": S / (ms * 1e-3) / 1e9 / peak, "energy": val["E"],
                    "_launches": qc.get_stat("launches_total") - l0})
        del st
    return out


_OPTS = {}


def qc_opt(key, default):
    return _OPTS.get(key, default)


def main():
    args = parse_args()
    if args.tile_bits is not None:
        _OPTS["tile_bits"] = args.tile_bits
    if args.low_bits is not None:
        _OPTS["low_bits"] = args.low_bits
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
