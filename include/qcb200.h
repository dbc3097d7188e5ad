/*
 * qcb200.h -- C ABI of libqcb200.so, the B200 (sm_100a) dense-statevector backend that
 * sits behind the QuantumCircuits.jl API (JamieMair/QuantumCircuits.jl v0.1.1).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / CUDA types.  Each
 * entry point names the reference interface it replaces (file:line relative to the
 * reference tree).  The Julia-side `ccall` bindings are in
 * quantumcircuits.jl_b200/julia/B200Backend.jl and INTEGRATION.md.
 *
 * Conventions (identical to the reference):
 *   - a state is 2^n complex amplitudes; qubit q (1-based, as in Julia) is bit q-1 of
 *     the 0-based linear index (column-major 2x2x...x2 tensor, src/utils.jl:61-66);
 *   - dtype QCB_C64 = ComplexF32 amplitudes, QCB_C128 = ComplexF64 amplitudes;
 *   - gate matrices are ALWAYS ComplexF64 (src/apply.jl:100-101), interleaved (re, im),
 *     column-major: 1-qubit u[out + 2*in] (4 complex), 2-qubit u[i + 2j + 4a + 8b]
 *     (16 complex; (i,a) <-> qubit K, (j,b) <-> qubit K+1, (i,j) output, (a,b) input),
 *     i.e. exactly the memory of the Julia arrays `mat(gate)`;
 *   - K is the 1-based `target_gate_dim`; a 2-qubit gate acts on K and K+1
 *     (src/gates.jl:126-131);
 *   - angle arrays are 15 x G Float64, column-major, column g = gate g in application
 *     order (src/circuits.jl:3-18).
 *
 * Every function returns 0 (QCB_OK) or an error code; qcb_last_error() gives the text.
 * Calls are stream-ordered on one stream per process (qcb_set_stream); functions that return
 * host data (download, measure, gradients) block until the data is valid.  One host
 * thread per process is assumed (like the reference: synchronize after every kernel,
 * src/apply.jl:139).  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with QCB_ECUDA.
 */
#ifndef QCB200_H
#define QCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCB_OK 0
#define QCB_EINVAL 1 /* bad argument: K out of range, size / dtype mismatch (src/apply.jl:106-115) */
#define QCB_ENOMEM 2
#define QCB_ECUDA 3
#define QCB_ENCCL 4
#define QCB_EUNSUPPORTED 5

#define QCB_C64 0  /* Julia ComplexF32 */
#define QCB_C128 1 /* Julia ComplexF64 */

#define QCB_HAM_TFIM 0 /* params (J, h, g)  src/hamiltonians/tfim.jl:11-15       */
#define QCB_HAM_EAST 1 /* params (c, A, B)  src/hamiltonians/east_model.jl:16-27 */

typedef struct qcb_state qcb_state; /* opaque device-resident state (replaces the CuArray in GPUApplyCache use) */

/* ---- library / device -------------------------------------------------------------- */
/* Replaces KernelAbstractions.get_backend / construct_apply_cache (src/apply.jl:95-104):
 * binds the calling process to one CUDA device. */
int qcb_init(int device);
int qcb_finalize(void);
int qcb_device_count(int *out);
const char *qcb_last_error(void);
const char *qcb_version(void);
/* KernelAbstractions.synchronize (src/apply.jl:139, hamiltonians.jl:39). */
int qcb_sync(void);
/* All work is issued on one stream: the CUDA default stream unless an externally owned cudaStream_t
 * (e.g. torch's current stream) is given here; NULL selects the default stream again. */
int qcb_set_stream(void *cuda_stream);
/* CUDA-event stopwatch on the library stream for hosts without their own event API. */
int qcb_timer_start(void);
int qcb_timer_stop(float *out_ms);
/* Tunables: "tile_bits", "low_bits", "max_gates_per_pass", "adjoint_tile_bits" (0 = auto), "adjoint_mma" (1 = ComplexF64 backward pass on the FP64
 * tensor cores, default; 0 = scalar-FMA pass), "adjoint_warps" (0 = auto, 4, 8: warps per CTA of that pass), "adjoint_prog_host" (1 = build that pass's
 * per-thread gate programs on the host and upload them: cross-check of the device-built table), "measure_ctas" (0 = auto, 2, 3: CTAs per SM of the tiled expectation-value kernel), "measure_persistent" (1 = persistent form of that kernel, 0 = one tile per CTA, default), "contract_coop" (1 = 16-lane cooperative
 * contraction of the environments with the gate derivatives, default; 0 = one thread per gate layer), "force_simple" (1 = one
 * unfused launch per gate). */
int qcb_set_option(const char *key, long value);
/* Counters of the most recent apply / gradient call: "passes", "stages", "launches", "gates",
 * and "launches_total" (all kernels launched by this library since qcb_init). */
int qcb_get_stat(const char *key, double *out);
/* Frees the library's cached work buffers (gradient work states, host-call staging state, reduction partials). */
int qcb_trim(void);

/* ---- states  (zero_state_tensor src/utils.jl:61-66; similar/copy/.= used at gradients.jl:74-76,90,118) */
int qcb_state_create(int nqubits, int dtype, qcb_state **out);
/* Wrap caller-owned device memory (2^nqubits amplitudes of dtype); the library never frees it. */
int qcb_state_wrap(int nqubits, int dtype, void *device_ptr, qcb_state **out);
int qcb_state_destroy(qcb_state *s);
int qcb_state_nqubits(const qcb_state *s, int *out);
int qcb_state_dtype(const qcb_state *s, int *out);
int qcb_state_device_ptr(const qcb_state *s, void **out);
int qcb_state_set_zero_state(qcb_state *s);                   /* |0...0>          */
int qcb_state_fill_zero(qcb_state *s);                        /* psi' .= 0 (apply.jl:135) */
int qcb_state_upload(qcb_state *s, const void *host);         /* Array -> device  */
int qcb_state_download(const qcb_state *s, void *host);       /* device -> Array  */
int qcb_state_copy(qcb_state *dst, const qcb_state *src);     /* psi'' .= psi     */
int qcb_state_norm2(const qcb_state *s, double *out);         /* sum |amp|^2      */

/* ---- gate algebra on the host (src/gates.jl:80-109, src/circuits.jl:9-12) -------- */
int qcb_build_general_unitary_gate(const double *angles15, double *out_u32);
int qcb_brickwork_num_gates(int nbits, int nlayers);
/* Partial derivative of the 15-angle gate with respect to angle k (0-based); the factor the
 * reference differentiates implicitly with its +-pi/2 shifts (src/gradients.jl:120-139). */
int qcb_general_unitary_gate_derivative(const double *angles15, int k, double *out_u32);
/* out15[k] = 2 Re sum_{r,c} env[r,c] (dU/dangle_k)[r,c], k = 0..14, for a 4x4 environment env (column-major r + 4c, (re, im)
 * pairs): the contraction the adjoint gradient applies to each gate's environment, all 15 angles from shared partial products. */
int qcb_general_unitary_gate_gradient(const double *angles15, const double *env32, double *out15);

/* ---- gate application (in place; the shim keeps the (psi', psi) signature) -------- */
/* apply!(psi', psi, ::Localised1SpinGate)  src/apply.jl:20-38 */
int qcb_apply_1q(qcb_state *s, const double *u8, int K);
/* apply!(cache, psi', psi, ::Localised2SpinAdjGate)  src/apply.jl:117-142, kernel :62-84 */
int qcb_apply_2q(qcb_state *s, const double *u32, int K);
/* apply(psi, gates::AbstractArray)  src/apply.jl:5-18 : mixed list, fused on chip.
 * arity[g] in {1,2}; mats holds 32 doubles per gate (a 1-qubit gate uses the first 8). */
int qcb_apply_gates(qcb_state *s, int ngates, const int *K, const int *arity, const double *mats);
/* The same with distinct states: dst = gates(src), src untouched -- apply!(psi', psi, gate) src/apply.jl:20,39 and
 * apply!(cache, psi', psi, gate) :117,:121 without the copy / zero fill (:135) in front of the pass. */
int qcb_apply_gates_from(qcb_state *dst, const qcb_state *src, int ngates, const int *K, const int *arity, const double *mats);
/* apply!(cache, psi', psi, circuit::GenericBrickworkCircuit)  src/circuits.jl:19-31.
 * first_gate/last_gate (0-based, half-open) select a slice of the circuit: (0, G) is the
 * whole circuit; propagate_forwards! (src/gradients.jl:54-71) is (gate_idx-1, G). */
int qcb_apply_brickwork(qcb_state *s, int nbits, int nlayers, const double *angles);
int qcb_apply_brickwork_range(qcb_state *s, int nbits, int nlayers, const double *angles, int first_gate,
                              int last_gate, int adjoint /* 1: apply U_g^dagger for g = last-1 .. first */);
/* The same with distinct states, dst = circuit(src) and src untouched (the reference's (psi', psi) contract): the first
 * fused pass reads src and writes dst, no copy of the state precedes it.  Unsharded states of equal size and type. */
int qcb_apply_brickwork_from(qcb_state *dst, const qcb_state *src, int nbits, int nlayers, const double *angles, int first_gate,
                             int last_gate, int adjoint);
/* One-shot form with HOST buffers (apply(psi, circuit) src/circuits.jl:37-42 on an Array):
 * H2D, circuit, D2H.  psi_in may be NULL for |0...0>. */
int qcb_apply_brickwork_host(int dtype, int nbits, int nlayers, const double *angles, const void *psi_in_host,
                             void *psi_out_host);

/* ---- expectation values: measure!(psi', H, psi)  src/hamiltonians/hamiltonians.jl:31-50 ---- */
/* out_E = (re, im) of sum(psi') so the shim can reproduce the imaginary-energy warning (:43-45).
 * Sharded states: collective call (every rank), every rank receives the full expectation value. */
int qcb_measure_tfim(const qcb_state *s, double J, double h, double g, double *out_E2); /* tfim.jl:17-54 */
int qcb_measure_east(const qcb_state *s, double c, double A, double B, double *out_E2); /* east_model.jl:30-52 */

/* measure(H::AbstractMatrix, psi) = real(dot(psi, H, psi))  src/circuits.jl:49-55 (also measure!(psi', H, psi) :56-58 and, after
 * qcb_apply_brickwork, measure(H, psi, circuit) :59-62).  H is 2^n x 2^n in CSR form on the host: rowptr[nrows+1], colidx[nnz]
 * (0-based, 64-bit), vals = nnz (re, im) pairs; a dense matrix is passed as the CSR holding every entry.  out_E2 = (re, im);
 * the reference asserts im < 1e-12.  Not for sharded states. */
int qcb_measure_csr(const qcb_state *s, long long nrows, const long long *rowptr, const long long *colidx, const double *vals_reim,
                    double *out_E2);

/* ---- gradients!(cache, psi', psi'', psi, H, psi0, circuit; calculate_energy)  src/gradients.jl:88-152 ----
 * Same numbers as the reference's parameter-shift sweep, computed by the O(G) adjoint sweep
 * (SURVEY.md appendix A).  psi0 is not modified.  out_grads is 15 x G column-major.
 * Sharded states: collective call (every rank), every rank receives the energy and the full gradient. */
int qcb_gradients_brickwork(const qcb_state *psi0, int nbits, int nlayers, const double *angles, int ham_kind,
                            const double *ham_params3, int want_energy, double *out_E, double *out_grads);
/* gradients(H::AbstractMatrix, psi0, circuit; calculate_energy): the same sweep for an explicit Hermitian Hamiltonian in CSR
 * form (layout as qcb_measure_csr) -- the reference's generic gradients! over measure!(psi', H::AbstractMatrix, psi)
 * (src/circuits.jl:56-58; examples/test_gradients.jl:25).  Not for sharded states. */
int qcb_gradients_brickwork_csr(const qcb_state *psi0, int nbits, int nlayers, const double *angles, long long nrows, const long long *rowptr,
                                const long long *colidx, const double *vals_reim, int want_energy, double *out_E, double *out_grads);
/* optimise!(circuit, H, psi0, epochs, lr)  src/gradients.jl:1-17 with quirk Q1 resolved
 * (energy + full gradient each epoch).  energies_out has epochs+1 entries laid out like the reference's. */
int qcb_optimise_brickwork(const qcb_state *psi0, int nbits, int nlayers, double *angles_inout, int ham_kind,
                           const double *ham_params3, int epochs, double lr, double *energies_out);

/* ---- building blocks of the statevector polar optimiser (ext/MPSExt/polar_optimise.jl:83-102) ----------------
 * qcb_state_inner: out = (re, im) of <a|b> = sum conj(a) b  (the overlap at polar_optimise.jl:140).
 * qcb_environment_2q: out_env32 = 4x4 complex, column-major E[r + 4c] = sum_rest conj(lam[r, rest]) psi[c, rest] for the
 * qubit pair (K, K+1) -- the contraction of upper and lower state over all other qubits (polar_optimise.jl:88-90). */
int qcb_state_inner(const qcb_state *a, const qcb_state *b, double *out2);
int qcb_environment_2q(const qcb_state *lam, const qcb_state *psi, int K, double *out_env32);

/* ---- multi-GPU: one process per GPU, amplitudes sharded by the top log2(nranks) qubits ---------
 * No reference counterpart (the reference holds one array in one memory).  Bootstrap: rank 0 calls
 * qcb_dist_unique_id and ships the 128 bytes to the other ranks by any host channel
 * (torch.distributed, MPI, a file); then every rank calls qcb_dist_init. */
int qcb_dist_unique_id(void *out_128_bytes);
int qcb_dist_init(int rank, int nranks, const void *unique_id_128_bytes);
int qcb_dist_finalize(void);
/* A sharded state holds 2^(nqubits - log2 nranks) amplitudes per rank. */
int qcb_state_create_sharded(int nqubits, int dtype, qcb_state **out);
int qcb_state_local_qubits(const qcb_state *s, int *out);
/* Optional: let the ranks read each other's shards directly (CUDA IPC over NVLink).  Every rank exports 128 bytes, the
 * host gathers them in rank order and every rank imports the whole table.  With peers mapped, a qubit swap is no longer
 * a separate all-to-all: it is fused into the next gate pass, whose tile loads come straight from the peers' memory.
 * Option "fuse_swaps" (default 1) switches the fusion off without unmapping.  When no rank could allocate a second buffer
 * (34 qubits on 2 GPUs) only the first buffers are mapped and a qubit swap is an in-place exchange over peer memory (one kernel,
 * no staging; option "peer_swap", default 1); when some ranks have a second buffer and others do not, nothing is mapped. */
int qcb_state_ipc_export(const qcb_state *s, void *out_128_bytes);
int qcb_state_ipc_import(qcb_state *s, const void *all_handles);
int qcb_state_ipc_close(qcb_state *s); /* unmap the peers' buffers; then synchronise the ranks before qcb_state_destroy */
/* Upload / download of this rank's canonical shard (amplitudes [rank * 2^nloc, (rank+1) * 2^nloc)). */
int qcb_state_upload_shard(qcb_state *s, const void *host_shard);
int qcb_state_download_shard(qcb_state *s, void *host_shard);

#ifdef __cplusplus
}
#endif
#endif /* QCB200_H */
