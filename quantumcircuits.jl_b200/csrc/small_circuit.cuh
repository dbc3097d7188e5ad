// small_circuit.cuh -- a whole gate list on a small state in ONE launch: the state lives in the shared memory of one
// thread-block cluster (up to 8 CTAs, distributed shared memory) for the entire circuit.
//
// Replaces, for states of up to 15 (ComplexF64) / 16 (ComplexF32) qubits, the sequence of fused tile passes of tile.cuh
// (a 12-qubit state is 64 KiB: two CTAs of the tile kernel, one launch per pass, every pass a round trip through L2), i.e.
// the reference's per-gate launches of _apply_2spin! (src/apply.jl:62-84, 134-142) for the 12-qubit configuration of
// BASELINE.json.  SURVEY.md section 7 "tiny states".
//
// Layout: CTA r of the cluster holds the amplitudes whose top `gbits` index bits are r (2^L of them, L = n - gbits), in two
// shared-memory buffers.  A gate on index bits (k, k+1):
//   * k + 1 <  L ("local"):  one thread per amplitude quad, in place, CTA barrier.
//   * k + 1 >= L ("remote"): owner computes -- every thread produces output amplitudes of its own CTA, one row of the 4x4
//     product each, from the four inputs of the quad, of which two or three sit in other CTAs and are read through
//     distributed shared memory (cluster.map_shared_rank); results go to the second buffer, one cluster barrier, and the
//     buffers swap roles.  No arithmetic is done twice: a row of the quad is computed by the CTA that owns it.
// The gate matrices are streamed from global memory into shared memory in chunks of kSmallChunk gates.
#pragma once
#if defined(__CUDACC__)
#include <cooperative_groups.h>
#endif
#include "tile.cuh"
#include "kernels.cuh"

namespace qcb {

#ifndef QCB_SMALL_THREADS
#define QCB_SMALL_THREADS 256
#endif
constexpr int kSmallThreads = QCB_SMALL_THREADS;
constexpr int kSmallChunk = 32;   // gates per shared-memory chunk of matrices
constexpr int kSmallMaxGbits = 3; // cluster of at most 8 CTAs (the portable maximum)

QCB_HD float gate_re(const GateMat<float> &U, int r, int c) { return U.v[4 * (r + 4 * c)]; }
QCB_HD float gate_im(const GateMat<float> &U, int r, int c) { return U.v[4 * (r + 4 * c) + 3]; }

// host: how many index bits go to the CTA rank for an n-qubit state
inline int small_circuit_gbits(int n) { return n <= 9 ? 0 : (n - 9 > kSmallMaxGbits ? kSmallMaxGbits : n - 9); }
inline int small_circuit_max_qubits(int amp_bytes) { return amp_bytes == 16 ? 15 : 16; }
// (a single CTA never takes the remote path: no second buffer)
template <typename R> inline size_t small_circuit_smem(int n, int gbits)
{
    const int L = n - gbits;
    return (gbits ? 2 : 1) * (sizeof(typename Vec2<R>::type) << L) + kSmallChunk * (sizeof(GateMat<R>) + sizeof(int));
}
// the gradient kernel: per-warp sums (two gates) + two buffers each for psi and lambda + the matrix chunk; at most 12 qubits
constexpr int kSmallGradMaxQubits = 12;
template <typename R> inline size_t small_gradient_smem(int n, int gbits)
{
    const int L = n - gbits;
    return 2 * (kSmallThreads / 32) * 32 * sizeof(double) + (gbits ? 4 : 2) * (sizeof(typename Vec2<R>::type) << L) + kSmallChunk * (sizeof(GateMat<R>) + sizeof(int));
}

// index algebra shared with the CPU replay (tests/emu)
// local gate: first amplitude of quad q for a gate on bits (k, k+1)
QCB_HD unsigned small_quad_base(unsigned q, int k) { return ((q >> k) << (k + 2)) | (q & ((1u << k) - 1u)); }
// remote gate: the amplitude `a` of CTA `rank` is row r of its quad; input c of that quad sits at local index loc[c] of CTA owner[c]
QCB_HD int small_remote_item(unsigned rank, int L, unsigned a, int k, unsigned (&owner)[4], unsigned (&loc)[4])
{
    const unsigned i = (rank << L) | a, ib = i & ~(3u << k);
    for (int c = 0; c < 4; c++) {
        const unsigned idx = ib | ((unsigned)c << k);
        owner[c] = idx >> L;
        loc[c] = idx & ((1u << L) - 1u);
    }
    return (int)((i >> k) & 3u);
}

// one output row of a gate: out = sum_c U[r][c] x_c
template <typename R, typename V> QCB_HD V small_row(const GateMat<R> &U, int r, const V (&x)[4])
{
    R re = 0, im = 0;
    for (int c = 0; c < 4; c++) {
        const R ur = gate_re(U, r, c), ui = gate_im(U, r, c);
        re = fma(ur, x[c].x, re);
        im = fma(ur, x[c].y, im);
        re = fma(-ui, x[c].y, re);
        im = fma(ui, x[c].x, im);
    }
    V o;
    o.x = re;
    o.y = im;
    return o;
}

#if defined(__CUDACC__)
namespace cg = cooperative_groups;

// shared-memory state of one CTA: the current buffer A, the spare B, and whether this CTA wrote A since the last cluster barrier
template <typename V> struct SmallBuf {
    V *A, *B;
};

// stage a chunk of the gate table in shared memory (all threads; barriers around it)
template <typename R>
__device__ __forceinline__ void small_load_chunk(GateMat<R> *gsm, int *ksm, const GateMat<R> *gates, const int *gate_k, int g0, int cnt, unsigned tid)
{
    __syncthreads(); // the previous chunk's matrices are no longer read
    const uint4 *gsrc = reinterpret_cast<const uint4 *>(gates + g0);
    uint4 *gdst = reinterpret_cast<uint4 *>(gsm);
    const int words = cnt * (int)(sizeof(GateMat<R>) / sizeof(uint4));
    for (int i = (int)tid; i < words; i += kSmallThreads) gdst[i] = gsrc[i];
    if ((int)tid < cnt) ksm[tid] = gate_k[g0 + tid];
    __syncthreads();
}

// grid = one cluster of 2^gbits CTAs of kSmallThreads threads.  gates[g] acts on index bits (gate_k[g], gate_k[g] + 1).
template <typename R>
__device__ __forceinline__ void small_apply_gates(cg::cluster_group &cluster, SmallBuf<typename Vec2<R>::type> &S, bool &dirty, GateMat<R> *gsm, int *ksm,
                                                  const GateMat<R> *__restrict__ gates, const int *__restrict__ gate_k, int ngates, unsigned rank, int L)
{
    using V = typename Vec2<R>::type;
    const unsigned nloc = 1u << L, tid = threadIdx.x;
    for (int g0 = 0; g0 < ngates; g0 += kSmallChunk) {
        const int cnt = min(kSmallChunk, ngates - g0);
        small_load_chunk<R>(gsm, ksm, gates, gate_k, g0, cnt, tid);
        for (int j = 0; j < cnt; j++) {
            const int k = ksm[j];
            const GateMat<R> &U = gsm[j];
            V *A = S.A, *B = S.B;
            if (k + 1 < L) {
                for (unsigned q = tid; q < (nloc >> 2); q += kSmallThreads) {
                    const unsigned base = small_quad_base(q, k);
                    V x0 = A[base], x1 = A[base | (1u << k)], x2 = A[base | (2u << k)], x3 = A[base | (3u << k)];
                    apply_quad(x0, x1, x2, x3, U);
                    A[base] = x0; A[base | (1u << k)] = x1; A[base | (2u << k)] = x2; A[base | (3u << k)] = x3;
                }
                __syncthreads();
                dirty = true;
            } else {
                // One cluster barrier per gate: every CTA's buffer A is complete before anybody reads it remotely.  The results
                // go to B, which nobody reads remotely before the NEXT cluster barrier -- and A is not written again before
                // that barrier either (gates in between work on B in place), by which time every CTA has finished reading it.
                cluster.sync();
                for (unsigned a0 = tid; a0 < nloc; a0 += 2 * kSmallThreads) {
                    // two amplitudes per thread and iteration: all remote loads in flight before the first use
                    const unsigned a1 = a0 + kSmallThreads;
                    const bool two = a1 < nloc;
                    unsigned owner0[4], loc0[4], owner1[4], loc1[4];
                    const int r0 = small_remote_item(rank, L, a0, k, owner0, loc0);
                    const int r1 = small_remote_item(rank, L, two ? a1 : a0, k, owner1, loc1);
                    V x0[4], x1[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const V *p0 = owner0[c] == rank ? A : cluster.map_shared_rank(A, owner0[c]);
                        const V *p1 = owner1[c] == rank ? A : cluster.map_shared_rank(A, owner1[c]);
                        x0[c] = p0[loc0[c]];
                        x1[c] = p1[loc1[c]];
                    }
                    B[a0] = small_row<R, V>(U, r0, x0);
                    if (two) B[a1] = small_row<R, V>(U, r1, x1);
                }
                __syncthreads(); // B is complete in this CTA (local gates may follow)
                S.A = B; S.B = A;
                dirty = true;
            }
        }
    }
}

template <typename R>
__global__ void __launch_bounds__(kSmallThreads, 1)
cluster_circuit_kernel(const typename Vec2<R>::type *__restrict__ src, typename Vec2<R>::type *__restrict__ dst,
                       const GateMat<R> *__restrict__ gates, const int *__restrict__ gate_k, int n, int gbits, int ngates)
{
    using V = typename Vec2<R>::type;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = gbits ? cluster.block_rank() : 0u;
    const int L = n - gbits;
    const unsigned nloc = 1u << L, tid = threadIdx.x;
    extern __shared__ __align__(16) unsigned char qcb_small_raw[];
    SmallBuf<V> S;
    S.A = reinterpret_cast<V *>(qcb_small_raw);
    S.B = gbits ? S.A + nloc : S.A;
    GateMat<R> *gsm = reinterpret_cast<GateMat<R> *>(S.B + nloc);
    int *ksm = reinterpret_cast<int *>(gsm + kSmallChunk);
    for (unsigned i = tid; i < nloc; i += kSmallThreads) S.A[i] = src[((size_t)rank << L) + i];
    bool dirty = true; // this CTA wrote its buffer since the last cluster barrier
    small_apply_gates<R>(cluster, S, dirty, gsm, ksm, gates, gate_k, ngates, rank, L);
    __syncthreads();
    for (unsigned i = tid; i < nloc; i += kSmallThreads) dst[((size_t)rank << L) + i] = S.A[i];
    if (gbits) cluster.sync(); // no CTA leaves while another one may still read its buffers
}

// ---- energy + all environments of a brickwork circuit in the same launch (the adjoint sweep of src/gradients.jl:88-152 as in
// api.cu: gradients_impl, for a state that fits the cluster) ------------------------------------------------------------------
// warp transpose-reduction: returns, in lane v, the sum over the 32 lanes of acc[v] (31 64-bit shuffles and 31 additions per
// lane instead of 160 for 32 butterfly reductions).  acc is destroyed.
__device__ __forceinline__ double small_warp_reduce32(double (&acc)[32], unsigned lane)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) { // s = half of the values still held
        const bool up = (lane & (unsigned)s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const double send = up ? acc[i] : acc[i + s], keep = up ? acc[i + s] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return acc[0];
}
constexpr int kSmallWarps = kSmallThreads / 32;
// cross-warp part of the sum, one barrier behind: part[buf][warp][v] was written before the last CTA barrier; warp 0 adds
// the kSmallWarps rows in fixed order and stores the 32 sums (reproducible)
__device__ __forceinline__ void small_flush_part(const double *part_buf, double *out, unsigned tid)
{
    if (tid < 32) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kSmallWarps; w++) s += part_buf[w * 32 + tid];
        out[tid] = s;
    }
}

// M[2(r + 4c)], M[2(r + 4c) + 1] += conj(l_r) x_c for row r of a quad (both states BEFORE the gate is un-applied)
template <typename V> __device__ __forceinline__ void small_outer_row(double (&acc)[32], int r, const V &l, const V (&x)[4])
{
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
        if (rr == r) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                acc[2 * (rr + 4 * c)] += (double)l.x * (double)x[c].x + (double)l.y * (double)x[c].y;
                acc[2 * (rr + 4 * c) + 1] += (double)l.x * (double)x[c].y - (double)l.y * (double)x[c].x;
            }
        }
}

// gates: U_g (forward), gates_dag: U_g^dagger, both in application order g = 0..G-1.  Outputs (global): Mpart[rank][g][32] =
// this CTA's part of M_g = sum_quads conj(lambda_g) psi_g^T (states after gate g), Epart[rank][32]: [0], [1] = its part of <psi|H|psi>.
template <typename R>
__global__ void __launch_bounds__(kSmallThreads, 1)
cluster_gradient_kernel(const typename Vec2<R>::type *__restrict__ src, const GateMat<R> *__restrict__ gates, const GateMat<R> *__restrict__ gates_dag,
                        const int *__restrict__ gate_k, int n, int gbits, int ngates, HamEval H, double *__restrict__ Mpart, double *__restrict__ Epart)
{
    using V = typename Vec2<R>::type;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = gbits ? cluster.block_rank() : 0u;
    const int L = n - gbits;
    const unsigned nloc = 1u << L, tid = threadIdx.x;
    extern __shared__ __align__(16) unsigned char qcb_small_raw[];
    double *part = reinterpret_cast<double *>(qcb_small_raw); // [2][kSmallWarps][32]: per-warp sums of the current / previous gate
    SmallBuf<V> P, Q;                                          // psi, lambda
    P.A = reinterpret_cast<V *>(part + 2 * kSmallWarps * 32);
    P.B = gbits ? P.A + nloc : P.A;
    Q.A = P.B + nloc;
    Q.B = gbits ? Q.A + nloc : Q.A;
    GateMat<R> *gsm = reinterpret_cast<GateMat<R> *>(Q.B + nloc);
    int *ksm = reinterpret_cast<int *>(gsm + kSmallChunk);
    for (unsigned i = tid; i < nloc; i += kSmallThreads) P.A[i] = src[((size_t)rank << L) + i];
    bool dirty = true;
#ifdef QCB_SMALL_TIMING
    long long tq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq0 = clock64();
#define QCB_TICK(i) { const long long now = clock64(); tq[i] += now - tq0; tq0 = now; }
#else
#define QCB_TICK(i)
#endif
    // forward circuit
    small_apply_gates<R>(cluster, P, dirty, gsm, ksm, gates, gate_k, ngates, rank, L);
    QCB_TICK(0)
    // lambda = H psi, E = <psi|lambda>   (kernels.cuh: k_apply_ham)
    if (dirty) cluster.sync();
    {
        const double oc = H.offdiag_coef();
        double e[32];
#pragma unroll
        for (int v = 0; v < 32; v++) e[v] = 0.0;
        for (unsigned a = tid; a < nloc; a += kSmallThreads) {
            const unsigned long long C = ((unsigned long long)rank << L) | a;
            const V p = P.A[a];
            double xr = 0, xi = 0;
            for (int i = 0; i < n; i++) {
                if (!H.flip_on(C, i)) continue;
                const unsigned idx = (unsigned)C ^ (1u << i), owner = idx >> L;
                const V *q = owner == rank ? P.A : cluster.map_shared_rank(P.A, owner);
                const V o = q[idx & (nloc - 1u)];
                xr += (double)o.x;
                xi += (double)o.y;
            }
            const double d = H.diag(C);
            const double lr = d * (double)p.x + oc * xr, li = d * (double)p.y + oc * xi;
            V lam;
            lam.x = (R)lr;
            lam.y = (R)li;
            Q.A[a] = lam;
            e[0] += (double)p.x * lr + (double)p.y * li;
            e[1] += (double)p.x * li - (double)p.y * lr;
        }
        const double es = small_warp_reduce32(e, tid & 31u);
        part[(tid >> 5) * 32 + (tid & 31u)] = es; // (only the first two of the 32 sums are used)
    }
    cluster.sync(); // lambda is complete everywhere and nobody reads psi remotely any more
    dirty = false;
    QCB_TICK(1)
    small_flush_part(part, Epart + 32 * rank, tid);
    unsigned cur = 1;      // part buffer of the gate being processed
    double *pending = nullptr; // where the sums in part[cur ^ 1] go (the previous gate's), once a barrier has passed
    // backward sweep
    for (int g1 = ngates; g1 > 0; g1 -= kSmallChunk) {
        const int g0 = max(0, g1 - kSmallChunk), cnt = g1 - g0;
        small_load_chunk<R>(gsm, ksm, gates_dag, gate_k, g0, cnt, tid);
        for (int j = cnt - 1; j >= 0; j--) {
            const int k = ksm[j];
            const GateMat<R> &Ud = gsm[j];
            QCB_TICK(2)
            double *prow = part + (cur * kSmallWarps + (tid >> 5)) * 32; // this warp's sums for this gate
            if (k + 1 < L) {
                // Local gate.  The outer products M = sum_quads conj(lambda) psi^T over the warp's quads run on the FP64 tensor
                // cores with operands fetched from shared memory (states BEFORE the un-application): A[m][q] = real component m of
                // lambda of quad q, B[q][n] = component n of psi of quad q (mma.sync m8n8k4, as in adjoint.cuh) -- two accumulators
                // per lane and one lane exchange instead of 64 DFMA and a 31-step shuffle reduction per thread.
                const unsigned lane = tid & 31u, g = lane >> 2, t4 = lane & 3u, jj = g >> 1, pt = g & 1u, nq = nloc >> 2;
                const R *pL = reinterpret_cast<const R *>(Q.A) + pt, *pP = reinterpret_cast<const R *>(P.A) + pt;
                double d0[2] = {0.0, 0.0}, d1[2] = {0.0, 0.0};
                for (unsigned q0 = tid & ~31u; q0 < nq; q0 += kSmallThreads) { // (warp-uniform bounds: mma.sync needs the whole warp)
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        const unsigned quad = q0 + 4u * (unsigned)m + t4;
                        const unsigned idx = small_quad_base(quad < nq ? quad : 0u, k) | (jj << k);
                        const double a = quad < nq ? (double)pL[2 * idx] : 0.0, b = quad < nq ? (double)pP[2 * idx] : 0.0;
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(d0[m & 1]), "+d"(d1[m & 1]) : "d"(a), "d"(b));
                    }
                    __syncwarp(); // every operand of the warp's quads is read before a lane rewrites its quad
                    const unsigned q = q0 + lane;
                    if (q < nq) {
                        const unsigned base = small_quad_base(q, k);
                        V x[4], l[4];
#pragma unroll
                        for (int c = 0; c < 4; c++) { x[c] = P.A[base | ((unsigned)c << k)]; l[c] = Q.A[base | ((unsigned)c << k)]; }
                        apply_quad(x[0], x[1], x[2], x[3], Ud);
                        apply_quad(l[0], l[1], l[2], l[3], Ud);
#pragma unroll
                        for (int c = 0; c < 4; c++) { P.A[base | ((unsigned)c << k)] = x[c]; Q.A[base | ((unsigned)c << k)] = l[c]; }
                    }
                    __syncwarp();
                }
                const double s0 = d0[0] + d0[1], s1 = d1[0] + d1[1];
                const double o0 = __shfl_xor_sync(0xffffffffu, s0, 4), o1 = __shfl_xor_sync(0xffffffffu, s1, 4);
                if (!pt) { // the lane of Re(lambda_jj) assembles M[jj][t4]
                    const unsigned kk = jj + 4u * t4;
                    prow[2 * kk] = s0 + o1;
                    prow[2 * kk + 1] = s1 - o0;
                }
                dirty = true;
                QCB_TICK(3)
            } else {
                double acc[32];
#pragma unroll
                for (int v = 0; v < 32; v++) acc[v] = 0.0;
                cluster.sync(); // (one barrier per gate, as in small_apply_gates)
                QCB_TICK(4)
                for (unsigned a = tid; a < nloc; a += kSmallThreads) {
                    unsigned owner[4], loc[4];
                    const int r = small_remote_item(rank, L, a, k, owner, loc);
                    V x[4], l[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const V *pp = owner[c] == rank ? P.A : cluster.map_shared_rank(P.A, owner[c]);
                        const V *qq = owner[c] == rank ? Q.A : cluster.map_shared_rank(Q.A, owner[c]);
                        x[c] = pp[loc[c]];
                        l[c] = qq[loc[c]];
                    }
                    small_outer_row(acc, r, Q.A[a], x);
                    P.B[a] = small_row<R, V>(Ud, r, x);
                    Q.B[a] = small_row<R, V>(Ud, r, l);
                }
                V *t = P.A; P.A = P.B; P.B = t;
                t = Q.A; Q.A = Q.B; Q.B = t;
                dirty = true;
                QCB_TICK(5)
                // per-warp sums by a transpose-reduction (31 shuffles per lane); across the warps one barrier later
                prow[tid & 31u] = small_warp_reduce32(acc, tid & 31u);
            }
            if (pending) small_flush_part(part + (cur ^ 1u) * kSmallWarps * 32, pending, tid);
            __syncthreads(); // the gate's writes to the state, its part[] rows; the flush has read the other buffer
            pending = Mpart + ((size_t)rank * ngates + (g0 + j)) * 32;
            cur ^= 1u;
            QCB_TICK(6)
        }
    }
    if (pending) small_flush_part(part + (cur ^ 1u) * kSmallWarps * 32, pending, tid);
    if (gbits) cluster.sync(); // no CTA leaves while another one may still read its buffers
#ifdef QCB_SMALL_TIMING
    if (tid == 0 && rank == 0)
        printf("cycles: forward %lld  Hpsi %lld  loop-top %lld  bwd-local %lld  bwd-remote-sync %lld  bwd-remote %lld  reduce+barrier %lld\n", tq[0], tq[1], tq[2], tq[3], tq[4], tq[5], tq[6]);
#endif
}

// env_g = (sum over the CTAs of M_g) conj(U_g)  -- the environment in the convention of k_adjoint_step (kernels.cuh): psi after,
// lambda before the un-application; valid because U is unitary (psi_g = U psi_{g-1}).  One warp per gate; also sums the energy.
template <typename R>
__global__ void k_small_env_finish(const double *__restrict__ Mpart, const double *__restrict__ Epart, const GateMat<R> *__restrict__ gates, int ngates,
                                   int nranks, double *__restrict__ env, double *__restrict__ energy2)
{
    const int g = blockIdx.x, lane = threadIdx.x;
    __shared__ double M[32];
    double m = 0;
    for (int rk = 0; rk < nranks; rk++) m += Mpart[((size_t)rk * ngates + g) * 32 + lane];
    M[lane] = m;
    __syncwarp();
    if (lane < 16) {
        const int r = lane & 3, d = lane >> 2;
        double er = 0, ei = 0;
        for (int c = 0; c < 4; c++) {
            const double mr = M[2 * (r + 4 * c)], mi = M[2 * (r + 4 * c) + 1];
            const double ur = (double)gate_re(gates[g], c, d), ui = -(double)gate_im(gates[g], c, d);
            er += mr * ur - mi * ui;
            ei += mr * ui + mi * ur;
        }
        env[(size_t)g * 32 + 2 * (r + 4 * d)] = er;
        env[(size_t)g * 32 + 2 * (r + 4 * d) + 1] = ei;
    }
    if (g == 0 && lane < 2) {
        double e = 0;
        for (int rk = 0; rk < nranks; rk++) e += Epart[32 * rk + lane];
        energy2[lane] = e;
    }
}
#endif

} // namespace qcb
