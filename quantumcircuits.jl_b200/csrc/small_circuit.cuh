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

namespace qcb {

constexpr int kSmallThreads = 256;
constexpr int kSmallChunk = 32;   // gates per shared-memory chunk of matrices
constexpr int kSmallMaxGbits = 3; // cluster of at most 8 CTAs (the portable maximum)

QCB_HD float gate_re(const GateMat<float> &U, int r, int c) { return U.v[4 * (r + 4 * c)]; }
QCB_HD float gate_im(const GateMat<float> &U, int r, int c) { return U.v[4 * (r + 4 * c) + 3]; }

// host: how many index bits go to the CTA rank for an n-qubit state
inline int small_circuit_gbits(int n) { return n <= 9 ? 0 : (n - 9 > kSmallMaxGbits ? kSmallMaxGbits : n - 9); }
inline int small_circuit_max_qubits(int amp_bytes) { return amp_bytes == 16 ? 15 : 16; }
template <typename R> inline size_t small_circuit_smem(int n)
{
    const int L = n - small_circuit_gbits(n);
    return 2 * (sizeof(typename Vec2<R>::type) << L) + kSmallChunk * (sizeof(GateMat<R>) + sizeof(int));
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

// grid = one cluster of 2^gbits CTAs of kSmallThreads threads.  gates[g] acts on index bits (gate_k[g], gate_k[g] + 1).
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
    V *A = reinterpret_cast<V *>(qcb_small_raw), *B = A + nloc;
    GateMat<R> *gsm = reinterpret_cast<GateMat<R> *>(B + nloc);
    int *ksm = reinterpret_cast<int *>(gsm + kSmallChunk);

    for (unsigned i = tid; i < nloc; i += kSmallThreads) A[i] = src[((size_t)rank << L) + i];
    bool dirty = true; // this CTA wrote its buffer since the last cluster barrier
    for (int g0 = 0; g0 < ngates; g0 += kSmallChunk) {
        const int cnt = min(kSmallChunk, ngates - g0);
        __syncthreads(); // the previous chunk's matrices are no longer read (and the shard load / last gate is complete)
        {
            const uint4 *gsrc = reinterpret_cast<const uint4 *>(gates + g0);
            uint4 *gdst = reinterpret_cast<uint4 *>(gsm);
            const int words = cnt * (int)(sizeof(GateMat<R>) / sizeof(uint4));
            for (int i = (int)tid; i < words; i += kSmallThreads) gdst[i] = gsrc[i];
            if ((int)tid < cnt) ksm[tid] = gate_k[g0 + tid];
        }
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            const int k = ksm[j];
            const GateMat<R> &U = gsm[j];
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
                if (dirty) cluster.sync(); // every CTA's buffer is complete before anybody reads it remotely
                for (unsigned a = tid; a < nloc; a += kSmallThreads) {
                    unsigned owner[4], loc[4];
                    const int r = small_remote_item(rank, L, a, k, owner, loc);
                    V x[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const V *p = owner[c] == rank ? A : cluster.map_shared_rank(A, owner[c]);
                        x[c] = p[loc[c]];
                    }
                    B[a] = small_row<R, V>(U, r, x);
                }
                cluster.sync(); // all remote reads of A are done; B is complete everywhere
                V *t = A; A = B; B = t;
                dirty = false;
            }
        }
    }
    __syncthreads();
    for (unsigned i = tid; i < nloc; i += kSmallThreads) dst[((size_t)rank << L) + i] = A[i];
}
#endif

} // namespace qcb
