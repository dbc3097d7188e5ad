// kernels.cuh -- the non-tiled device kernels of the backend:
//   * single-gate in-place kernels (small states, `force_simple` cross-check path),
//   * fused sweep-and-reduce expectation values for the kernel Hamiltonians
//     (replaces _tfim_measure! src/hamiltonians/tfim.jl:17-54 and _eastmodel_measure!
//     src/hamiltonians/east_model.jl:30-52 *plus* the separate sum(psi') pass of
//     src/hamiltonians/hamiltonians.jl:41 -- no scratch state is written),
//   * lambda = H psi and the 4x4 environment reduction of the adjoint gradient sweep
//     (replaces the O(G^2) shift sweep of src/gradients.jl:88-152).
// Reductions are deterministic: warp shuffles -> one partial per CTA -> fixed-order final sum.
#pragma once
#include "tile.cuh"

namespace qcb {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sums NV doubles per thread over a CTA of NWARPS warps; result in out[0..NV) (valid after the call for all threads)
template <int NV, int NWARPS> __device__ __forceinline__ void block_sum_n(double (&v)[NV], double *out)
{
    __shared__ double red[NWARPS][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const double s = warp_sum(v[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; w++) s += red[w][threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}
template <int NV> __device__ __forceinline__ void block_sum(double (&v)[NV], double *out) { block_sum_n<NV, kRedThreads / 32>(v, out); }

// ---- in-place single gates --------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) k_apply_2q_simple(typename Vec2<R>::type *psi, unsigned long long nquads, int kb,
                                                         const __grid_constant__ GateMat<R> U)
{
    using V = typename Vec2<R>::type;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long low = (1ull << kb) - 1ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nquads; i += stride) {
        const unsigned long long base = ((i >> kb) << (kb + 2)) | (i & low);
        V x0 = psi[base], x1 = psi[base | (1ull << kb)], x2 = psi[base | (2ull << kb)], x3 = psi[base | (3ull << kb)];
        apply_quad(x0, x1, x2, x3, U);
        psi[base] = x0; psi[base | (1ull << kb)] = x1; psi[base | (2ull << kb)] = x2; psi[base | (3ull << kb)] = x3;
    }
}

// 1-qubit gate on bit kb: u[out + 2 in] as 8 reals
template <typename R> struct Gate1 { R v[8]; };
template <typename R>
__global__ void __launch_bounds__(256) k_apply_1q_simple(typename Vec2<R>::type *psi, unsigned long long npairs, int kb,
                                                         const __grid_constant__ Gate1<R> U)
{
    using V = typename Vec2<R>::type;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long low = (1ull << kb) - 1ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += stride) {
        const unsigned long long base = ((i >> kb) << (kb + 1)) | (i & low);
        const V a = psi[base], b = psi[base | (1ull << kb)];
        V o0, o1;
        o0.x = U.v[0] * a.x - U.v[1] * a.y + U.v[4] * b.x - U.v[5] * b.y;
        o0.y = U.v[0] * a.y + U.v[1] * a.x + U.v[4] * b.y + U.v[5] * b.x;
        o1.x = U.v[2] * a.x - U.v[3] * a.y + U.v[6] * b.x - U.v[7] * b.y;
        o1.y = U.v[2] * a.y + U.v[3] * a.x + U.v[6] * b.y + U.v[7] * b.x;
        psi[base] = o0; psi[base | (1ull << kb)] = o1;
    }
}

template <typename R>
__global__ void __launch_bounds__(256) k_set_zero_state(typename Vec2<R>::type *psi, unsigned long long n, int set_one)
{
    using V = typename Vec2<R>::type;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        V z; z.x = (i == 0 && set_one) ? (R)1 : (R)0; z.y = (R)0;
        psi[i] = z;
    }
}

// ---- Hamiltonian diagonal / off-diagonal pieces -----------------------------------
// C is the full logical configuration.  diag(C): coefficient of |psi[C]|^2 ; the off-diagonal
// part is coef * sum_i w_i(C) psi[C ^ 2^i].
QCB_HD int popc64(unsigned long long x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

struct HamEval {
    int kind, n;          // 0 TFIM (p0 = J, p1 = h, p2 = g); 1 East (p0 = c, p1 = A, p2 = B)
    double p0, p1, p2;
    QCB_HD double diag(unsigned long long C) const
    {
        if (kind == 0) { // -J (zz + h z)        tfim.jl:25-32,48-52
            const unsigned long long pairmask = (n >= 2) ? ((1ull << (n - 1)) - 1ull) : 0ull;
            const int eq = popc64(~(C ^ (C >> 1)) & pairmask);
            const int zz = 2 * eq - (n - 1);
            const int z = n - 2 * popc64(C);
            return -p0 * ((double)zz + p1 * (double)z);
        }
        // East: B * (n_0 + sum_{i>=1} n_{i-1} n_i + c n_{i-1}) + c       east_model.jl:38-50
        const unsigned long long all = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
        const unsigned long long zero = ~C & all; // bit i set iff n_i = 1
        const unsigned long long prev = zero & ((n >= 2) ? ((1ull << (n - 1)) - 1ull) : 0ull); // n_{i-1}, i = 1..n-1
        const int nn = popc64(prev & (zero >> 1));
        const int np = popc64(prev);
        return p2 * ((double)(zero & 1ull) + (double)nn + p0 * (double)np) + p0;
    }
    QCB_HD double offdiag_coef() const { return kind == 0 ? -p0 * p2 : p1; }
    // weight of the flip of bit i at configuration C
    QCB_HD bool flip_on(unsigned long long C, int i) const
    {
        if (kind == 0 || i == 0) return true;
        return ((C >> (i - 1)) & 1ull) == 0ull; // n_{i-1}
    }
};

// E = sum_C conj(psi[C]) (H psi)[C], accumulated in double; one pass, no scratch state.
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_measure(const typename Vec2<R>::type *__restrict__ psi,
                                                         unsigned long long namps, HamEval H, double2 *partial)
{
    using V = typename Vec2<R>::type;
    double acc[2] = {0.0, 0.0};
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const double oc = H.offdiag_coef();
    for (unsigned long long C = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; C < namps; C += stride) {
        const V p = psi[C];
        const double pr = (double)p.x, pi = (double)p.y;
        double xr = 0, xi = 0;
        for (int i = 0; i < H.n; i++) {
            if (!H.flip_on(C, i)) continue;
            const V o = psi[C ^ (1ull << i)];
            xr += (double)o.x;
            xi += (double)o.y;
        }
        // conj(x) * p  (tfim.jl:39-46: x = sum conj(psi_other); tf = x * psi)
        acc[0] += H.diag(C) * (pr * pr + pi * pi) + oc * (xr * pr + xi * pi);
        acc[1] += oc * (xr * pi - xi * pr);
    }
    __shared__ double out[2];
    block_sum<2>(acc, out);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
}

// partner loads in flight per thread in the gathers of k_apply_ham_tiled (measured: 4 / 8 / 16 -> 6.80 / 6.72 / 7.46 ms at 28 q
// ComplexF32 against 7.57 ms for a plain loop; the same batching does nothing for the DRAM-bound k_apply_ham<double>)
constexpr int kHamTileBatch = 8;
// lambda = H psi (out of place); optionally also E = <psi|H|psi> = sum_C conj(psi[C]) lambda[C] in the same sweep
// (partial != nullptr: one double2 per CTA, like k_measure).
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_apply_ham(typename Vec2<R>::type *__restrict__ lam,
                                                           const typename Vec2<R>::type *__restrict__ psi, unsigned long long namps,
                                                           HamEval H, double2 *partial)
{
    using V = typename Vec2<R>::type;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const double oc = H.offdiag_coef();
    double acc[2] = {0.0, 0.0};
    for (unsigned long long C = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; C < namps; C += stride) {
        const V p = psi[C];
        double xr = 0, xi = 0;
        for (int i = 0; i < H.n; i++) {
            if (!H.flip_on(C, i)) continue;
            const V o = psi[C ^ (1ull << i)];
            xr += (double)o.x;
            xi += (double)o.y;
        }
        const double d = H.diag(C);
        const double lr = d * (double)p.x + oc * xr, li = d * (double)p.y + oc * xi;
        V r;
        r.x = (R)lr;
        r.y = (R)li;
        lam[C] = r;
        acc[0] += (double)p.x * lr + (double)p.y * li;
        acc[1] += (double)p.x * li - (double)p.y * lr;
    }
    if (partial) {
        __shared__ double out[2];
        block_sum<2>(acc, out);
        if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
    }
}

// The same sweep with the flips of the low TBL index bits served from shared memory: a CTA stages 2^TBL consecutive amplitudes
// (32 KiB) and only the flips of bits >= TBL go to L1 / L2.  Chunks are taken in index order by the whole grid.  The sums run in
// the same order as in k_apply_ham: lambda is bit-identical to it.  Default for ComplexF32, where the gather form is bound by
// load instructions (28 q: 9.1 -> 6.5 ms); the ComplexF64 gather is bound by DRAM (its high-bit partner streams miss L2:
// 9.5 state-sized reads) and gets slower with tiles (more CTAs' partner streams compete for L2: 12-13 reads) -- DESIGN 2.2.
// The two per-thread phases are __host__ __device__ so that the CPU suite replays them (tests/emu).
template <typename R, int TBL>
QCB_HD void ham_tile_load(typename Vec2<R>::type *sm, const typename Vec2<R>::type *__restrict__ psi, unsigned long long base, uint32_t t)
{
#pragma unroll
    for (int j = 0; j < (1 << TBL) / kRedThreads; j++) sm[t + j * kRedThreads] = psi[base + t + j * kRedThreads];
}
template <typename R, int TBL>
QCB_HD void ham_tile_rows(typename Vec2<R>::type *__restrict__ lam, const typename Vec2<R>::type *__restrict__ psi,
                          const typename Vec2<R>::type *sm, unsigned long long base, uint32_t t, const HamEval &H, double oc, double &er, double &ei)
{
    using V = typename Vec2<R>::type;
#pragma unroll 1
    for (int j = 0; j < (1 << TBL) / kRedThreads; j++) {
        const uint32_t e = t + j * kRedThreads;
        const unsigned long long C = base + e;
        const V p = sm[e];
        double xr = 0, xi = 0;
#pragma unroll
        for (int i = 0; i < TBL; i++) { // (H.n >= TBL)
            if (!H.flip_on(C, i)) continue;
            const V o = sm[e ^ (1u << i)];
            xr += (double)o.x;
            xi += (double)o.y;
        }
        // gathers in batches of kHamTileBatch loads in flight (a flip that is off, or a bit past the last one, loads a valid
        // address and is not added): the sweep is latency-bound otherwise (ncu: long_scoreboard, 3 loads per thread in flight)
#pragma unroll 1
        for (int i0 = TBL; i0 < H.n; i0 += kHamTileBatch) {
            V o[kHamTileBatch];
#pragma unroll
            for (int k = 0; k < kHamTileBatch; k++) o[k] = psi[C ^ ((i0 + k < H.n) ? (1ull << (i0 + k)) : 0ull)];
#pragma unroll
            for (int k = 0; k < kHamTileBatch; k++) {
                if (i0 + k >= H.n || !H.flip_on(C, i0 + k)) continue;
                xr += (double)o[k].x;
                xi += (double)o[k].y;
            }
        }
        const double d = H.diag(C);
        const double lr = d * (double)p.x + oc * xr, li = d * (double)p.y + oc * xi;
        V r;
        r.x = (R)lr;
        r.y = (R)li;
        lam[C] = r;
        er += (double)p.x * lr + (double)p.y * li;
        ei += (double)p.x * li - (double)p.y * lr;
    }
}
#if defined(__CUDACC__)
template <typename R, int TBL>
__global__ void __launch_bounds__(kRedThreads) k_apply_ham_tiled(typename Vec2<R>::type *__restrict__ lam,
                                                                 const typename Vec2<R>::type *__restrict__ psi, unsigned long long nchunks,
                                                                 HamEval H, double2 *partial)
{
    __shared__ typename Vec2<R>::type sm[1 << TBL];
    const double oc = H.offdiag_coef();
    double acc[2] = {0.0, 0.0};
    for (unsigned long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        ham_tile_load<R, TBL>(sm, psi, ch << TBL, threadIdx.x);
        __syncthreads();
        ham_tile_rows<R, TBL>(lam, psi, sm, ch << TBL, threadIdx.x, H, oc, acc[0], acc[1]);
        __syncthreads();
    }
    if (partial) {
        __shared__ double out[2];
        block_sum<2>(acc, out);
        if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
    }
}
#endif

// E = <psi| H |psi> for an explicit (sparse) Hamiltonian in CSR form: src/circuits.jl:49-55, real(dot(psi, H, psi)).
// One row per thread (the reference's matrices have n+1 entries per row), accumulated in double.
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_measure_csr(const typename Vec2<R>::type *__restrict__ psi, long long nrows,
                                                             const long long *__restrict__ rowptr, const long long *__restrict__ colidx,
                                                             const double2 *__restrict__ vals, double2 *partial)
{
    double acc[2] = {0.0, 0.0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
        double hr = 0, hi = 0; // (H psi)[r]
        for (long long k = rowptr[r]; k < rowptr[r + 1]; k++) {
            const double2 v = vals[k];
            const auto x = psi[colidx[k]];
            hr += v.x * (double)x.x - v.y * (double)x.y;
            hi += v.x * (double)x.y + v.y * (double)x.x;
        }
        const auto p = psi[r];
        acc[0] += (double)p.x * hr + (double)p.y * hi;
        acc[1] += (double)p.x * hi - (double)p.y * hr;
    }
    __shared__ double out[2];
    block_sum<2>(acc, out);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
}

// ---- lambda = H psi on a sharded state ------------------------------------------------------------------------------
// Layout of a shard: local index bit phys[q] (< nloc) or rank bit phys[q] - nloc holds logical qubit q.  A flip of logical
// qubit i pairs amplitudes of the same shard only while phys[i] < nloc, so lambda is built in two sweeps with one qubit
// swap between them (canonical layout: the qubits below nloc; after swap_top: the others); `mask` selects the logical
// qubits whose flips this sweep adds, `first` also writes the diagonal term (otherwise lambda is accumulated).
struct ShardMap {
    unsigned char phys[40];
    int n, nloc;
    unsigned long long hi; // rank << nloc
};
QCB_HD unsigned long long logical_config(const ShardMap &M, unsigned long long x)
{
    const unsigned long long X = x | M.hi;
    unsigned long long C = 0;
    for (int q = 0; q < M.n; q++) C |= ((X >> M.phys[q]) & 1ull) << q;
    return C;
}
// one amplitude; adds conj(psi[x]) * (this sweep's contribution to lambda[x]) to (er, ei)
template <typename V>
QCB_HD void ham_sharded_amp(V *lam, const V *psi, unsigned long long x, const ShardMap &M, const HamEval &H, unsigned long long mask,
                            bool first, double &er, double &ei)
{
    const unsigned long long C = logical_config(M, x);
    const V p = psi[x];
    double xr = 0, xi = 0;
    for (int i = 0; i < H.n; i++) {
        if (!((mask >> i) & 1ull) || !H.flip_on(C, i)) continue;
        const V o = psi[x ^ (1ull << M.phys[i])];
        xr += (double)o.x;
        xi += (double)o.y;
    }
    const double oc = H.offdiag_coef();
    double lr = oc * xr, li = oc * xi;
    if (first) {
        const double d = H.diag(C);
        lr += d * (double)p.x;
        li += d * (double)p.y;
    }
    er += (double)p.x * lr + (double)p.y * li;
    ei += (double)p.x * li - (double)p.y * lr;
    V r = first ? V{} : lam[x];
    r.x = (decltype(r.x))((double)r.x + lr);
    r.y = (decltype(r.y))((double)r.y + li);
    lam[x] = r;
}
#if defined(__CUDACC__)
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_apply_ham_sharded(typename Vec2<R>::type *lam, const typename Vec2<R>::type *__restrict__ psi,
                                                                   unsigned long long namps, const __grid_constant__ ShardMap M, HamEval H,
                                                                   unsigned long long mask, int first, double2 *partial)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    double acc[2] = {0.0, 0.0};
    for (unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; x < namps; x += stride)
        ham_sharded_amp(lam, psi, x, M, H, mask, first != 0, acc[0], acc[1]);
    __shared__ double out[2];
    block_sum<2>(acc, out);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
}
#endif

// lambda = H psi for an explicit CSR Hamiltonian (and E = <psi|lambda> partials): the adjoint gradient with a matrix
// Hamiltonian (the reference's gradients!(..., H::AbstractMatrix, ...) through measure!, examples/test_gradients.jl:25)
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_apply_csr(typename Vec2<R>::type *lam, const typename Vec2<R>::type *__restrict__ psi, long long nrows,
                                                           const long long *__restrict__ rowptr, const long long *__restrict__ colidx,
                                                           const double2 *__restrict__ vals, double2 *partial)
{
    using V = typename Vec2<R>::type;
    double acc[2] = {0.0, 0.0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
        double hr = 0, hi = 0;
        for (long long k = rowptr[r]; k < rowptr[r + 1]; k++) {
            const double2 v = vals[k];
            const V x = psi[colidx[k]];
            hr += v.x * (double)x.x - v.y * (double)x.y;
            hi += v.x * (double)x.y + v.y * (double)x.x;
        }
        const V p = psi[r];
        V o;
        o.x = (R)hr;
        o.y = (R)hi;
        lam[r] = o;
        acc[0] += (double)p.x * hr + (double)p.y * hi;
        acc[1] += (double)p.x * hi - (double)p.y * hr;
    }
    __shared__ double out[2];
    block_sum<2>(acc, out);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
}

// generic fixed-order final reduction: out[v] = sum_b partial[b * nv + v]
static __global__ void k_final_sum(const double *__restrict__ partial, int nblocks, int nv, double *out)
{
    const int v = blockIdx.x;
    double s = 0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(size_t)b * nv + v];
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        out[v] = t;
    }
}

// sum |amp|^2
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_norm2(const typename Vec2<R>::type *__restrict__ psi, unsigned long long namps,
                                                       double *partial)
{
    double acc[1] = {0.0};
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long C = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; C < namps; C += stride) {
        const auto p = psi[C];
        acc[0] += (double)p.x * (double)p.x + (double)p.y * (double)p.y;
    }
    __shared__ double out[1];
    block_sum<1>(acc, out);
    if (threadIdx.x == 0) partial[blockIdx.x] = out[0];
}

// <a|b> = sum conj(a) b
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_inner(const typename Vec2<R>::type *__restrict__ a, const typename Vec2<R>::type *__restrict__ b,
                                                       unsigned long long namps, double2 *partial)
{
    double acc[2] = {0.0, 0.0};
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long C = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; C < namps; C += stride) {
        const auto x = a[C];
        const auto y = b[C];
        acc[0] += (double)x.x * (double)y.x + (double)x.y * (double)y.y;
        acc[1] += (double)x.x * (double)y.y - (double)x.y * (double)y.x;
    }
    __shared__ double out[2];
    block_sum<2>(acc, out);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
}

// 4x4 environment of the qubit pair (kb, kb+1): env[r + 4c] = sum_rest conj(lam[r, rest]) psi[c, rest]  (no state is modified)
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_env_2q(const typename Vec2<R>::type *__restrict__ psi, const typename Vec2<R>::type *__restrict__ lam,
                                                        unsigned long long nquads, int kb, double *partial)
{
    using V = typename Vec2<R>::type;
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long low = (1ull << kb) - 1ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nquads; i += stride) {
        const unsigned long long base = ((i >> kb) << (kb + 2)) | (i & low);
        V x[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { x[j] = psi[base | ((unsigned long long)j << kb)]; l[j] = lam[base | ((unsigned long long)j << kb)]; }
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const double lr = (double)l[r].x, li = (double)l[r].y, xr = (double)x[c].x, xi = (double)x[c].y;
                acc[2 * (r + 4 * c)] += lr * xr + li * xi;
                acc[2 * (r + 4 * c) + 1] += lr * xi - li * xr;
            }
    }
    __shared__ double out[32];
    block_sum<32>(acc, out);
    if (threadIdx.x < 32) partial[(size_t)blockIdx.x * 32 + threadIdx.x] = out[threadIdx.x];
}

// ---- adjoint sweep, unfused step for one gate on bits (kb, kb+1) ---------------------
//   psi <- U^dagger psi ;  env[r + 4c] += conj(lam[r]) psi[c] ;  lam <- U^dagger lam
// Ud holds U^dagger.  partial: [gridDim.x][32] doubles.
template <typename R>
__global__ void __launch_bounds__(kRedThreads) k_adjoint_step(typename Vec2<R>::type *psi, typename Vec2<R>::type *lam,
                                                              unsigned long long nquads, int kb,
                                                              const __grid_constant__ GateMat<R> Ud, double *partial)
{
    using V = typename Vec2<R>::type;
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long low = (1ull << kb) - 1ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nquads; i += stride) {
        const unsigned long long base = ((i >> kb) << (kb + 2)) | (i & low);
        V x[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { x[j] = psi[base | ((unsigned long long)j << kb)]; l[j] = lam[base | ((unsigned long long)j << kb)]; }
        apply_quad(x[0], x[1], x[2], x[3], Ud);
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const double lr = (double)l[r].x, li = (double)l[r].y, xr = (double)x[c].x, xi = (double)x[c].y;
                acc[2 * (r + 4 * c)] += lr * xr + li * xi;     // Re conj(l) x
                acc[2 * (r + 4 * c) + 1] += lr * xi - li * xr; // Im conj(l) x
            }
        apply_quad(l[0], l[1], l[2], l[3], Ud);
#pragma unroll
        for (int j = 0; j < 4; j++) { psi[base | ((unsigned long long)j << kb)] = x[j]; lam[base | ((unsigned long long)j << kb)] = l[j]; }
    }
    __shared__ double out[32];
    block_sum<32>(acc, out);
    if (threadIdx.x < 32) partial[(size_t)blockIdx.x * 32 + threadIdx.x] = out[threadIdx.x];
}

} // namespace qcb
