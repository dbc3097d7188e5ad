// measure_tile.cuh -- tiled sweep-and-reduce expectation value of the kernel Hamiltonians.
//
// Replaces _tfim_measure! / _eastmodel_measure! (src/hamiltonians/tfim.jl:17-54, east_model.jl:30-52) plus the
// separate sum(psi') pass (src/hamiltonians/hamiltonians.jl:41).  The reference gathers N+1 amplitudes per
// amplitude; k_measure (kernels.cuh) does the same from L2/HBM.  Here the flips are done inside shared-memory
// tiles: a pass stages 2^TB amplitudes, every thread pulls 16-amplitude register windows (4 consecutive tile bits)
// and accumulates  2 w_i(C) Re(conj(psi[C | 2^i]) psi[C])  for the window bits i (each pair once; the imaginary
// parts of the two orientations cancel exactly) and, in the first stage of the first pass, diag(C) |psi[C]|^2.
// HBM traffic: one read of the state per pass, ceil((n - TB) / (TB - low_bits)) + 1 passes; nothing is written.
#pragma once
#include "kernels.cuh"

namespace qcb {

constexpr int kMeasMaxStages = 4;

struct MeasStage {
    StageDesc sd;
    uint8_t active;        // which of the 4 window bits contribute flips in this stage
    uint8_t diag;          // 1: this stage also accumulates the diagonal part
    uint8_t wphys[kRegBits]; // physical index bit of each window bit
    int8_t ctl[kRegBits];    // East: physical index bit that holds the control qubit i-1 of window bit b (-1: none, qubit 0);
                             // positions >= nloc address the rank bits of a sharded state (c_hi)
    uint8_t pad[2];
};

template <typename R> struct MeasParams {
    int32_t nstages, nseg;
    uint8_t seg_start[kMaxSeg], seg_len[kMaxSeg];
    uint8_t ld_phys[kMaxTB], ld_lpos[kMaxTB];
    unsigned long long ld_goff[kRegAmps];
    uint16_t ld_eoff[kRegAmps];
    MeasStage stage[kMeasMaxStages];
    HamEval H;
    unsigned long long c_hi; // index bits above the local shard (rank << nloc), 0 for unsharded states
};

QCB_HD double ham_diag(const HamEval &h, unsigned long long C)
{
    const int n = h.n;
    const unsigned long long pairmask = (n >= 2) ? ((1ull << (n - 1)) - 1ull) : 0ull;
    if (h.kind == 0) {
        const int eq = popc64(~(C ^ (C >> 1)) & pairmask);
        return -h.p0 * ((double)(2 * eq - (n - 1)) + h.p1 * (double)(n - 2 * popc64(C)));
    }
    const unsigned long long all = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    const unsigned long long zero = ~C & all;
    const unsigned long long prev = zero & pairmask;
    return h.p2 * ((double)(zero & 1ull) + (double)popc64(prev & (zero >> 1)) + h.p0 * (double)popc64(prev)) + h.p0;
}

// one stage of one thread: adds its contribution to (diagonal sum, flip sum).  KIND is the Hamiltonian (compile time:
// the TFIM flips need no index arithmetic at all, the East flips need one control bit per pair).
template <typename R, int TB, int KIND>
QCB_HD void measure_stage(const MeasParams<R> &P, int s, const typename Vec2<R>::type *sm, uint32_t tid, unsigned long long tile_base,
                          double &acc_diag, double &acc_flip)
{
    using V = typename Vec2<R>::type;
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    const MeasStage &S = P.stage[s];
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) e |= ((tid >> k) & 1u) << S.sd.tpos[k];
    const uint32_t es = swz<SWB>(e);
    V a[kRegAmps];
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) a[r] = sm[es ^ S.sd.roff[r]];
    // physical index of this thread's element r = 0 (needed for the diagonal and for the East control bits)
    unsigned long long C0 = tile_base | P.c_hi;
    if (S.diag || KIND == 1) {
#pragma unroll
        for (int k = 0; k < TB; k++) C0 |= (unsigned long long)((e >> P.ld_lpos[k]) & 1u) << P.ld_phys[k];
    }
    if (S.diag) {
#pragma unroll
        for (int r = 0; r < kRegAmps; r++) {
            unsigned long long C = C0;
#pragma unroll
            for (int b = 0; b < kRegBits; b++) C |= (unsigned long long)((r >> b) & 1) << S.wphys[b];
            acc_diag += ham_diag(P.H, C) * ((double)a[r].x * (double)a[r].x + (double)a[r].y * (double)a[r].y);
        }
    }
#pragma unroll
    for (int b = 0; b < kRegBits; b++) {
        if (!((S.active >> b) & 1)) continue;
        // East: the flip of qubit i is weighted by n_{i-1} = [bit i-1 is 0]; bit i-1 is either the previous window
        // bit (then it depends on r) or fixed for this thread
        uint32_t allowed = 0xffffu;
        if (KIND == 1 && S.ctl[b] >= 0) {
            const int ctl = S.ctl[b];
            if (b > 0 && S.wphys[b - 1] == ctl) allowed = b == 1 ? 0x5555u : (b == 2 ? 0x3333u : 0x0f0fu);
            else allowed = ((C0 >> ctl) & 1ull) ? 0u : 0xffffu;
        }
#pragma unroll
        for (int r = 0; r < kRegAmps; r++) {
            if ((r >> b) & 1) continue;
            if (KIND == 1 && !((allowed >> r) & 1u)) continue;
            const V x0 = a[r], x1 = a[r | (1 << b)];
            acc_flip += 2.0 * ((double)x0.x * (double)x1.x + (double)x0.y * (double)x1.y);
        }
    }
}

template <typename R, int TB>
QCB_HD void measure_tile_load(const typename Vec2<R>::type *__restrict__ src, const MeasParams<R> &P, typename Vec2<R>::type *sm,
                              uint32_t tid, unsigned long long bid, unsigned long long &tile_base)
{
    using V = typename Vec2<R>::type;
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    unsigned long long g = 0, b = bid;
    for (int s = 0; s < P.nseg; s++) {
        const int len = P.seg_len[s];
        g |= (b & ((1ull << len) - 1ull)) << P.seg_start[s];
        b >>= len;
    }
    tile_base = g;
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) {
        const uint32_t bit = (tid >> k) & 1u;
        g |= (unsigned long long)bit << P.ld_phys[k];
        e |= bit << P.ld_lpos[k];
    }
    e = swz<SWB>(e);
    V tmp[kRegAmps];
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) tmp[i] = src[g | P.ld_goff[i]];
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) sm[e ^ P.ld_eoff[i]] = tmp[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent form (option "measure_persistent", off by default until it has been measured on a GPU).  ncu on the kernel
// below (one tile per CTA) shows it bound by instruction issue, not by HBM: 93-126 warp instructions per amplitude and
// pass, a third of them in the tile load (segment loops, per-thread index deposits) and another third in the stage
// prologues (window offsets, physical index) -- all of it work that does not depend on the tile.  Here a CTA walks many
// tiles and keeps the tile-independent part in registers: its share of the global index and its load offset, and per
// stage the swizzled window offset and the thread's share of the physical index.  Same arithmetic, same summation order
// per thread; the partial sums are per persistent CTA instead of per tile.
template <typename R, int TB> struct MeasThread {
    unsigned long long gthread;               // this thread's bits of the global index (tile load)
    uint32_t e_load;                          // swizzled tile-local offset of its element 0 (tile load)
    uint32_t es[kMeasMaxStages];              // per stage: swizzled offset of its window
    unsigned long long cth[kMeasMaxStages];   // per stage: its bits of the physical index (without tile base and c_hi)
};
template <typename R, int TB> QCB_HD MeasThread<R, TB> measure_thread_consts(const MeasParams<R> &P, uint32_t tid)
{
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    MeasThread<R, TB> T;
    T.gthread = 0;
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) {
        const uint32_t bit = (tid >> k) & 1u;
        T.gthread |= (unsigned long long)bit << P.ld_phys[k];
        e |= bit << P.ld_lpos[k];
    }
    T.e_load = swz<SWB>(e);
#pragma unroll
    for (int s = 0; s < kMeasMaxStages; s++) {
        uint32_t es = 0;
        if (s < P.nstages) {
#pragma unroll
            for (int k = 0; k < TBITS; k++) es |= ((tid >> k) & 1u) << P.stage[s].sd.tpos[k];
        }
        unsigned long long cth = 0;
#pragma unroll
        for (int k = 0; k < TB; k++) cth |= (unsigned long long)((es >> P.ld_lpos[k]) & 1u) << P.ld_phys[k];
        T.es[s] = swz<SWB>(es);
        T.cth[s] = cth;
    }
    return T;
}
// global index bits of tile `bid` (the non-tile bits, segment by segment)
template <typename R> QCB_HD unsigned long long measure_tile_base(const MeasParams<R> &P, unsigned long long bid)
{
    unsigned long long g = 0;
    for (int s = 0; s < P.nseg; s++) {
        const int len = P.seg_len[s];
        g |= (bid & ((1ull << len) - 1ull)) << P.seg_start[s];
        bid >>= len;
    }
    return g;
}
template <typename R, int TB>
QCB_HD void measure_tile_load_pre(const typename Vec2<R>::type *__restrict__ src, const MeasParams<R> &P, typename Vec2<R>::type *sm,
                                  const MeasThread<R, TB> &T, unsigned long long tile_base)
{
    using V = typename Vec2<R>::type;
    const unsigned long long g = tile_base | T.gthread;
    V tmp[kRegAmps];
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) tmp[i] = src[g | P.ld_goff[i]];
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) sm[T.e_load ^ P.ld_eoff[i]] = tmp[i];
}
// measure_stage with the thread's window offset `es` (swizzled) and physical index C0 given
template <typename R, int TB, int KIND>
QCB_HD void measure_stage_pre(const MeasParams<R> &P, int s, const typename Vec2<R>::type *sm, uint32_t es, unsigned long long C0,
                              double &acc_diag, double &acc_flip)
{
    using V = typename Vec2<R>::type;
    const MeasStage &S = P.stage[s];
    V a[kRegAmps];
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) a[r] = sm[es ^ S.sd.roff[r]];
    if (S.diag) {
#pragma unroll
        for (int r = 0; r < kRegAmps; r++) {
            unsigned long long C = C0;
#pragma unroll
            for (int b = 0; b < kRegBits; b++) C |= (unsigned long long)((r >> b) & 1) << S.wphys[b];
            acc_diag += ham_diag(P.H, C) * ((double)a[r].x * (double)a[r].x + (double)a[r].y * (double)a[r].y);
        }
    }
#pragma unroll
    for (int b = 0; b < kRegBits; b++) {
        if (!((S.active >> b) & 1)) continue;
        uint32_t allowed = 0xffffu;
        if (KIND == 1 && S.ctl[b] >= 0) {
            const int ctl = S.ctl[b];
            if (b > 0 && S.wphys[b - 1] == ctl) allowed = b == 1 ? 0x5555u : (b == 2 ? 0x3333u : 0x0f0fu);
            else allowed = ((C0 >> ctl) & 1ull) ? 0u : 0xffffu;
        }
#pragma unroll
        for (int r = 0; r < kRegAmps; r++) {
            if ((r >> b) & 1) continue;
            if (KIND == 1 && !((allowed >> r) & 1u)) continue;
            const V x0 = a[r], x1 = a[r | (1 << b)];
            acc_flip += 2.0 * ((double)x0.x * (double)x1.x + (double)x0.y * (double)x1.y);
        }
    }
}

#if defined(__CUDACC__)
template <typename R, int TB, int MINB, int KIND>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
measure_tile_persistent_kernel(const typename Vec2<R>::type *psi, const __grid_constant__ MeasParams<R> P, unsigned long long ntiles,
                               double2 *partial /* [gridDim.x] */)
{
    using V = typename Vec2<R>::type;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    V *sm = reinterpret_cast<V *>(qcb_smem_raw);
    const MeasThread<R, TB> T = measure_thread_consts<R, TB>(P, threadIdx.x);
    double acc[2] = {0.0, 0.0};
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned long long base = measure_tile_base(P, tile);
        measure_tile_load_pre<R, TB>(psi, P, sm, T, base);
        __syncthreads();
#pragma unroll
        for (int s = 0; s < kMeasMaxStages; s++)
            if (s < P.nstages) measure_stage_pre<R, TB, KIND>(P, s, sm, T.es[s], base | P.c_hi | T.cth[s], acc[0], acc[1]);
        __syncthreads(); // the next tile's stores overwrite what other threads still read
    }
    __shared__ double out[2];
    block_sum_n<2, (1 << (TB - kRegBits)) / 32>(acc, out);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
}
#endif

#if defined(__CUDACC__)
template <typename R, int TB, int MINB, int KIND>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
measure_tile_kernel(const typename Vec2<R>::type *psi, const __grid_constant__ MeasParams<R> P, double2 *partial)
{
    using V = typename Vec2<R>::type;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    V *sm = reinterpret_cast<V *>(qcb_smem_raw);
    const uint32_t tid = threadIdx.x;
    unsigned long long tile_base;
    measure_tile_load<R, TB>(psi, P, sm, tid, blockIdx.x, tile_base);
    __syncthreads();
    double acc[2] = {0.0, 0.0};
#pragma unroll 1
    for (int s = 0; s < P.nstages; s++) measure_stage<R, TB, KIND>(P, s, sm, tid, tile_base, acc[0], acc[1]);
    __shared__ double out[2];
    block_sum_n<2, (1 << (TB - kRegBits)) / 32>(acc, out);
    if (tid == 0) partial[blockIdx.x] = make_double2(out[0], out[1]);
}
#endif

} // namespace qcb
