// flip_tile.cuh -- expectation values (and lambda = H psi, flip_ham.cuh) of the kernel Hamiltonians as a persistent,
// TMA-fed sweep over shared-memory tiles.
//
// Replaces _tfim_measure! / _eastmodel_measure! (src/hamiltonians/tfim.jl:17-54, east_model.jl:30-52) plus the separate
// sum(psi') pass (src/hamiltonians/hamiltonians.jl:41):
//     E = sum_C diag(C) |psi_C|^2  +  coef * sum_i sum_{C : bit i clear} w_i(C) * 2 Re(conj(psi[C | 2^i]) psi[C])
// (each flip pair once; the two orientations have exactly opposite imaginary parts).  The reference gathers n + 1
// amplitudes per amplitude; here a CTA per SM walks 64 KiB tiles that stream into a three-slot shared-memory ring with TMA
// (cp.async.bulk.tensor, one mbarrier per slot), and two groups of eight warps pull 16-unit register windows out of the
// ring and do the flips in registers.
//
// Tile geometry (unit = 16 bytes = one ComplexF64 or two ComplexF32 amplitudes; 4096 units per tile, 12 unit bits):
//   level 0   : 2^12 contiguous units (amplitude bits [0, 12) / [0, 13)): all flips inside the tile + the diagonal term
//   level >= 1: 128-byte rows (unit bits 0..2) x a window of 9 index bits [w0, w0 + 9) (unit bits 3..11): the flips of
//               the window bits not covered by an earlier level
// Work order: levels 0 and 1 are interleaved chunk by chunk (chunk = 2^21 / 2^22 amplitudes = 32 MiB) so that the
// level-1 tiles re-read their chunk from L2 while it is still resident; further levels are separate sweeps.  HBM
// traffic at 30 qubits: 2 reads of the state (was 3), nothing written.
//
// The TMA boxes are 128-byte rows with the 128-byte swizzle: unit u = (row << 3) | chunk sits at u ^ ((u >> 3) & 7).
// A register stage is conflict free when the three lane bits of a quarter warp are unit bits {0,1,2} or {3,4,5}; the
// stage tables below are built that way.  All index algebra is __host__ __device__: tests/emu replays it on the CPU.
#pragma once
#include "kernels.cuh"
#include "tma.cuh"

namespace qcb {

constexpr int kFlipUnitBits = 12;                       // units per tile
constexpr int kFlipTileBytes = 16 << kFlipUnitBits;     // 64 KiB
constexpr int kFlipWin = 9;                             // window bits of a level >= 1 tile
constexpr int kFlipRing = 3;                            // tiles in flight per CTA
constexpr int kFlipConsumers = 256;                     // consumer threads per tile (8 warps: 16 units per thread and stage)
constexpr int kFlipGroups = 2;                          // consumer groups per CTA, each on its own tile (tile k -> group k mod 2);
                                                        // the warp after the last group is the producer
constexpr int kFlipMaxLevels = 5;

template <typename R> struct FlipGeom;
template <> struct FlipGeom<double> { static constexpr int NA = 1, SUB = 0, TBA = 12, C = 3; }; // amplitudes per unit, tile amplitude bits, row bits
template <> struct FlipGeom<float> { static constexpr int NA = 2, SUB = 1, TBA = 13, C = 4; };

// stage tables (unit bits).  LV = 0: level-0 tiles, LV = 1: window tiles.
QCB_HD constexpr int flip_win(int lv, int s, int b)
{
    return lv == 0 ? (s == 0 ? (b < 3 ? b : 6) : s == 1 ? (b < 3 ? 3 + b : 7) : 8 + b)
                   : (s == 0 ? 3 + b : s == 1 ? 7 + b : (b < 3 ? b : 11));
}
QCB_HD constexpr bool flip_in_win(int lv, int s, int u)
{
    return flip_win(lv, s, 0) == u || flip_win(lv, s, 1) == u || flip_win(lv, s, 2) == u || flip_win(lv, s, 3) == u;
}
QCB_HD constexpr int flip_win_index(int lv, int s, int u)
{
    return flip_win(lv, s, 0) == u ? 0 : flip_win(lv, s, 1) == u ? 1 : flip_win(lv, s, 2) == u ? 2 : flip_win(lv, s, 3) == u ? 3 : -1;
}
// thread bit k (0..7) -> unit bit: the three lane bits first ({3,4,5} when the window holds {0,1,2}, else {0,1,2}), then
// the remaining free unit bits in ascending order
QCB_HD constexpr int flip_tpos(int lv, int s, int k)
{
    const bool low_in_win = flip_in_win(lv, s, 0);
    if (k < 3) return low_in_win ? 3 + k : k;
    int seen = 0;
    for (int u = 0; u < kFlipUnitBits; u++) {
        if (flip_in_win(lv, s, u)) continue;
        if (low_in_win ? (u >= 3 && u <= 5) : (u <= 2)) continue;
        if (seen == k - 3) return u;
        seen++;
    }
    return -1;
}

struct FlipLevel {
    int32_t w0;        // level >= 1: first index bit of the window
    int32_t mid_bits;  // w0 - C: bits of the row coordinate between the 128-byte row and the window
    uint32_t active;   // unit bits whose flips this level adds (level 0: all 12)
    int32_t pad;
};

struct FlipParams {
    HamEval H;
    int32_t n, nlevels;
    int32_t tiles_log2;       // tiles per level = 2^(n - TBA)
    int32_t chunk_log2;       // tiles per chunk and level in the interleaved region
    int32_t interleaved;      // 1: levels 0 and 1 are interleaved chunk by chunk
    int32_t pad;
    unsigned long long n_items;
    FlipLevel level[kFlipMaxLevels];
    double diag_T[32];        // in-window part of diag(C) for the diagonal stage (see flip_fill_diag_table)
    unsigned long long c_hi;  // index bits above the local state (sharded states: rank << nloc)
};

struct FlipItem {
    int32_t lv;               // level
    unsigned long long t;     // tile of that level
};

QCB_HD FlipItem flip_decode(const FlipParams &P, unsigned long long j)
{
    FlipItem it;
    const unsigned long long ntiles = 1ull << P.tiles_log2;
    const unsigned long long n01 = P.interleaved ? 2ull * ntiles : ntiles;
    if (j < n01) {
        if (P.interleaved) {
            const unsigned long long chunk = j >> (P.chunk_log2 + 1), jj = j & ((2ull << P.chunk_log2) - 1ull);
            it.lv = (int32_t)(jj >> P.chunk_log2);
            it.t = (chunk << P.chunk_log2) | (jj & ((1ull << P.chunk_log2) - 1ull));
        } else {
            it.lv = 0;
            it.t = j;
        }
    } else {
        const unsigned long long r = j - n01;
        it.lv = (P.interleaved ? 2 : 1) + (int32_t)(r >> P.tiles_log2);
        it.t = r & (ntiles - 1ull);
    }
    return it;
}

// first amplitude index of a tile (window / tile bits zero)
template <typename R> QCB_HD unsigned long long flip_tile_base(const FlipParams &P, const FlipItem &it)
{
    using G = FlipGeom<R>;
    if (it.lv == 0) return it.t << G::TBA;
    const FlipLevel &L = P.level[it.lv];
    const unsigned long long mid = it.t & ((1ull << L.mid_bits) - 1ull), hi = it.t >> L.mid_bits;
    return (mid << G::C) | (hi << (L.w0 + kFlipWin));
}

// physical index bit of unit bit u of a tile of level lv
template <typename R> QCB_HD int flip_phys_of_unit(const FlipParams &P, int lv, int u)
{
    using G = FlipGeom<R>;
    if (lv == 0 || u < 3) return u + G::SUB;
    return P.level[lv].w0 + (u - 3);
}

// host: levels for an n-qubit local state.  Returns false if n is too small for this path.
template <typename R> inline bool flip_plan(FlipParams &P, int nloc, const HamEval &H, unsigned long long c_hi)
{
    using G = FlipGeom<R>;
    if (nloc < G::TBA + 2) return false; // (level >= 1 windows need w0 > C: nloc - 9 > C)
    P = FlipParams();
    P.H = H;
    P.n = nloc;
    P.c_hi = c_hi;
    P.tiles_log2 = nloc - G::TBA;
    int covered = G::TBA, nl = 1;
    P.level[0].w0 = 0; P.level[0].mid_bits = 0; P.level[0].active = (1u << kFlipUnitBits) - 1u;
    while (covered < nloc) {
        if (nl >= kFlipMaxLevels) return false;
        FlipLevel &L = P.level[nl];
        L.w0 = covered < nloc - kFlipWin ? covered : nloc - kFlipWin;
        L.mid_bits = L.w0 - G::C;
        L.active = 0;
        for (int u = 3; u < kFlipUnitBits; u++)
            if (L.w0 + (u - 3) >= covered) L.active |= 1u << u;
        covered = L.w0 + kFlipWin;
        nl++;
    }
    P.nlevels = nl;
    // levels 0 and 1 chunk by chunk when level 1's window sits directly on top of the level-0 tile bits
    P.interleaved = nl >= 2 ? 1 : 0;
    P.chunk_log2 = (nl >= 2 && P.level[1].w0 == G::TBA) ? kFlipWin : P.tiles_log2; // else: one chunk = the whole state
    if (P.chunk_log2 > P.tiles_log2) P.chunk_log2 = P.tiles_log2;
    P.n_items = (unsigned long long)nl << P.tiles_log2;
    return true;
}

// diag(C0 ^ w) = diag(C0) + sum over the boundary window bits set in w of (diag(C0 | bit) - diag(C0)) + T[w]: both
// Hamiltonians are quadratic forms in the bits with nearest-neighbour couplings only, so the couplings between window
// bits and the single-bit terms of the interior window bits do not depend on C0.  Window of the diagonal stage (level 0,
// stage 2): unit bits 8..11 (+ the in-unit amplitude bit for ComplexF32); boundary bits: unit 8, unit 11 (+ amplitude
// bit 0).  T is indexed by the register index r (bit 0 = in-unit bit for ComplexF32, then unit bits 8..11).
template <typename R> inline void flip_fill_diag_table(FlipParams &P)
{
    using G = FlipGeom<R>;
    const HamEval Hn{P.H.kind, P.H.n, P.H.p0, P.H.p1, P.H.p2};
    const double d0 = Hn.diag(0);
    const unsigned long long lo = 1ull << (8 + G::SUB), hi = 1ull << (11 + G::SUB);
    for (int r = 0; r < (16 << G::SUB); r++) {
        unsigned long long w = 0;
        if (G::SUB && (r & 1)) w |= 1ull;
        for (int b = 0; b < 4; b++)
            if ((r >> (b + G::SUB)) & 1) w |= 1ull << (8 + b + G::SUB);
        double t = Hn.diag(w) - d0;
        if (w & lo) t -= Hn.diag(lo) - d0;
        if (w & hi) t -= Hn.diag(hi) - d0;
        if (G::SUB && (w & 1ull)) t -= Hn.diag(1ull) - d0;
        P.diag_T[r] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// one register stage of one consumer thread on one tile
// ---------------------------------------------------------------------------------------------------------------------
template <typename R> struct FlipAmp;
template <> struct FlipAmp<double> { using V = double2; using Acc = double; };
template <> struct FlipAmp<float> { using V = float2; using Acc = float; };

// swizzled byte offset of the thread's window origin (window unit bits zero) -- depends on (lv, s, tid) only
template <int LV, int S> QCB_HD uint32_t flip_thread_unit(uint32_t tid)
{
    uint32_t u0 = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) u0 |= ((tid >> k) & 1u) << flip_tpos(LV, S, k);
    return u0;
}

template <typename R, int LV, int S>
QCB_HD void flip_load_window(const unsigned char *tile, uint32_t u0, typename FlipAmp<R>::V (&a)[16 << FlipGeom<R>::SUB])
{
    using G = FlipGeom<R>;
    const uint32_t off0 = tma_swz_unit(u0) << 4;
#pragma unroll
    for (int ru = 0; ru < 16; ru++) {
        uint32_t wu = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) wu |= (uint32_t)((ru >> b) & 1) << flip_win(LV, S, b);
        const uint32_t woff = tma_swz_unit(wu) << 4; // the swizzle is XOR-linear: swz(u0 | wu) = swz(u0) ^ swz(wu)
        if constexpr (G::NA == 1) {
            a[ru] = *reinterpret_cast<const double2 *>(tile + (off0 ^ woff));
        } else {
            const float4 v = *reinterpret_cast<const float4 *>(tile + (off0 ^ woff));
            a[2 * ru].x = v.x; a[2 * ru].y = v.y;         // register bit 0 = the in-unit amplitude bit
            a[2 * ru + 1].x = v.z; a[2 * ru + 1].y = v.w;
        }
    }
}

// runtime context of a tile for one thread
struct FlipTileCtx {
    unsigned long long base;  // first amplitude index of the tile | c_hi
    uint32_t active;          // unit bits whose flips count
    int tile_ctl_ok;          // East, level >= 1: the control qubit of the lowest window bit (index bit w0 - 1, outside the tile) is 0
};

// sum over the allowed pairs of window register bit JB of Re(conj(a[r | bit]) a[r]); CTL = register bit that holds the
// East control qubit (-1: none in the window)
template <typename R, int NR, int JB, int CTL>
QCB_HD typename FlipAmp<R>::Acc flip_pair_sum(const typename FlipAmp<R>::V (&a)[NR])
{
    typename FlipAmp<R>::Acc s0 = 0, s1 = 0;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        if ((r >> JB) & 1) continue;
        if (CTL >= 0 && ((r >> CTL) & 1)) continue;
        const typename FlipAmp<R>::V x0 = a[r], x1 = a[r | (1 << JB)];
        s0 = fma(x0.x, x1.x, s0);
        s1 = fma(x0.y, x1.y, s1);
    }
    return s0 + s1;
}

// flips of window index B (unit bit flip_win(LV, S, B)) of this stage
template <typename R, int KIND, int LV, int S, int B>
QCB_HD void flip_window_bit(const typename FlipAmp<R>::V (&a)[16 << FlipGeom<R>::SUB], uint32_t u0, const FlipTileCtx &T, double &acc)
{
    using G = FlipGeom<R>;
    constexpr int NR = 16 << G::SUB;
    constexpr int U = flip_win(LV, S, B);  // unit bit
    constexpr int JB = B + G::SUB;          // register bit
    if constexpr (LV == 1 && U < 3) {
        return; // the 128-byte row bits of a window tile were covered by level 0
    } else {
        if (!((T.active >> U) & 1u)) return;
        if constexpr (KIND == 0) { // TFIM: every pair counts
            acc += 2.0 * (double)flip_pair_sum<R, NR, JB, -1>(a);
        } else if constexpr (LV == 0 && U == 0 && G::SUB == 0) { // East: the flip of index bit i counts when bit i - 1 is 0; bit 0: no control
            acc += 2.0 * (double)flip_pair_sum<R, NR, JB, -1>(a);
        } else if constexpr (LV == 0 && U == 0) { // ComplexF32: unit bit 0 is index bit 1, its control is the in-unit bit (register bit 0)
            acc += 2.0 * (double)flip_pair_sum<R, NR, JB, 0>(a);
        } else if constexpr (LV == 1 && U == 3) { // lowest window bit: control is index bit w0 - 1, outside the tile
            if (T.tile_ctl_ok) acc += 2.0 * (double)flip_pair_sum<R, NR, JB, -1>(a);
        } else {
            constexpr int UC = U - 1; // the control is the unit bit below
            constexpr int BC = flip_win_index(LV, S, UC);
            if constexpr (BC >= 0) {
                acc += 2.0 * (double)flip_pair_sum<R, NR, JB, BC + G::SUB>(a);
            } else {
                const typename FlipAmp<R>::Acc s = flip_pair_sum<R, NR, JB, -1>(a);
                if (!((u0 >> UC) & 1u)) acc += 2.0 * (double)s;
            }
        }
    }
}

template <typename R, int KIND, int LV, int S>
QCB_HD void flip_stage(const unsigned char *tile, uint32_t tid, const FlipParams &P, const FlipTileCtx &T, double &acc_diag, double &acc_flip)
{
    using G = FlipGeom<R>;
    using V = typename FlipAmp<R>::V;
    constexpr int NR = 16 << G::SUB;
    // skip a stage none of whose window bits is active (upper levels of small states)
    constexpr uint32_t wmask = (1u << flip_win(LV, S, 0)) | (1u << flip_win(LV, S, 1)) | (1u << flip_win(LV, S, 2)) | (1u << flip_win(LV, S, 3));
    if (!(LV == 0) && !(T.active & wmask)) return;
    const uint32_t u0 = flip_thread_unit<LV, S>(tid);
    V a[NR];
    flip_load_window<R, LV, S>(tile, u0, a);
    flip_window_bit<R, KIND, LV, S, 0>(a, u0, T, acc_flip);
    flip_window_bit<R, KIND, LV, S, 1>(a, u0, T, acc_flip);
    flip_window_bit<R, KIND, LV, S, 2>(a, u0, T, acc_flip);
    flip_window_bit<R, KIND, LV, S, 3>(a, u0, T, acc_flip);
    if constexpr (LV == 0 && S == 2) {
        if constexpr (G::SUB == 1) acc_flip += 2.0 * (double)flip_pair_sum<R, NR, 0, -1>(a); // ComplexF32: index bit 0 lives inside the unit (no control)
        // diagonal term: window = unit bits 8..11 (+ in-unit bit); boundary bits: unit 8, unit 11 (+ index bit 0)
        const unsigned long long C0 = T.base | ((unsigned long long)u0 << G::SUB);
        const unsigned long long lo = 1ull << (8 + G::SUB), hi = 1ull << (11 + G::SUB);
        const double d0 = P.H.diag(C0);
        const double dlo = P.H.diag(C0 | lo) - d0, dhi = P.H.diag(C0 | hi) - d0;
        const double dsub = G::SUB ? P.H.diag(C0 | 1ull) - d0 : 0.0;
        typename FlipAmp<R>::Acc part = 0;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            double d = d0 + P.diag_T[r];
            if ((r >> G::SUB) & 1) d += dlo;
            if ((r >> (3 + G::SUB)) & 1) d += dhi;
            if (G::SUB && (r & 1)) d += dsub;
            const typename FlipAmp<R>::Acc w = fma(a[r].x, a[r].x, a[r].y * a[r].y);
            part = fma((typename FlipAmp<R>::Acc)d, w, part);
        }
        acc_diag += (double)part;
    }
}

template <typename R, int KIND>
QCB_HD void flip_tile_consume(const unsigned char *tile, uint32_t tid, const FlipParams &P, const FlipItem &it, double &acc_diag, double &acc_flip)
{
    FlipTileCtx T;
    T.base = flip_tile_base<R>(P, it) | P.c_hi;
    T.active = P.level[it.lv].active;
    T.tile_ctl_ok = 1;
    if (it.lv == 0) {
        flip_stage<R, KIND, 0, 0>(tile, tid, P, T, acc_diag, acc_flip);
        flip_stage<R, KIND, 0, 1>(tile, tid, P, T, acc_diag, acc_flip);
        flip_stage<R, KIND, 0, 2>(tile, tid, P, T, acc_diag, acc_flip);
    } else {
        if (KIND == 1) T.tile_ctl_ok = ((T.base >> (P.level[it.lv].w0 - 1)) & 1ull) ? 0 : 1;
        flip_stage<R, KIND, 1, 0>(tile, tid, P, T, acc_diag, acc_flip);
        flip_stage<R, KIND, 1, 1>(tile, tid, P, T, acc_diag, acc_flip);
        flip_stage<R, KIND, 1, 2>(tile, tid, P, T, acc_diag, acc_flip);
    }
}

// CPU replay helper (tests/emu): what the two TMA loads of a tile put into shared memory
template <typename R>
inline void flip_emulate_tile_load(const typename Vec2<R>::type *psi, const FlipParams &P, const FlipItem &it, unsigned char *tile)
{
    using G = FlipGeom<R>;
    using V = typename Vec2<R>::type;
    const unsigned long long base = flip_tile_base<R>(P, it);
    for (uint32_t u = 0; u < (1u << kFlipUnitBits); u++) {
        // unit u = (row << 3) | chunk; row = window value (level >= 1) or consecutive 128-byte rows (level 0)
        unsigned long long amp;
        if (it.lv == 0) amp = base + ((unsigned long long)u << G::SUB);
        else amp = base | ((unsigned long long)(u & 7u) << G::SUB) | ((unsigned long long)(u >> 3) << P.level[it.lv].w0);
        V *dst = reinterpret_cast<V *>(tile + ((size_t)tma_swz_unit(u) << 4));
        for (int k = 0; k < G::NA; k++) dst[k] = psi[amp + k];
    }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------------------
// the kernel: one CTA per SM, 16 warps = two consumer groups that also issue their own TMA refills
// ---------------------------------------------------------------------------------------------------------------------
struct FlipMaps {
    CUtensorMap m[kFlipMaxLevels];
};

template <typename R> __device__ __forceinline__ void flip_issue_tile(const FlipMaps &M, const FlipParams &P, const FlipItem &it, unsigned char *dst, uint64_t *bar)
{
    tma::mbar_expect_tx(bar, kFlipTileBytes);
    if (it.lv == 0) {
        const int row0 = (int)(it.t << (kFlipUnitBits - 3));
        tma::load_2d(dst, &M.m[0], bar, 0, row0);
        tma::load_2d(dst + kFlipTileBytes / 2, &M.m[0], bar, 0, row0 + 256);
    } else {
        const FlipLevel &L = P.level[it.lv];
        const int mid = (int)(it.t & ((1ull << L.mid_bits) - 1ull)), hi = (int)(it.t >> L.mid_bits);
        tma::load_4d(dst, &M.m[it.lv], bar, 0, mid, 0, hi);
        tma::load_4d(dst + kFlipTileBytes / 2, &M.m[it.lv], bar, 0, mid, 256, hi);
    }
}

// Two consumer groups of 8 warps, no producer warp: with one group the kernel was latency bound (ncu at 28 qubits: issue
// slots 41 % busy, top stalls `wait` and `long_scoreboard`, DRAM at 62 % of its peak with two reads of the state), and a 17th
// warp would cost registers (a CTA's warps are allocated in fours: 20 x 32 x 96).  16 warps x 32 x 128 registers fill the SM.
// (Also tried: all 16 warps on ONE tile, two threads per register window sharing its work, so that two slots are in flight --
// 2.82 ms instead of 2.11 ms at 28 qubits: +36 % instructions for the redundant window loads and a CTA-wide barrier per tile.)
// Tile k of the CTA lives in ring slot k mod 3 and is consumed by group k mod 2; when a group has finished a tile (group
// barrier), its first thread refills that very slot with tile k + 3 -- which the OTHER group will consume -- so a slot is
// never written while it is read and no "empty" barriers are needed: one `full` mbarrier per slot, plus a plain counter
// `issued[slot]` (round of the last tile issued into the slot, + 1).  The counter is what makes the parity wait sound: a
// parity test cannot tell "phase r has completed" from "phase r - 1 has not completed yet", and the group that waits for
// round r of a slot did not consume round r - 1 of it.  Once it has seen issued[slot] == r + 1 it knows that the other
// group has passed its own wait for round r - 1 (it issues only after consuming), so the barrier is in phase r.
template <typename R, int KIND>
__global__ void __launch_bounds__(kFlipGroups * kFlipConsumers, 1)
flip_measure_kernel(const __grid_constant__ FlipMaps M, const __grid_constant__ FlipParams P, double2 *partial /* [gridDim.x] */)
{
    extern __shared__ unsigned char flip_smem_raw[];
    // (pointer arithmetic on the shared array itself: a round trip through uintptr_t makes every tile access a generic LD / ST
    // with a 64-bit address instead of LDS / STS)
    unsigned char *ring = flip_smem_raw + ((1024u - (tma::smem_u32(flip_smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + kFlipRing * kFlipTileBytes);
    volatile uint32_t *issued = reinterpret_cast<volatile uint32_t *>(full + kFlipRing); // [kFlipRing], inside the 2 * kFlipRing words reserved for barriers
    double *red = reinterpret_cast<double *>(full + 2 * kFlipRing); // [16 warps][2]
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t kWarpsPerGroup = kFlipConsumers / 32, kConsumerWarps = kFlipGroups * kWarpsPerGroup;
    if (tid == 0) {
        for (int i = 0; i < kFlipRing; i++) { tma::mbar_init(&full[i], 1); issued[i] = 0; }
        tma::fence_barrier_init();
        for (int l = 0; l < P.nlevels; l++) tma::prefetch_map(&M.m[l]);
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t k = 0; k < (uint32_t)kFlipRing; k++) {
            const unsigned long long j = blockIdx.x + (unsigned long long)k * gridDim.x;
            if (j < P.n_items) {
                flip_issue_tile<R>(M, P, flip_decode(P, j), ring + (size_t)k * kFlipTileBytes, &full[k]);
                __threadfence_block();
                issued[k] = 1;
            }
        }
    }
    double acc_diag = 0.0, acc_flip = 0.0;
    const uint32_t group = warp / kWarpsPerGroup, gtid = tid % kFlipConsumers;
    uint32_t k = group;
    for (unsigned long long j = blockIdx.x + (unsigned long long)group * gridDim.x; j < P.n_items; j += (unsigned long long)kFlipGroups * gridDim.x, k += kFlipGroups) {
        const uint32_t slot = k % kFlipRing, round = k / kFlipRing;
        while (issued[slot] < round + 1u) { } // the tile of this round has been issued (=> the previous round of the slot has landed)
        __threadfence_block();
        tma::mbar_wait(&full[slot], round & 1u);
        flip_tile_consume<R, KIND>(ring + (size_t)slot * kFlipTileBytes, gtid, P, flip_decode(P, j), acc_diag, acc_flip);
        tma::named_bar_sync(1 + (int)group, kFlipConsumers); // every warp of the group is done with the slot
        if (gtid == 0) {
            const unsigned long long jn = j + (unsigned long long)kFlipRing * gridDim.x; // tile k + 3 of this CTA
            if (jn < P.n_items) {
                tma::fence_proxy_async(); // the group's reads of the slot come before the bulk copy that overwrites it
                flip_issue_tile<R>(M, P, flip_decode(P, jn), ring + (size_t)slot * kFlipTileBytes, &full[slot]);
                __threadfence_block();
                issued[slot] = round + 2u;
            }
        }
    }
    acc_diag = warp_sum(acc_diag);
    acc_flip = warp_sum(acc_flip);
    if (lane == 0) { red[2 * warp] = acc_diag; red[2 * warp + 1] = acc_flip; }
    __syncthreads();
    if (tid == 0) {
        double d = 0, f = 0;
        for (uint32_t w = 0; w < kConsumerWarps; w++) { d += red[2 * w]; f += red[2 * w + 1]; } // fixed order: reproducible
        partial[blockIdx.x] = make_double2(d, f);
    }
}

inline size_t flip_smem_bytes() { return (size_t)kFlipRing * kFlipTileBytes + 1024 + 2 * kFlipRing * sizeof(uint64_t) + 2 * kFlipGroups * (kFlipConsumers / 32) * sizeof(double); }

// host: tensor maps of all levels for a state at `base`
template <typename R> inline int flip_make_maps(FlipMaps &M, const void *base, const FlipParams &P, const char **why)
{
    using G = FlipGeom<R>;
    const bool f64 = sizeof(R) == 8;
    const unsigned inner = f64 ? 16u : 32u; // elements per 128-byte row
    {
        TmaDims d;
        d.rank = 2;
        d.dims[0] = inner; d.dims[1] = 1ull << (P.n - G::C);
        d.strides[0] = 128;
        d.box[0] = inner; d.box[1] = 256;
        if (tma_encode(&M.m[0], base, f64, d, why)) return 1;
    }
    for (int l = 1; l < P.nlevels; l++) {
        const FlipLevel &L = P.level[l];
        TmaDims d;
        d.rank = 4;
        d.dims[0] = inner;
        d.dims[1] = 1ull << L.mid_bits;                       d.strides[0] = 128;
        d.dims[2] = 1ull << kFlipWin;                          d.strides[1] = (unsigned long long)sizeof(R) * 2ull << L.w0;
        d.dims[3] = 1ull << (P.n - L.w0 - kFlipWin);           d.strides[2] = (unsigned long long)sizeof(R) * 2ull << (L.w0 + kFlipWin);
        d.box[0] = inner; d.box[1] = 1; d.box[2] = 256; d.box[3] = 1;
        d.strided_rows = true;
        if (tma_encode(&M.m[l], base, f64, d, why)) return 1;
    }
    return 0;
}
#endif

} // namespace qcb
