// adjoint_f32.cuh -- fused backward pass of the adjoint gradient sweep for ComplexF32 states.
//
// Same contract as adjoint_pass_kernel (adjoint.cuh; it replaces the reference's 15 G (G+1) re-propagations,
// src/gradients.jl:107-139): one launch un-applies a trapezoid of gates of the reversed circuit on a psi tile and a
// lambda tile and reduces, per gate, env[r, c] = sum_rest conj(lambda_g[r, rest]) psi_{g-1}[c, rest].
//
// What adjoint.cuh does for ComplexF32 and why it is slow: one amplitude quad per thread and three shared-memory round
// trips per gate (psi, the operand fetch of the reduction, lambda), and the reduction itself on the FP64 tensor cores
// after converting every operand to double (F2F + DMMA at the FP64 rate for a single-precision state).
//
// Here everything between two stage boundaries happens in registers:
//   * the 16-amplitude register windows of BOTH states stay in registers for all (up to 4) gates of a stage, exactly as
//     in the forward pass (tile.cuh); a gate is un-applied with packed FFMA2;
//   * the thread accumulates the 4x4 complex outer products conj(lambda_r) psi_c of its own 4 quads with packed FFMA2
//     (32 FFMA2 per quad, the same issue cost as one un-application; the operands are the amplitudes as they sit in the
//     registers -- see adjf_outer);
//   * per gate and tile, the warp sums its 32 x 32 partial values with two transpose-reductions of 16 values (32 shuffles
//     and adds per lane instead of 160 for 32 butterfly reductions): every lane ends up with one component of the warp's
//     sum, which it adds to the double accumulators in shared memory (one row per gate and warp, as in adjoint.cuh).
// Shared memory is touched once per stage (load both windows, store both windows); nothing runs on the FP64 pipe
// except 32 DADDs per warp, gate and tile.  FP32 accumulation covers the 128 quads a warp owns in one tile only.
//
// A variant that fed the reduction to the single-precision tensor cores (mma.sync.m16n8k8 TF32 with a hi/lo operand
// split to keep FP32 accuracy: one MMA = 4 quads x all four hi/lo cross terms) was built first and measured slower than
// adjoint.cuh (28 q gradient 81.9 ms vs 66.2 ms): ~20 issue slots of operand preparation per MMA, 672 of the 1225
// instructions per warp and gate, instruction-cache misses on top (ncu: no_instruction the top stall, ALU pipe 42 %,
// tensor pipe 12 %), and the legacy MMA path takes issue bandwidth from FFMA2 (tools/fp64_peak.cu: FFMA2 drops from
// 62.5 to 39.2 TFLOP/s with one TF32 MMA per 8 FFMA2).  profiles/r02_adjoint_f32_tf32_variant_ncu.txt keeps the numbers.
//
// The index algebra is __host__ __device__ and replayed on the CPU by tests/emu (adjoint_stage_f32_host).
#pragma once
#include <cstring>

#include "adjoint.cuh"

namespace qcb {

// the two bits of q spread over the window bits other than LO, LO+1 (as apply_window_gate does)
QCB_HD constexpr int adjf_other(int LO, int q) { return LO == 0 ? (q << 2) : (LO == 1 ? ((q & 1) | ((q & 2) << 2)) : q); }

// Outer products conj(l_r) x_c for the columns c = C0, C0 + 1 over the 4 quads of a window, gate on window bits (LO, LO+1).
// No operand is broadcast or negated: with P = (l.re x.re, l.im x.im) and Q = (l.re x.im, l.im x.re) accumulated as packed
// pairs (two FFMA2 whose operands are the amplitudes as they sit in the registers; the swap is an operand modifier),
//     Re conj(l) x = P.lo + P.hi,     Im conj(l) x = Q.lo - Q.hi
// are two scalar adds per matrix element, once per gate instead of once per quad.  accP / accQ index: r + 4 (c - C0).
template <int LO, int C0> QCB_HD void adjf_outer(const float2 (&l)[kRegAmps], const float2 (&x)[kRegAmps], float2 (&accP)[8], float2 (&accQ)[8])
{
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int other = adjf_other(LO, q);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const float2 xv = x[other | ((C0 + c) << LO)];
            const float2 xs = float2{xv.y, xv.x};
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const float2 lv = l[other | (r << LO)];
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
                accP[r + 4 * c] = __ffma2_rn(lv, xv, accP[r + 4 * c]);
                accQ[r + 4 * c] = __ffma2_rn(lv, xs, accQ[r + 4 * c]);
#else
                accP[r + 4 * c].x = fmaf(lv.x, xv.x, accP[r + 4 * c].x); accP[r + 4 * c].y = fmaf(lv.y, xv.y, accP[r + 4 * c].y);
                accQ[r + 4 * c].x = fmaf(lv.x, xs.x, accQ[r + 4 * c].x); accQ[r + 4 * c].y = fmaf(lv.y, xs.y, accQ[r + 4 * c].y);
#endif
            }
        }
    }
}
// v[2 k] = Re, v[2 k + 1] = Im, k = r + 4 (c - C0): half an environment row (adjoint.cuh: env_accumulate)
QCB_HD void adjf_combine(const float2 (&accP)[8], const float2 (&accQ)[8], float (&v)[16])
{
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[2 * k] = accP[k].x + accP[k].y;
        v[2 * k + 1] = accQ[k].x - accQ[k].y;
    }
}

// Transpose-reduction of 16 values per lane over a warp: step s = 16, 8, 4, 2 halves the number of live components (a lane
// whose bit s is set keeps the upper half and sends the lower half to lane ^ s, and vice versa), a last plain exchange
// with lane ^ 1 completes the sum: lanes 2i and 2i + 1 both end up with component i of the warp's sum in v[0].
// (16 shuffles and adds per lane and round instead of 80 for 16 butterfly reductions.)  Host form: all 32 lanes at once.
inline void adjf_transpose_reduce_host(float (*v)[16] /* [32 lanes][16] */)
{
    for (int s = 16, n = 16; s >= 2; s >>= 1, n >>= 1) {
        float nv[32][16];
        for (int lane = 0; lane < 32; lane++) {
            const bool up = (lane & s) != 0;
            for (int i = 0; i < n / 2; i++) {
                const float keep = up ? v[lane][i + n / 2] : v[lane][i];
                const float recv = up ? v[lane ^ s][i + n / 2] : v[lane ^ s][i]; // the partner sends its half of the kind I keep
                nv[lane][i] = keep + recv;
            }
        }
        for (int lane = 0; lane < 32; lane++)
            for (int i = 0; i < n / 2; i++) v[lane][i] = nv[lane][i];
    }
    float last[32];
    for (int lane = 0; lane < 32; lane++) last[lane] = v[lane][0] + v[lane ^ 1][0];
    for (int lane = 0; lane < 32; lane++) v[lane][0] = last[lane];
}
// where a lane adds its value after the two rounds: even lanes own round 0 (columns 0, 1), odd lanes round 1 (columns 2, 3)
QCB_HD int adjf_acc_index(uint32_t lane) { return (int)((lane & 1u) * 16u + (lane >> 1)); }

// swizzled tile-local offset of thread `geom`'s window origin
template <int TB> QCB_HD uint32_t adjf_win_base(const StageDesc &S, uint32_t geom)
{
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TB - kRegBits; k++) e |= ((geom >> k) & 1u) << S.tpos[k];
    return swz<4>(e);
}

// CPU replay of one stage of one tile: warp by warp with the device's arithmetic.  acc: [warps][kAdjMaxGates][32] doubles
template <int TB>
inline void adjoint_stage_f32_host(const AdjParams<float> &A, int s, float2 *smP, float2 *smL, double *acc)
{
    constexpr int NT = 1 << (TB - kRegBits);
    static_assert(NT >= 32, "a warp of windows");
    const StageDesc &S = A.P.stage[s];
    for (int warp = 0; warp < NT / 32; warp++) {
        float2 x[32][kRegAmps], l[32][kRegAmps];
        uint32_t e[32];
        for (uint32_t lane = 0; lane < 32; lane++) {
            e[lane] = adjf_win_base<TB>(S, (uint32_t)warp * 32u + lane);
            for (int r = 0; r < kRegAmps; r++) { x[lane][r] = smP[e[lane] ^ S.roff[r]]; l[lane][r] = smL[e[lane] ^ S.roff[r]]; }
        }
        for (int slot = 0; slot < kSlots; slot++) {
            if (!((S.mask >> slot) & 1u)) continue;
            const int LO = slot_lo(slot);
            const GateMat<float> &Ud = A.P.U[s * kSlots + slot];
            double *acc_gate = acc + ((size_t)warp * kAdjMaxGates + A.slot_gate[s * kSlots + slot]) * 32;
            float v[2][32][16];
            for (uint32_t lane = 0; lane < 32; lane++) {
                float2 p0[8], q0[8], p1[8], q1[8];
                for (int i = 0; i < 8; i++) p0[i] = q0[i] = p1[i] = q1[i] = float2{0.f, 0.f};
                if (LO == 0) { apply_window_gate<0, float, float2>(x[lane], Ud); adjf_outer<0, 0>(l[lane], x[lane], p0, q0); adjf_outer<0, 2>(l[lane], x[lane], p1, q1); apply_window_gate<0, float, float2>(l[lane], Ud); }
                else if (LO == 1) { apply_window_gate<1, float, float2>(x[lane], Ud); adjf_outer<1, 0>(l[lane], x[lane], p0, q0); adjf_outer<1, 2>(l[lane], x[lane], p1, q1); apply_window_gate<1, float, float2>(l[lane], Ud); }
                else { apply_window_gate<2, float, float2>(x[lane], Ud); adjf_outer<2, 0>(l[lane], x[lane], p0, q0); adjf_outer<2, 2>(l[lane], x[lane], p1, q1); apply_window_gate<2, float, float2>(l[lane], Ud); }
                adjf_combine(p0, q0, v[0][lane]);
                adjf_combine(p1, q1, v[1][lane]);
            }
            adjf_transpose_reduce_host(v[0]);
            adjf_transpose_reduce_host(v[1]);
            for (uint32_t lane = 0; lane < 32; lane++) acc_gate[adjf_acc_index(lane)] += (double)v[lane & 1u][lane][0];
        }
        for (uint32_t lane = 0; lane < 32; lane++)
            for (int r = 0; r < kRegAmps; r++) { smP[e[lane] ^ S.roff[r]] = x[lane][r]; smL[e[lane] ^ S.roff[r]] = l[lane][r]; }
    }
}

#if defined(__CUDACC__)
// one round of the reduction: columns C0, C0 + 1 of the outer products, transpose-reduced over the warp
template <int LO, int C0>
__device__ __forceinline__ float adjf_round(const float2 (&l)[kRegAmps], const float2 (&x)[kRegAmps], uint32_t lane)
{
    float2 accP[8], accQ[8];
#pragma unroll
    for (int i = 0; i < 8; i++) accP[i] = accQ[i] = make_float2(0.f, 0.f);
    adjf_outer<LO, C0>(l, x, accP, accQ);
    float v[16];
    adjf_combine(accP, accQ, v);
#pragma unroll
    for (int s = 16, n = 16; s >= 2; s >>= 1, n >>= 1) {
        const bool up = (lane & (uint32_t)s) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; i++) {
            const float keep = up ? v[i + n / 2] : v[i];
            const float send = up ? v[i] : v[i + n / 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// one gate slot of a stage: x, l = the thread's psi / lambda windows (registers)
template <int LO>
__device__ __forceinline__ void adjf_slot(const GateMat<float> &Ud, float2 (&x)[kRegAmps], float2 (&l)[kRegAmps], uint32_t lane, double *acc_gate)
{
    apply_window_gate<LO, float, float2>(x, Ud); // psi_{g-1} = U^dagger psi_g
    const float r0 = adjf_round<LO, 0>(l, x, lane); // conj(lambda_g) psi_{g-1}^T, summed over the warp's 128 quads
    const float r1 = adjf_round<LO, 2>(l, x, lane);
    apply_window_gate<LO, float, float2>(l, Ud); // lambda_{g-1} = U^dagger lambda_g
    acc_gate[adjf_acc_index(lane)] += (double)((lane & 1u) ? r1 : r0);
}

template <int TB>
__device__ __forceinline__ void adjoint_stage_f32(const AdjParams<float> &A, int s, float2 *smP, float2 *smL, double *acc_warp, uint32_t tid)
{
    const StageDesc &S = A.P.stage[s];
    const uint32_t lane = tid & 31u;
    const uint32_t e = adjf_win_base<TB>(S, tid);
    float2 x[kRegAmps], l[kRegAmps];
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) { x[r] = smP[e ^ S.roff[r]]; l[r] = smL[e ^ S.roff[r]]; }
    const uint32_t mask = S.mask;
    const uint8_t *sg = A.slot_gate + s * kSlots;
    // slots 0 and 3 act on the same window bits (1, 2): one copy of that code, entered twice (the kernel is bound by the FMA
    // pipe with 16 warps per SM; ncu showed instruction-fetch stalls on the fully unrolled form)
#pragma unroll 1
    for (int rep = 0; rep < 2; rep++) {
        const int sl = rep ? 3 : 0;
        if ((mask >> sl) & 1u) adjf_slot<1>(A.P.U[s * kSlots + sl], x, l, lane, acc_warp + 32 * sg[sl]);
        if (rep == 0) {
            if (mask & 2u) adjf_slot<0>(A.P.U[s * kSlots + 1], x, l, lane, acc_warp + 32 * sg[1]);
            if (mask & 4u) adjf_slot<2>(A.P.U[s * kSlots + 2], x, l, lane, acc_warp + 32 * sg[2]);
        }
    }
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) { smP[e ^ S.roff[r]] = x[r]; smL[e ^ S.roff[r]] = l[r]; }
}

// tile_load (tile.cuh) in two halves, so that the next tile's global loads are in flight during the write-back of the current one
template <int TB>
__device__ __forceinline__ void adjf_fetch(const float2 *__restrict__ src, const PassParams<float> &P, uint32_t tid, unsigned long long bid, float2 (&tmp)[kRegAmps])
{
    unsigned long long g = cta_base(P, bid);
#pragma unroll
    for (int k = 0; k < TB - kRegBits; k++) g |= (unsigned long long)((tid >> k) & 1u) << P.ld_phys[k];
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) tmp[i] = src[g | P.ld_goff[i]];
}
template <int TB>
__device__ __forceinline__ void adjf_fill(const PassParams<float> &P, float2 *sm, uint32_t tid, const float2 (&tmp)[kRegAmps])
{
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TB - kRegBits; k++) e |= ((tid >> k) & 1u) << P.ld_lpos[k];
    e = swz<4>(e);
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) sm[e ^ P.ld_eoff[i]] = tmp[i];
}

// persistent CTAs; partial: [gridDim.x][kAdjMaxGates][32] (the layout adjoint_pass_kernel writes: k_env_final sums it)
template <int TB, int MINB>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
adjoint_pass_f32_kernel(float2 *psi, float2 *lam, const __grid_constant__ AdjParams<float> A, unsigned long long ntiles, double *partial,
                        const __grid_constant__ EnvIds ids, int *id_row /* may be null: [kAdjMaxGates] gate index of every accumulator row */)
{
    if (id_row && blockIdx.x == 0 && threadIdx.x < kAdjMaxGates) id_row[threadIdx.x] = (int)threadIdx.x < ids.n ? ids.id[threadIdx.x] : -1;
    constexpr int NT = 1 << (TB - kRegBits);
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    float2 *smP = reinterpret_cast<float2 *>(qcb_smem_raw);
    float2 *smL = smP + (1 << TB);
    double *acc = reinterpret_cast<double *>(smL + (1 << TB)); // [NW][kAdjMaxGates][32]
    const uint32_t tid = threadIdx.x;
    for (int i = tid; i < NW * kAdjMaxGates * 32; i += NT) acc[i] = 0.0;
    __syncthreads();
    double *acc_warp = acc + (tid >> 5) * (kAdjMaxGates * 32);
    unsigned long long tile = blockIdx.x; // (the host launches at most ntiles CTAs)
    {
        float2 rp[kRegAmps], rl[kRegAmps];
        adjf_fetch<TB>(psi, A.P, tid, tile, rp);
        adjf_fetch<TB>(lam, A.P, tid, tile, rl);
        adjf_fill<TB>(A.P, smP, tid, rp);
        adjf_fill<TB>(A.P, smL, tid, rl);
    }
    while (true) {
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < A.P.nstages; s++) {
            adjoint_stage_f32<TB>(A, s, smP, smL, acc_warp, tid);
            __syncthreads();
        }
        const unsigned long long next = tile + gridDim.x;
        if (next >= ntiles) {
            tile_store<float, TB>(psi, A.P, smP, tid, tile);
            tile_store<float, TB>(lam, A.P, smL, tid, tile);
            break;
        }
        float2 rp[kRegAmps], rl[kRegAmps];
        adjf_fetch<TB>(psi, A.P, tid, next, rp); // in flight during the write-back
        adjf_fetch<TB>(lam, A.P, tid, next, rl);
        tile_store<float, TB>(psi, A.P, smP, tid, tile);
        tile_store<float, TB>(lam, A.P, smL, tid, tile);
        __syncthreads(); // every thread has read its part of the tile: the buffers may be overwritten
        adjf_fill<TB>(A.P, smP, tid, rp);
        adjf_fill<TB>(A.P, smL, tid, rl);
        tile = next;
    }
    __syncthreads();
    for (int i = tid; i < kAdjMaxGates * 32; i += NT) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) s += acc[w * (kAdjMaxGates * 32) + i];
        partial[(size_t)blockIdx.x * (kAdjMaxGates * 32) + i] = s;
    }
}
#endif

} // namespace qcb
