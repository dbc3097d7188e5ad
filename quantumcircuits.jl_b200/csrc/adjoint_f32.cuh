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
// Here:
//   * the 16-amplitude register windows of BOTH states stay in registers for all (up to 4) gates of a stage, exactly as
//     in the forward pass (tile.cuh); a gate is un-applied with packed FFMA2 on registers;
//   * after a gate, the thread writes its updated window back in place (the write the stage needs anyway at its end),
//     and the warp reduces its 128 quads with single-precision tensor-core MMAs (mma.sync.m16n8k8, TF32 inputs, FP32
//     accumulate) whose operands come straight from the tile: each element is read exactly once;
//   * single-precision accuracy is kept by the split x = hi + lo (hi = the upper 19 bits = an exact TF32 number, lo = x - hi
//     exact in FP32): ONE m16n8k8 MMA contracts 4 quads and produces all four hi/lo cross terms:
//         rows  m = (amplitude j of lambda, plain / "tilde", hi / lo)      16 = 4 x 2 x 2
//         k       = (quad q, Re / Im)                                        8 = 4 x 2
//         cols  n = (amplitude c of psi, hi / lo)                            8 = 4 x 2
//     A[(j, plain)][(q, Re / Im)] = (Re, Im) lambda_j,   A[(j, tilde)][..] = (-Im, Re) lambda_j,   B[(q, Re / Im)][c] = (Re, Im) psi_c
//     so that  D[(j, plain)][c] = Re conj(lambda_j) psi_c  and  D[(j, tilde)][c] = Im conj(lambda_j) psi_c  summed over the
//     quads.  Lane (g, t) loads ONE float2 of each state per MMA (amplitude g & 3 of quad t) and builds its 4 + 2 operand
//     registers from it; no shuffle, no double conversion, nothing on the FP64 pipe;
//   * FP32 accumulation runs over the 128 quads a warp owns in one tile only; the per-tile sums go to double accumulators
//     in shared memory (one row per gate and warp), as in adjoint.cuh.
//
// The index algebra is __host__ __device__ and replayed on the CPU by tests/emu with a software model of the m16n8k8
// fragment layout (adjoint_stage_f32_host).
#pragma once
#include <cstring>

#include "adjoint.cuh"

namespace qcb {

QCB_HD uint32_t adjf_bits(float x)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u;
    std::memcpy(&u, &x, 4);
    return u;
#endif
}
QCB_HD float adjf_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x;
    std::memcpy(&x, &u, 4);
    return x;
#endif
}
// upper 19 bits of x: exactly representable as TF32 (sign, 8 exponent bits, 10 mantissa bits)
QCB_HD float adjf_hi(float x) { return adjf_float(adjf_bits(x) & 0xffffe000u); }

// the two bits of q spread over the window bits other than LO, LO+1 (as apply_window_gate does)
QCB_HD constexpr int adjf_other(int LO, int q) { return LO == 0 ? (q << 2) : (LO == 1 ? ((q & 1) | ((q & 2) << 2)) : q); }

// operand registers of lane (g, t) for one reduction MMA, from lambda_j and psi_j (j = g & 3) of quad t
struct AdjfFrag {
    uint32_t a[4]; // A[g][t], A[g+8][t], A[g][t+4], A[g+8][t+4]:  (hi, lo) of the Re-slot value, (hi, lo) of the Im-slot value
    uint32_t b[2]; // B[t][g], B[t+4][g]
};
QCB_HD AdjfFrag adjf_operands(float2 lv, float2 pv, uint32_t lane)
{
    const bool upper = ((lane >> 4) & 1u) != 0; // g >= 4: the "tilde" rows of A, the lo columns of B
    const float u = upper ? -lv.y : lv.x, v = upper ? lv.x : lv.y;
    const float uh = adjf_hi(u), vh = adjf_hi(v);
    const float ph = adjf_hi(pv.x), qh = adjf_hi(pv.y);
    AdjfFrag F;
    F.a[0] = adjf_bits(uh);
    F.a[1] = adjf_bits(u - uh);
    F.a[2] = adjf_bits(vh);
    F.a[3] = adjf_bits(v - vh);
    F.b[0] = adjf_bits(upper ? pv.x - ph : ph);
    F.b[1] = adjf_bits(upper ? pv.y - qh : qh);
    return F;
}

// where lane (g, t) adds its two sums after the MMAs of a gate: s0 = C[g][2t] + C[g+8][2t], s1 = C[g][2t+1] + C[g+8][2t+1],
// each summed with lane t ^ 2 (the lo columns); lanes with t < 2 own env element (j = g & 3, c = 2t, 2t+1), Re (g < 4) or Im
QCB_HD int adjf_acc_index(uint32_t lane, int which)
{
    const uint32_t g = lane >> 2, t = lane & 3u;
    const int j = (int)(g & 3u), c = (int)(2u * t) + which;
    return 2 * (j + 4 * c) + (int)(g >> 2);
}

// swizzled tile-local offset of thread `geom`'s window origin (XOR-linear in geom)
template <int TB> QCB_HD uint32_t adjf_win_base(const StageDesc &S, uint32_t geom)
{
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TB - kRegBits; k++) e |= ((geom >> k) & 1u) << S.tpos[k];
    return swz<4>(e);
}
// swizzled offset of register element r of a window at p0 (= StageDesc::roff[r], recomputed: a per-lane index into the
// kernel-parameter bank would be a divergent constant fetch)
QCB_HD uint32_t adjf_roff(int p0, int r) { return swz<4>((uint32_t)r << p0); }

// ---- software model of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (CPU replay) ----------------------------------------
// a[lane][4], b[lane][2], c[lane][4]; lane = 4 g + t.  TF32 inputs: the low 13 bits are dropped.
inline void adjf_mma_host(const uint32_t (*a)[4], const uint32_t (*b)[2], float (*c)[4])
{
    float A[16][8], B[8][8];
    for (int lane = 0; lane < 32; lane++) {
        const int g = lane >> 2, t = lane & 3;
        A[g][t] = adjf_float(a[lane][0] & 0xffffe000u);
        A[g + 8][t] = adjf_float(a[lane][1] & 0xffffe000u);
        A[g][t + 4] = adjf_float(a[lane][2] & 0xffffe000u);
        A[g + 8][t + 4] = adjf_float(a[lane][3] & 0xffffe000u);
        B[t][g] = adjf_float(b[lane][0] & 0xffffe000u);
        B[t + 4][g] = adjf_float(b[lane][1] & 0xffffe000u);
    }
    for (int lane = 0; lane < 32; lane++) {
        const int g = lane >> 2, t = lane & 3;
        for (int i = 0; i < 4; i++) {
            const int m = g + 8 * (i >> 1), n = 2 * t + (i & 1);
            double s = c[lane][i];
            for (int k = 0; k < 8; k++) s += (double)A[m][k] * (double)B[k][n];
            c[lane][i] = (float)s;
        }
    }
}

template <int LO> inline void adjf_apply_window_host(float2 (&a)[kRegAmps], const GateMat<float> &U) { apply_window_gate<LO, float, float2>(a, U); }

// CPU replay of one stage of one tile: warp by warp, phase by phase, with the device's address algebra.
// acc: [warps][kAdjMaxGates][32] doubles
template <int TB>
inline void adjoint_stage_f32_host(const AdjParams<float> &A, int s, float2 *smP, float2 *smL, double *acc)
{
    constexpr int NT = 1 << (TB - kRegBits);
    static_assert(NT >= 32, "a warp of windows");
    const StageDesc &S = A.P.stage[s];
    for (int warp = 0; warp < NT / 32; warp++) {
        for (int slot = 0; slot < kSlots; slot++) {
            if (!((S.mask >> slot) & 1u)) continue;
            const int LO = slot_lo(slot);
            const GateMat<float> &Ud = A.P.U[s * kSlots + slot];
            double *acc_gate = acc + ((size_t)warp * kAdjMaxGates + A.slot_gate[s * kSlots + slot]) * 32;
            auto window = [&](float2 *sm, uint32_t lane) {
                const uint32_t e = adjf_win_base<TB>(S, (uint32_t)warp * 32u + lane);
                float2 w[kRegAmps];
                for (int r = 0; r < kRegAmps; r++) w[r] = sm[e ^ S.roff[r]];
                if (LO == 0) adjf_apply_window_host<0>(w, Ud);
                else if (LO == 1) adjf_apply_window_host<1>(w, Ud);
                else adjf_apply_window_host<2>(w, Ud);
                for (int r = 0; r < kRegAmps; r++) sm[e ^ S.roff[r]] = w[r];
            };
            for (uint32_t lane = 0; lane < 32; lane++) window(smP, lane);      // psi <- U^dagger psi
            float c[32][4];
            for (int lane = 0; lane < 32; lane++) c[lane][0] = c[lane][1] = c[lane][2] = c[lane][3] = 0.f;
            for (uint32_t d = 0; d < 32; d++) {                                  // MMA d contracts the 4 quads of lane d's window
                uint32_t a[32][4], b[32][2];
                const uint32_t ed = adjf_win_base<TB>(S, (uint32_t)warp * 32u + d);
                for (uint32_t lane = 0; lane < 32; lane++) {
                    const uint32_t g = lane >> 2, t = lane & 3u;
                    const uint32_t idx = ed ^ adjf_roff(S.p0, adjf_other(LO, (int)t) | (int)((g & 3u) << LO));
                    const AdjfFrag F = adjf_operands(smL[idx], smP[idx], lane);
                    for (int i = 0; i < 4; i++) a[lane][i] = F.a[i];
                    b[lane][0] = F.b[0]; b[lane][1] = F.b[1];
                }
                adjf_mma_host(a, b, c);
            }
            for (uint32_t lane = 0; lane < 32; lane++) {
                if ((lane & 3u) >= 2) continue;
                const float s0 = (c[lane][0] + c[lane][2]) + (c[lane ^ 2][0] + c[lane ^ 2][2]);
                const float s1 = (c[lane][1] + c[lane][3]) + (c[lane ^ 2][1] + c[lane ^ 2][3]);
                acc_gate[adjf_acc_index(lane, 0)] += (double)s0;
                acc_gate[adjf_acc_index(lane, 1)] += (double)s1;
            }
            for (uint32_t lane = 0; lane < 32; lane++) window(smL, lane);      // lambda <- U^dagger lambda
        }
    }
}

#if defined(__CUDACC__)
__device__ __forceinline__ void adjf_mma(float (&c)[4], const AdjfFrag &F)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(F.a[0]), "r"(F.a[1]), "r"(F.a[2]), "r"(F.a[3]), "r"(F.b[0]), "r"(F.b[1]));
}

// one gate slot of a stage: x, l = the thread's psi / lambda windows (registers), e = swizzled origin of its window
template <int TB, int LO>
__device__ __forceinline__ void adjf_slot(const StageDesc &S, const GateMat<float> &Ud, float2 (&x)[kRegAmps], float2 (&l)[kRegAmps], float2 *smP,
                                          float2 *smL, uint32_t e, uint32_t ewarp, const uint32_t (&cbit)[5], uint32_t lane, double *acc_gate)
{
    // (1) psi <- U^dagger psi on registers; write the window back in place
    apply_window_gate<LO, float, float2>(x, Ud);
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) smP[e ^ S.roff[r]] = x[r];
    __syncwarp();
    // (2) env += conj(lambda) psi^T over the warp's 32 windows: MMA d takes the 4 quads of lane d's window
    const uint32_t g = lane >> 2, t = lane & 3u;
    const uint32_t mine = ewarp ^ adjf_roff(S.p0, adjf_other(LO, (int)t) | (int)((g & 3u) << LO));
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 32; d++) {
        uint32_t ed = 0; // window origin of lane d relative to the warp's: XOR of the lane-bit constants (uniform)
#pragma unroll
        for (int k = 0; k < 5; k++) ed ^= ((d >> k) & 1) ? cbit[k] : 0u;
        const uint32_t idx = mine ^ ed;
        adjf_mma(c, adjf_operands(smL[idx], smP[idx], lane));
    }
    {
        float s0 = c[0] + c[2], s1 = c[1] + c[3];
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        if (t < 2u) {
            acc_gate[adjf_acc_index(lane, 0)] += (double)s0;
            acc_gate[adjf_acc_index(lane, 1)] += (double)s1;
        }
    }
    __syncwarp();
    // (3) lambda <- U^dagger lambda
    apply_window_gate<LO, float, float2>(l, Ud);
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) smL[e ^ S.roff[r]] = l[r];
}

template <int TB>
__device__ __forceinline__ void adjoint_stage_f32(const AdjParams<float> &A, int s, float2 *smP, float2 *smL, double *acc_warp, uint32_t tid)
{
    const StageDesc &S = A.P.stage[s];
    const uint32_t lane = tid & 31u;
    const uint32_t ewarp = adjf_win_base<TB>(S, tid & ~31u);
    uint32_t cbit[5];
#pragma unroll
    for (int k = 0; k < 5; k++) cbit[k] = swz<4>(1u << S.tpos[k]);
    uint32_t e = ewarp;
#pragma unroll
    for (int k = 0; k < 5; k++) e ^= ((lane >> k) & 1u) ? cbit[k] : 0u;
    float2 x[kRegAmps], l[kRegAmps];
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) { x[r] = smP[e ^ S.roff[r]]; l[r] = smL[e ^ S.roff[r]]; }
    const uint32_t mask = S.mask;
    const uint8_t *sg = A.slot_gate + s * kSlots;
    if (mask & 1u) adjf_slot<TB, 1>(S, A.P.U[s * kSlots + 0], x, l, smP, smL, e, ewarp, cbit, lane, acc_warp + 32 * sg[0]);
    if (mask & 2u) adjf_slot<TB, 0>(S, A.P.U[s * kSlots + 1], x, l, smP, smL, e, ewarp, cbit, lane, acc_warp + 32 * sg[1]);
    if (mask & 4u) adjf_slot<TB, 2>(S, A.P.U[s * kSlots + 2], x, l, smP, smL, e, ewarp, cbit, lane, acc_warp + 32 * sg[2]);
    if (mask & 8u) adjf_slot<TB, 1>(S, A.P.U[s * kSlots + 3], x, l, smP, smL, e, ewarp, cbit, lane, acc_warp + 32 * sg[3]);
}

// persistent CTAs; partial: [gridDim.x][kAdjMaxGates][32] (the layout adjoint_pass_kernel writes: k_env_final sums it)
template <int TB, int MINB>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
adjoint_pass_f32_kernel(float2 *psi, float2 *lam, const __grid_constant__ AdjParams<float> A, unsigned long long ntiles, double *partial)
{
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
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        tile_load<float, TB>(psi, A.P, smP, tid, tile);
        tile_load<float, TB>(lam, A.P, smL, tid, tile);
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < A.P.nstages; s++) {
            adjoint_stage_f32<TB>(A, s, smP, smL, acc_warp, tid);
            __syncthreads();
        }
        tile_store<float, TB>(psi, A.P, smP, tid, tile);
        tile_store<float, TB>(lam, A.P, smL, tid, tile);
        __syncthreads();
    }
    for (int i = tid; i < kAdjMaxGates * 32; i += NT) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) s += acc[w * (kAdjMaxGates * 32) + i];
        partial[(size_t)blockIdx.x * (kAdjMaxGates * 32) + i] = s;
    }
}
#endif

} // namespace qcb
