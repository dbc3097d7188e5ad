// adjoint.cuh -- fused backward pass of the adjoint gradient sweep.
//
// The reference computes parameter gradients by un-applying one gate and re-propagating the rest of the
// circuit 30 times (src/gradients.jl:107-139: 15 G (G+1) gate applications).  The same numbers follow from one
// backward sweep:
//     for g = G..1:   psi <- U_g^dagger psi ;  env_g[r,c] = sum_rest conj(lam[r,rest]) psi[c,rest] ;  lam <- U_g^dagger lam
// with lam = H psi_G, and grads[k,g] = 2 Re sum_{r,c} env_g[r,c] dU_g/dtheta_k[r,c].
//
// One launch of adjoint_pass_kernel does this for a whole trapezoid of gates (the reversed circuit is scheduled
// by the same pass planner as the forward circuit): a persistent CTA stages a psi tile and a lambda tile in
// shared memory, walks the register stages, and for every gate un-applies it on both tiles and accumulates the
// 4x4 environment: thread-local partial sums -> warp transpose-reduction with shuffles (31 exchanges for 32
// values) -> per-warp shared-memory accumulators that live across all tiles of the CTA -> one partial per CTA
// -> fixed-order final sum (deterministic).  HBM traffic per pass: read + write of both states, independent of
// the number of gates fused.
#pragma once
#include "tile.cuh"

namespace qcb {

constexpr int kAdjMaxGates = 10; // gates per backward pass (bounds the shared-memory accumulators)

template <typename R> struct AdjParams {
    PassParams<R> P;                          // geometry, stages, U^dagger matrices
    uint8_t slot_gate[kMaxStages * kSlots];   // ordinal of the slot's gate inside this pass (accumulator row)
    int32_t ngates;
};

// env[2*(r+4c)], env[2*(r+4c)+1] += Re, Im of conj(l_r) x_c
template <typename V, typename A> QCB_HD void env_accumulate(A (&env)[32], const V (&l)[4], const V (&x)[4])
{
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) {
            env[2 * (r + 4 * c)] = fma((A)l[r].x, (A)x[c].x, fma((A)l[r].y, (A)x[c].y, env[2 * (r + 4 * c)]));
            env[2 * (r + 4 * c) + 1] = fma((A)l[r].x, (A)x[c].y, fma(-(A)l[r].y, (A)x[c].x, env[2 * (r + 4 * c) + 1]));
        }
}

#if defined(__CUDA_ARCH__)
// after the call lane L holds in v[0] the warp-wide sum of element L
__device__ __forceinline__ void warp_transpose_reduce(double (&v)[32])
{
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const double send = up ? v[i] : v[i + s];
            const double keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
}
#endif

// One stage for one thread.  Unlike the forward pass nothing is kept in registers between gates, so the slot and the
// amplitude quad are *runtime* loop indices: one copy of the 192-FMA quad body (un-apply psi, environment, un-apply
// lambda) instead of sixteen -- the fully unrolled form overflowed the instruction cache (ncu: `no_instructions` was the
// top stall reason).  acc_warp = this warp's accumulator rows [kAdjMaxGates][32]; on the host replay every thread adds
// all 32 values of a row.
template <typename R, int TB>
QCB_HD void adjoint_stage(const AdjParams<R> &A, int s, typename Vec2<R>::type *smP, typename Vec2<R>::type *smL,
                          double *acc_warp, uint32_t tid)
{
    using V = typename Vec2<R>::type;
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    const StageDesc &S = A.P.stage[s];
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) e |= ((tid >> k) & 1u) << S.tpos[k];
    e = swz<SWB>(e);
    const uint32_t mask = S.mask;
#pragma unroll 1
    for (int slot = 0; slot < kSlots; slot++) {
        if (!((mask >> slot) & 1u)) continue;
        const int LO = slot_lo(slot);
        const GateMat<R> &Ud = A.P.U[s * kSlots + slot];
        double *acc_gate = acc_warp + 32 * A.slot_gate[s * kSlots + slot];
        R env[32];
#pragma unroll
        for (int i = 0; i < 32; i++) env[i] = (R)0;
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            // spread the two bits of q over the window bits other than LO, LO+1
            const int other = LO == 0 ? (q << 2) : (LO == 1 ? ((q & 1) | ((q & 2) << 2)) : q);
            V x[4], l[4];
            uint32_t off[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                off[j] = e ^ S.roff[other | (j << LO)];
                x[j] = smP[off[j]];
                l[j] = smL[off[j]];
            }
            apply_quad(x[0], x[1], x[2], x[3], Ud);
            env_accumulate<V, R>(env, l, x);
            apply_quad(l[0], l[1], l[2], l[3], Ud);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                smP[off[j]] = x[j];
                smL[off[j]] = l[j];
            }
        }
#if defined(__CUDA_ARCH__)
        double v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = (double)env[i];
        warp_transpose_reduce(v);
        acc_gate[threadIdx.x & 31u] += v[0]; // the row is private to this warp: plain read-modify-write
#else
        for (int i = 0; i < 32; i++) acc_gate[i] += (double)env[i];
#endif
    }
}

#if defined(__CUDACC__)
template <typename R, int TB, int MINB>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
adjoint_pass_kernel(typename Vec2<R>::type *psi, typename Vec2<R>::type *lam, const __grid_constant__ AdjParams<R> A,
                    unsigned long long ntiles, double *partial /* [gridDim.x][kAdjMaxGates][32] */)
{
    using V = typename Vec2<R>::type;
    constexpr int NT = 1 << (TB - kRegBits);
    constexpr int NW = (NT + 31) / 32;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    V *smP = reinterpret_cast<V *>(qcb_smem_raw);
    V *smL = smP + (1 << TB);
    double *acc = reinterpret_cast<double *>(smL + (1 << TB)); // [NW][kAdjMaxGates][32]
    const uint32_t tid = threadIdx.x;
    for (int i = tid; i < NW * kAdjMaxGates * 32; i += NT) acc[i] = 0.0;
    __syncthreads();
    double *acc_warp = acc + (tid >> 5) * (kAdjMaxGates * 32);
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        tile_load<R, TB>(psi, A.P, smP, tid, tile);
        tile_load<R, TB>(lam, A.P, smL, tid, tile);
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < A.P.nstages; s++) {
            adjoint_stage<R, TB>(A, s, smP, smL, acc_warp, tid);
            __syncthreads();
        }
        tile_store<R, TB>(psi, A.P, smP, tid, tile);
        tile_store<R, TB>(lam, A.P, smL, tid, tile);
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
