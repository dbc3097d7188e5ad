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
// 4x4 environment with FP64 tensor-core MMAs (2 accumulators per lane) -> per-warp shared-memory accumulators that
// live across all tiles of the CTA -> one partial per CTA -> fixed-order final sum (deterministic).  HBM traffic per pass: read + write of both states, independent of
// the number of gates fused.
#pragma once
#include "tile.cuh"

namespace qcb {

constexpr int kAdjMaxGates = 10; // gates per backward pass (bounds the shared-memory accumulators)

struct EnvIds { int id[kAdjMaxGates]; int n; }; // accumulator row (ordinal inside the pass) -> gate index of the circuit

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


// One stage for one thread.  Nothing is kept in registers between gates, so the slot and the amplitude quad are
// *runtime* loop indices: one copy of each quad body instead of sixteen (the fully unrolled form overflowed the
// instruction cache; ncu: `no_instructions` was the top stall reason).
//
// Device path, per gate:  (1) every thread un-applies the gate on its 4 psi quads (shared memory -> registers ->
// shared memory);  (2) the warp contracts lambda with the updated psi over its 128 quads with FP64 tensor-core MMAs
// (mma.sync m8n8k4): A[m][k] = component m of lambda (Re/Im of the 4 amplitudes) of quad k, B[k][n] = component n of
// psi of quad k, so D[m][n] = sum_quads lambda_m psi_n lands distributed over the warp, 2 accumulators per lane
// instead of 32 per thread -- this is what lifts the register cap on occupancy, and no shuffle reduction is needed;
// (3) every thread un-applies the gate on its 4 lambda quads.  env[r,c] = sum conj(lambda_r) psi_c is assembled from D
// with one lane exchange: Re = D[2r][2c] + D[2r+1][2c+1], Im = D[2r][2c+1] - D[2r+1][2c].
// Host replay: the same three phases with a scalar environment accumulation.
// acc_warp = this warp's accumulator rows [kAdjMaxGates][32] (on the host every thread adds all 32 values of a row).
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
#if defined(__CUDA_ARCH__)
    const uint32_t lane = tid & 31u, g = lane >> 2, t = lane & 3u;
#endif
#pragma unroll 1
    for (int slot = 0; slot < kSlots; slot++) {
        if (!((mask >> slot) & 1u)) continue;
        const int LO = slot_lo(slot);
        const GateMat<R> &Ud = A.P.U[s * kSlots + slot];
        double *acc_gate = acc_warp + 32 * A.slot_gate[s * kSlots + slot];
#if defined(__CUDA_ARCH__)
        // (1) psi <- U^dagger psi
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            const int other = LO == 0 ? (q << 2) : (LO == 1 ? ((q & 1) | ((q & 2) << 2)) : q);
            V x[4];
            uint32_t off[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { off[j] = e ^ S.roff[other | (j << LO)]; x[j] = smP[off[j]]; }
            apply_quad(x[0], x[1], x[2], x[3], Ud);
#pragma unroll
            for (int j = 0; j < 4; j++) smP[off[j]] = x[j];
        }
        __syncwarp();
        // (2) D += Lambda^T Psi over the warp's 128 quads.  MMA number d contracts the 4 quads owned by lane d: lane
        // (g, t) supplies component g (amplitude g/2, Re or Im) of that lane's quad t.  The 16 amplitudes touched by one
        // MMA are the owner's 4-bit register window, which the XOR swizzle spreads evenly over the banks.
        {
            const int j = (int)(g >> 1), part = (int)(g & 1u);
            const int other = LO == 0 ? ((int)t << 2) : (LO == 1 ? (((int)t & 1) | (((int)t & 2) << 2)) : (int)t);
            const uint32_t myoff = S.roff[other | (j << LO)];
            const R *pL = reinterpret_cast<const R *>(smL) + part;
            const R *pP = reinterpret_cast<const R *>(smP) + part;
            double d0[4] = {0.0, 0.0, 0.0, 0.0}, d1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int d = 0; d < 32; d++) {
                const uint32_t idx = __shfl_sync(0xffffffffu, e, d) ^ myoff;
                const double a = (double)pL[2 * idx], b = (double)pP[2 * idx];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d0[d & 3]), "+d"(d1[d & 3]) : "d"(a), "d"(b));
            }
            const double s0 = (d0[0] + d0[1]) + (d0[2] + d0[3]), s1 = (d1[0] + d1[1]) + (d1[2] + d1[3]);
            const double o0 = __shfl_xor_sync(0xffffffffu, s0, 4), o1 = __shfl_xor_sync(0xffffffffu, s1, 4);
            if (!part) { // lane holds Re(lambda_j) rows: env[j, t]
                const int k = j + 4 * (int)t;
                acc_gate[2 * k] += s0 + o1;
                acc_gate[2 * k + 1] += s1 - o0;
            }
        }
        __syncwarp();
        // (3) lambda <- U^dagger lambda
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            const int other = LO == 0 ? (q << 2) : (LO == 1 ? ((q & 1) | ((q & 2) << 2)) : q);
            V l[4];
            uint32_t off[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { off[j] = e ^ S.roff[other | (j << LO)]; l[j] = smL[off[j]]; }
            apply_quad(l[0], l[1], l[2], l[3], Ud);
#pragma unroll
            for (int j = 0; j < 4; j++) smL[off[j]] = l[j];
        }
        __syncwarp();
#else
        for (int q = 0; q < 4; q++) {
            const int other = LO == 0 ? (q << 2) : (LO == 1 ? ((q & 1) | ((q & 2) << 2)) : q);
            V x[4], l[4];
            uint32_t off[4];
            R env[32];
            for (int i = 0; i < 32; i++) env[i] = (R)0;
            for (int j = 0; j < 4; j++) {
                off[j] = e ^ S.roff[other | (j << LO)];
                x[j] = smP[off[j]];
                l[j] = smL[off[j]];
            }
            apply_quad(x[0], x[1], x[2], x[3], Ud);
            env_accumulate<V, R>(env, l, x);
            apply_quad(l[0], l[1], l[2], l[3], Ud);
            for (int j = 0; j < 4; j++) {
                smP[off[j]] = x[j];
                smL[off[j]] = l[j];
            }
            for (int i = 0; i < 32; i++) acc_gate[i] += (double)env[i];
        }
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
