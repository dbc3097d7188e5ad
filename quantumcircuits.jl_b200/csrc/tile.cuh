// tile.cuh -- the fused gate pass: one sweep over HBM applies a whole trapezoid of
// 2-qubit gates to shared-memory tiles of the statevector.
//
// Replaces the reference's one-launch-per-gate KernelAbstractions kernel
// (_apply_2spin!, src/apply.jl:62-84, launched from src/apply.jl:134-142 after a full
// zero fill) -- same arithmetic per gate, psi'[i,j] = sum_{a,b} u[i,j,a,b] psi[a,b],
// but in place and with many gates per pass.
//
// Data layout in HBM: the reference's -- 2^n interleaved complex amplitudes, qubit q
// (1-based) = bit q-1 of the linear index.  A pass picks TB "tile bits" of the index
// (always including the lowest physical bits so every global access is a >= 128 B run);
// each CTA owns the 2^TB amplitudes that share the remaining n-TB bits, stages them in
// shared memory, and runs `nstages` register stages on them:
//
//   stage: every thread pulls the 16 amplitudes spanned by 4 consecutive local bits
//          [p0, p0+4) into registers and applies up to 4 gates on the bit pairs
//          C=(1,2), A=(0,1), B=(2,3), C=(1,2) of that window (a brickwork diamond),
//          64 real FMAs per amplitude quad per gate (48 products + 4 sums in the 3M form the
//          deep ComplexF64 passes run, apply_window_gate_3m); gate matrices are kernel
//          parameters, so ptxas feeds them to DFMA/FFMA from uniform registers.
//
// Shared memory is XOR-swizzled (low SWB bits of the 16 B / 8 B chunk index ^= xor of
// the higher SWB-bit groups) so that every stage window and the load/store phases are
// bank-conflict free; the host scheduler (schedule.hpp) orders the thread bits so that
// the lanes of a quarter warp land on distinct bank groups.
//
// All functions are __host__ __device__: tests/emu/ replays the identical code on the
// CPU, thread by thread, to check the index algebra and the scheduler without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define QCB_HD __host__ __device__ __forceinline__
#else
#define QCB_HD inline
struct double2 { double x, y; };
struct float2 { float x, y; };
#endif

namespace qcb {

constexpr int kRegBits = 4;    // qubits of the per-thread register window
constexpr int kRegAmps = 16;   // 2^kRegBits amplitudes per thread
constexpr int kSlots = 4;      // gate slots per stage (C A B C)
constexpr int kMaxStages = 12; // register stages per pass (bounds the kernel-parameter block)
constexpr int kMaxSeg = 8;
constexpr int kMinTB = 9;
constexpr int kMaxTB = 13;

template <typename R> struct Vec2;
template <> struct Vec2<double> { using type = double2; static constexpr int SWB = 3; };
template <> struct Vec2<float> { using type = float2; static constexpr int SWB = 4; };

// local bit pair (lo, lo+1) of each slot inside the 4-bit register window
QCB_HD constexpr int slot_lo(int slot) { return slot == 1 ? 0 : (slot == 2 ? 2 : 1); }

// XOR swizzle of a tile-local element index (modifies only the low SWB bits = the bank group of the 16 B / 8 B chunk).
// Every index bit b >= SWB toggles the bank-group bits given by bank_vector<SWB>(b).
//   SWB = 4 (8 B chunks, 16 groups): bit b toggles group bit b mod 4.
//   SWB = 3 (16 B chunks, 8 groups): the columns repeat with period 4 as 001, 010, 100, 111, so that ANY three of
//            four consecutive index bits are linearly independent: a quarter warp that varies three bits of a 4-bit
//            register window (stage loads, and the MMA operand loads of the backward pass) is conflict free.
//   SWB = -3 (ComplexF64 tiles written by TMA with the 128-byte hardware swizzle, tile_tma.cuh): 16 B chunk index ^= row & 7,
//            i.e. index bit b in {3,4,5} toggles group bit b - 3, higher bits toggle nothing: a register window at
//            p0 >= 3 leaves the lane bits {0,1,2} (conflict free), windows at p0 < 3 see 2-way conflicts.
template <int SWB> QCB_HD constexpr uint32_t bank_vector(int b)
{
    return SWB == -3 ? (b < 3 ? (1u << b) : (b < 6 ? (1u << (b - 3)) : 0u))
                     : SWB == 4 ? (1u << (b & 3)) : ((b & 3) == 3 ? 7u : (1u << (b & 3)));
}
template <int SWB> QCB_HD uint32_t swz(uint32_t e)
{
    if (SWB == -3) return e ^ ((e >> 3) & 7u);
    if (SWB == 3) {
        // index bits 3.. : fold in groups of 4 (bit i of h is index bit i + 3, i.e. column (i + 3) mod 4)
        const uint32_t h = e >> 3;
        const uint32_t x = (h ^ (h >> 4) ^ (h >> 8)) & 15u; // x0 <-> column 3 (111), x1 <-> 001, x2 <-> 010, x3 <-> 100
        return e ^ ((x >> 1) ^ ((x & 1u) * 7u));
    }
    return e ^ (((e >> 4) ^ (e >> 8) ^ (e >> 12)) & 15u);
}

// Gate matrix as the kernels consume it (r = output index, c = input index of the 4x4 form).
//   double: with A = Re u, B = Im u and k = 4r + c: v[k] = A_rc, v[16 + k] = B_rc (the ordinary form and the tensor-core
//           backward pass read these two, 256 contiguous bytes), v[32 + k] = -(A + B)_rc, v[48 + k] = (B - A)_rc (with A:
//           the constants of the three-multiplication form of the complex product, apply_window_gate_3m below).
//   float : packed for FFMA2 (sm_100 fma.rn.f32x2): v[4k..4k+3] = (Re u, Re u, -Im u, Im u), so that
//           (re, im) += (Re u, Re u) * (x.re, x.im) + (-Im u, Im u) * (x.im, x.re) is two packed FMAs.
template <typename R> struct alignas(16) GateMat { R v[64]; };
QCB_HD void gate_set(GateMat<double> &U, int r, int c, double re, double im)
{
    const int k = 4 * r + c;
    U.v[k] = re; U.v[16 + k] = im; U.v[32 + k] = -(re + im); U.v[48 + k] = im - re;
}
QCB_HD double gate_re(const GateMat<double> &U, int r, int c) { return U.v[4 * r + c]; }
QCB_HD double gate_im(const GateMat<double> &U, int r, int c) { return U.v[16 + 4 * r + c]; }
QCB_HD void gate_set(GateMat<float> &U, int r, int c, double re, double im)
{
    float *p = U.v + 4 * (r + 4 * c);
    p[0] = (float)re; p[1] = (float)re; p[2] = (float)-im; p[3] = (float)im;
}

struct StageDesc {
    uint8_t p0;                        // first local bit of the register window
    uint8_t mask;                      // active slots, bit s = slot s
    uint8_t tpos[kMaxTB - kRegBits];   // thread bit k -> local bit position (outside the window)
    uint8_t pad;
    uint16_t roff[kRegAmps];           // swizzled offset of register element r:  swz(r << p0)
};

template <typename R> struct PassParams {
    GateMat<R> U[kMaxStages * kSlots];
    StageDesc stage[kMaxStages];
    int32_t nstages;
    int32_t nseg;                      // the CTA index is deposited into the non-tile bits, segment by segment
    uint8_t seg_start[kMaxSeg], seg_len[kMaxSeg];
    uint8_t ld_phys[kMaxTB];           // memory-order bit k (thread bits first) -> physical index bit
    uint8_t ld_lpos[kMaxTB];           // memory-order bit k -> tile-local bit when loading
    uint8_t st_lpos[kMaxTB];           // ... when storing (differs only for bit-permuting passes)
    unsigned long long ld_goff[kRegAmps]; // global element offset of per-thread element i (top 4 memory-order bits)
    uint16_t ld_eoff[kRegAmps];        // swizzled local offset of element i when loading
    uint16_t st_eoff[kRegAmps];        // ... when storing
};

// ---------------------------------------------------------------------------------
// one gate on one amplitude quad held in registers
// ---------------------------------------------------------------------------------
QCB_HD void apply_quad(double2 &x0, double2 &x1, double2 &x2, double2 &x3, const GateMat<double> &U)
{
    const double2 in[4] = {x0, x1, x2, x3};
    double2 out[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        double re = gate_re(U, r, 0) * in[0].x;
        double im = gate_re(U, r, 0) * in[0].y;
        re = fma(-gate_im(U, r, 0), in[0].y, re);
        im = fma(gate_im(U, r, 0), in[0].x, im);
#pragma unroll
        for (int c = 1; c < 4; c++) {
            const double ur = gate_re(U, r, c), ui = gate_im(U, r, c);
            re = fma(ur, in[c].x, re);
            im = fma(ur, in[c].y, im);
            re = fma(-ui, in[c].y, re);
            im = fma(ui, in[c].x, im);
        }
        out[r].x = re;
        out[r].y = im;
    }
    x0 = out[0]; x1 = out[1]; x2 = out[2]; x3 = out[3];
}

QCB_HD void apply_quad(float2 &x0, float2 &x1, float2 &x2, float2 &x3, const GateMat<float> &U)
{
    const float2 in[4] = {x0, x1, x2, x3};
    float2 out[4];
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    // packed FP32: one FFMA2 issue slot per two FMAs, operands straight from uniform registers
    float2 sw[4];
#pragma unroll
    for (int c = 0; c < 4; c++) sw[c] = make_float2(in[c].y, in[c].x);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const float2 *u = reinterpret_cast<const float2 *>(U.v) + 2 * r;
        float2 acc = __fmul2_rn(u[0], in[0]);
        acc = __ffma2_rn(u[1], sw[0], acc);
#pragma unroll
        for (int c = 1; c < 4; c++) {
            acc = __ffma2_rn(u[8 * c], in[c], acc);
            acc = __ffma2_rn(u[8 * c + 1], sw[c], acc);
        }
        out[r] = acc;
    }
#else
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const float *u = U.v + 4 * r;
        float re = u[0] * in[0].x;
        float im = u[1] * in[0].y;
        re = fmaf(u[2], in[0].y, re);
        im = fmaf(u[3], in[0].x, im);
#pragma unroll
        for (int c = 1; c < 4; c++) {
            const float *w = U.v + 4 * (r + 4 * c);
            re = fmaf(w[0], in[c].x, re);
            im = fmaf(w[1], in[c].y, im);
            re = fmaf(w[2], in[c].y, re);
            im = fmaf(w[3], in[c].x, im);
        }
        out[r].x = re;
        out[r].y = im;
    }
#endif
    x0 = out[0]; x1 = out[1]; x2 = out[2]; x3 = out[3];
}

// gate on register-window bits (LO, LO+1): 4 quads per thread
template <int LO> QCB_HD constexpr int quad_other(int q)
{
    // spread the two bits of q over the window bits other than LO, LO+1
    int other = 0, qq = q;
    for (int b = 0; b < 4; b++)
        if (b != LO && b != LO + 1) { other |= (qq & 1) << b; qq >>= 1; }
    return other;
}

// ComplexF64, three real products per complex product (Gauss / "3M").  With u = A + iB and x = p + iq,
//   Re(u x) = A (p + q) - (A + B) q,   Im(u x) = A (p + q) + (B - A) p,
// the sum S_r = sum_c A_rc (p_c + q_c) is shared by both parts of output r and t_c = p_c + q_c by all four outputs:
// 48 DFMA/DMUL + 4 DADD per quad instead of 64 on the FP64 pipe that bounds the deep passes.  Normwise error bound as
// for the ordinary form (the component-wise one is lost: |err| <= c eps |u||x|, which is what a unitary needs).
// Order of evaluation: all four quads, two output rows at a time.  The 24 constants of a row pair are 48 uniform
// registers and are fetched once per gate for all four quads, as the 32 constants of the ordinary form are; asking
// ptxas for all 48 constants at once (one quad after the other) exceeds the 63 uniform registers, it then keeps the
// constants in vector registers (LDC), spills, and the pass runs 35 % slower than the ordinary form (measured).
template <int LO> QCB_HD void apply_window_gate_3m(double2 (&a)[kRegAmps], const GateMat<double> &U)
{
    double2 out01[4][2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int o = quad_other<LO>(q);
            double2 *x[4] = {&a[o], &a[o | (1 << LO)], &a[o | (2 << LO)], &a[o | (3 << LO)]};
            double t[4]; // (the compiler keeps t of the first row pair alive instead of adding twice)
#pragma unroll
            for (int c = 0; c < 4; c++) t[c] = x[c]->x + x[c]->y;
            double2 out[2];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const double *u = U.v + 4 * (2 * h + rr);
                double s = u[0] * t[0];
#pragma unroll
                for (int c = 1; c < 4; c++) s = fma(u[c], t[c], s);
                double re = s, im = s;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    re = fma(u[32 + c], x[c]->y, re);
                    im = fma(u[48 + c], x[c]->x, im);
                }
                out[rr].x = re;
                out[rr].y = im;
            }
            if (h == 0) { out01[q][0] = out[0]; out01[q][1] = out[1]; }
            else { *x[0] = out01[q][0]; *x[1] = out01[q][1]; *x[2] = out[0]; *x[3] = out[1]; }
        }
    }
}

// M3 selects the 3M form for ComplexF64 (ignored for ComplexF32: FFMA2 does two real products per issue slot anyway)
template <int LO, typename R, typename V, bool M3 = false> QCB_HD void apply_window_gate(V (&a)[kRegAmps], const GateMat<R> &U)
{
    if constexpr (sizeof(R) == 8 && M3) {
        apply_window_gate_3m<LO>(a, U);
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int other = quad_other<LO>(q);
            apply_quad(a[other], a[other | (1 << LO)], a[other | (2 << LO)], a[other | (3 << LO)], U);
        }
    }
}

// ---------------------------------------------------------------------------------
// phases of a pass (each is followed by a CTA barrier in the kernel)
// ---------------------------------------------------------------------------------
template <typename R> QCB_HD unsigned long long cta_base(const PassParams<R> &P, unsigned long long b)
{
    unsigned long long base = 0;
    for (int s = 0; s < P.nseg; s++) {
        const int len = P.seg_len[s];
        base |= (b & ((1ull << len) - 1ull)) << P.seg_start[s];
        b >>= len;
    }
    return base;
}

template <typename R, int TB>
QCB_HD void tile_load(const typename Vec2<R>::type *__restrict__ src, const PassParams<R> &P,
                      typename Vec2<R>::type *sm, uint32_t tid, unsigned long long bid)
{
    using V = typename Vec2<R>::type;
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    unsigned long long g = cta_base(P, bid);
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

// Source of a pass that is fused with a global<->local qubit swap (sharded states, dist.cu): after swap_top the element
// with new local index x = (r << shift) | off is the element (me << shift) | off of rank r's *old* shard, so the pass
// loads straight from the peers' memory (NVLink P2P, IPC-mapped pointers) instead of waiting for an all-to-all to land:
// base[r] = rank r's old buffer + (me << shift).
struct PeerSrc {
    const void *base[8];
    int shift; // nloc - log2(number of ranks)
};

template <typename R, int TB>
QCB_HD void tile_load_peer(const PeerSrc &S, const PassParams<R> &P, typename Vec2<R>::type *sm, uint32_t tid, unsigned long long bid)
{
    using V = typename Vec2<R>::type;
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    unsigned long long g = cta_base(P, bid);
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) {
        const uint32_t bit = (tid >> k) & 1u;
        g |= (unsigned long long)bit << P.ld_phys[k];
        e |= bit << P.ld_lpos[k];
    }
    e = swz<SWB>(e);
    const unsigned long long low = (1ull << S.shift) - 1ull;
    V tmp[kRegAmps];
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) {
        const unsigned long long x = g | P.ld_goff[i];
        tmp[i] = static_cast<const V *>(S.base[x >> S.shift])[x & low];
    }
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) sm[e ^ P.ld_eoff[i]] = tmp[i];
}

template <typename R, int TB>
QCB_HD void tile_store(typename Vec2<R>::type *__restrict__ dst, const PassParams<R> &P,
                       const typename Vec2<R>::type *sm, uint32_t tid, unsigned long long bid)
{
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    unsigned long long g = cta_base(P, bid);
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) {
        const uint32_t bit = (tid >> k) & 1u;
        g |= (unsigned long long)bit << P.ld_phys[k];
        e |= bit << P.st_lpos[k];
    }
    e = swz<SWB>(e);
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) dst[g | P.ld_goff[i]] = sm[e ^ P.st_eoff[i]];
}

// the slots of MASK, 3M arithmetic, no conditionals
template <uint32_t MASK, typename R, typename V> QCB_HD void apply_slots(V (&a)[kRegAmps], const GateMat<R> *U)
{
    if constexpr ((MASK & 1u) != 0) apply_window_gate<1, R, V, true>(a, U[0]);
    if constexpr ((MASK & 2u) != 0) apply_window_gate<0, R, V, true>(a, U[1]);
    if constexpr ((MASK & 4u) != 0) apply_window_gate<2, R, V, true>(a, U[2]);
    if constexpr ((MASK & 8u) != 0) apply_window_gate<1, R, V, true>(a, U[3]);
}
template <bool FORGET, typename V> QCB_HD void stage_store(V *sm, uint32_t e, const StageDesc &S, const V (&a)[kRegAmps])
{
#if defined(__CUDA_ARCH__)
    if (FORGET) asm volatile("" : "+r"(e));
#endif
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) sm[e ^ S.roff[r]] = a[r];
}

// FORGET: make the compiler forget the 16 element addresses between the loads and the stores (they are one XOR away from
// `e`); the persistent kernels need those registers for their ring bookkeeping.
// M3 (ComplexF64 only): 0 = ordinary arithmetic; 1 = the 3M form, four conditional gates; 2 = the 3M form for passes
// whose stages are all full diamonds, as straight-line code (four conditional gates cost ~43 register moves per gate at
// their joins -- the window must sit in the same registers on both sides).  One code path per kernel: instruction fetch
// is a limiter here.  A kernel holding the straight-line diamond AND a generic path (36 KB of stage code) ran the mixed
// passes with `no_instruction` as the top stall (ncu) and 4 % slower overall than the one-path forms; straight-line code
// for every partial mask the brickwork schedules produce (118 KB) was 5 % slower still (profiles/).
template <typename R, int TB, int SWB = Vec2<R>::SWB, bool FORGET = false, int M3 = 0>
QCB_HD void tile_stage(const PassParams<R> &P, int s, typename Vec2<R>::type *sm, uint32_t tid)
{
    using V = typename Vec2<R>::type;
    constexpr int TBITS = TB - kRegBits;
    const StageDesc &S = P.stage[s];
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) e |= ((tid >> k) & 1u) << S.tpos[k];
    e = swz<SWB>(e);
    V a[kRegAmps];
#pragma unroll
    for (int r = 0; r < kRegAmps; r++) a[r] = sm[e ^ S.roff[r]];
    const uint32_t mask = S.mask;
    if (M3 == 2) {
        // a pass of full diamonds only (the host checks every mask): straight-line code, no joins with the window live
        apply_slots<15, R, V>(a, &P.U[s * kSlots]);
        stage_store<FORGET>(sm, e, S, a);
        return;
    }
    if (M3 == 1) {
        if (mask & 1u) apply_window_gate<1, R, V, true>(a, P.U[s * kSlots + 0]);
        if (mask & 2u) apply_window_gate<0, R, V, true>(a, P.U[s * kSlots + 1]);
        if (mask & 4u) apply_window_gate<2, R, V, true>(a, P.U[s * kSlots + 2]);
        if (mask & 8u) apply_window_gate<1, R, V, true>(a, P.U[s * kSlots + 3]);
        stage_store<FORGET>(sm, e, S, a);
        return;
    }
    if (mask & 1u) apply_window_gate<1, R, V>(a, P.U[s * kSlots + 0]);
    if (mask & 2u) apply_window_gate<0, R, V>(a, P.U[s * kSlots + 1]);
    if (mask & 4u) apply_window_gate<2, R, V>(a, P.U[s * kSlots + 2]);
    if (mask & 8u) apply_window_gate<1, R, V>(a, P.U[s * kSlots + 3]);
    stage_store<FORGET>(sm, e, S, a);
}

#if defined(__CUDACC__)
template <typename R, int TB, int MINB, int M3>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
tile_pass_kernel(const typename Vec2<R>::type *src, typename Vec2<R>::type *dst,
                 const __grid_constant__ PassParams<R> P)
{
    using V = typename Vec2<R>::type;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    V *sm = reinterpret_cast<V *>(qcb_smem_raw);
    const uint32_t tid = threadIdx.x;
    const unsigned long long bid = blockIdx.x;
    tile_load<R, TB>(src, P, sm, tid, bid);
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < P.nstages; s++) {
        tile_stage<R, TB, Vec2<R>::SWB, false, M3>(P, s, sm, tid);
        __syncthreads();
    }
    tile_store<R, TB>(dst, P, sm, tid, bid);
}

// the same pass with its loads taken from the peers' shards (fused qubit swap), always out of place
template <typename R, int TB, int MINB, int M3>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
tile_pass_peer_kernel(const __grid_constant__ PeerSrc S, typename Vec2<R>::type *dst, const __grid_constant__ PassParams<R> P)
{
    using V = typename Vec2<R>::type;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    V *sm = reinterpret_cast<V *>(qcb_smem_raw);
    const uint32_t tid = threadIdx.x;
    const unsigned long long bid = blockIdx.x;
    tile_load_peer<R, TB>(S, P, sm, tid, bid);
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < P.nstages; s++) {
        tile_stage<R, TB, Vec2<R>::SWB, false, M3>(P, s, sm, tid);
        __syncthreads();
    }
    tile_store<R, TB>(dst, P, sm, tid, bid);
}
#endif

} // namespace qcb
