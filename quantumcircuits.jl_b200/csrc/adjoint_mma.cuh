// adjoint_mma.cuh -- backward pass of the adjoint gradient sweep, ComplexF64, entirely on the FP64 tensor cores.
//
// Same contract as adjoint_pass_kernel (adjoint.cuh): one launch un-applies a trapezoid of gates of the reversed
// circuit on a psi tile and a lambda tile and reduces one 4x4 cross-density matrix per gate (src/gradients.jl:107-139 is
// what it replaces).  What changes is where the arithmetic and the data movement happen.  adjoint.cuh keeps one amplitude
// quad per thread, so every gate costs three shared-memory round trips (psi, lambda, and the operand fetch of the
// reduction MMAs): ncu showed it bound by shared-memory wavefronts (LSU 72 %, FP64 36 %).  Here a warp works on 8 quads at
// a time in the MMA fragment layout (mma.sync.m8n8k4.f64; lane = 4 g + t):
//
//   un-apply  D[m][n] = sum_k A[m][k] B[k][n]   A = the real 8x8 form of U^dagger (rows m = (amplitude, Re/Im) of the
//             output, 4 columns per MMA: the Re parts, then the Im parts of the 4 input amplitudes), B[k][n] = amplitude k
//             of quad n: lane (g, t) feeds both MMAs from ONE 16-byte shared-memory load (amplitude t of quad g), and
//             receives component g of quads 2t, 2t+1.
//   reduce    E[m][n] += sum_k L'[m][k] P'[k][n]   with k = quad: lane (g, t) must supply component g of quad k for both
//             operands -- exactly what the un-apply MMAs just left in its registers (quads 2t and 2t+1 -> two MMAs).
//             No operand fetch at all; E = sum_quads lambda'_m psi'_n stays distributed over the warp (2 doubles/lane).
//   store     two 8-byte stores per state per lane (component g of quads 2t, 2t+1).
//
// Shared-memory traffic per gate is one read + one write of both tiles (the minimum), no amplitude is held in a
// register across gates, and all FP64 work runs at the DMMA rate.  The reduction uses the states AFTER the gate has been
// un-applied on both:
//   M[r', c] = sum_rest conj(lambda_{g-1}[r', rest]) psi_{g-1}[c, rest],
// and the environment the gradient needs, env[r, c] = sum conj(lambda_g[r]) psi_{g-1}[c], follows on the host as
// env = conj(U) M because lambda_g = U lambda_{g-1} (U unitary -- the adjoint method already relies on it to recover
// psi_{g-1} = U^dagger psi_g).
//
// Windows, stages and the XOR swizzle are those of tile.cuh: a warp owns WPW = 2^(TB-4)/NW of the tile's 16-amplitude
// windows for the duration of a stage (gates of a stage stay inside a window, so warps only meet at stage boundaries); a
// group of 8 quads = 2 whole windows that differ in warp-window bit `sel` (chosen on the host so that the 8-byte stores
// of a half warp hit 16 distinct bank pairs; the 16-byte loads of a quarter warp vary 3 of the 4 window bits and are
// conflict free by construction of the swizzle).
//
// Everything a thread needs for one gate -- its two MMA operands, its load / store offsets, the Gray-code steps from group
// to group, the kind of barrier that follows -- is a 32-byte AdjmLaneProg record computed on the host with the functions
// below (the same ones the CPU replay uses) and fetched from global memory one gate ahead: two coalesced 16-byte loads
// instead of a chain of dependent constant-bank reads per gate (ncu on the first version of this kernel: 19 % of the
// warps' stall samples sat on a 32-way divergent constant fetch of the operands, 17 % on the per-gate reduction).
#pragma once
#include <cmath>
#include <cstring>

#include "adjoint.cuh"

namespace qcb {

struct AdjMmaParams {
    AdjParams<double> A;
    uint8_t sel[kMaxStages * kSlots]; // warp-window bit that pairs the two windows of a group, per (stage, slot)
};

// one thread's program for one gate of the pass (index [gate ordinal][thread of the CTA])
struct AdjmLaneProg {
    double a_re, a_im; // A operands of the two un-apply MMAs: row (amplitude g>>1, part g&1) of the real form of U^dagger
    uint16_t ld;       // element offset of amplitude t of quad g of the warp's first group (load)
    uint16_t st[2];    // ... of amplitude g>>1 of quads 2t, 2t+1 (store; the lane writes component g&1)
    uint16_t cg[3];    // offsets toggled by the group-index bits (Gray-code walk over the groups)
    uint16_t flags;    // bit 0: last gate of its stage (CTA barrier after it instead of a warp barrier)
    uint16_t pad;      // 0 (the kernel uses it as a zero the compiler cannot see through)
};
static_assert(sizeof(AdjmLaneProg) == 32, "two 16-byte loads per gate");

// the two bits of q spread over the window bits other than LO, LO+1
QCB_HD int adjm_other(int LO, int q) { return LO == 0 ? (q << 2) : (LO == 1 ? ((q & 1) | ((q & 2) << 2)) : q); }

// swizzled tile-local offset of window `geom` (the thread index of tile_stage); XOR-linear in geom
template <int TB> QCB_HD uint32_t adjm_win_base(const StageDesc &S, uint32_t geom)
{
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TB - kRegBits; k++) e |= ((geom >> k) & 1u) << S.tpos[k];
    return swz<3>(e);
}

// Offsets are XOR-linear in the window id, so a stage needs only the warp's base window and one constant per warp-window
// bit; a gate picks `sel` out of them and walks its groups in Gray-code order (one XOR per group).
template <int N> struct Log2 { static constexpr int value = 1 + Log2<N / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

template <int TB, int WB> struct AdjmStage {
    uint32_t wbase;  // offset of the warp's first window
    uint32_t cw[WB]; // offset toggled by warp-window bit k
};
template <int TB, int WB> QCB_HD AdjmStage<TB, WB> adjm_stage_consts(const StageDesc &S, uint32_t warp)
{
    AdjmStage<TB, WB> C;
    C.wbase = adjm_win_base<TB>(S, warp << WB);
#pragma unroll
    for (int k = 0; k < WB; k++) C.cw[k] = adjm_win_base<TB>(S, 1u << k);
    return C;
}
// per-gate lane constants: XOR them with the group offset to get the element index
struct AdjmLane {
    uint32_t ld;    // amplitude t of quad g (load)
    uint32_t st[2]; // amplitude g>>1 of quads 2t, 2t+1 (store; the lane writes component g&1 of it)
};
template <int TB, int WB> struct AdjmGate {
    AdjmLane L;
    uint32_t cg[WB > 1 ? WB - 1 : 1]; // offsets toggled by the group-index bits (the warp-window bits other than sel)
};
template <int TB, int WB> QCB_HD AdjmGate<TB, WB> adjm_gate_consts(const AdjmStage<TB, WB> &C, const StageDesc &S, int LO, int sel, uint32_t lane)
{
    const uint32_t g = lane >> 2, t = lane & 3u;
    AdjmGate<TB, WB> G;
    uint32_t csel = 0;
#pragma unroll
    for (int k = 0; k < WB; k++) csel = k == sel ? C.cw[k] : csel;
    G.cg[0] = 0;
#pragma unroll
    for (int j = 0; j + 1 < WB; j++) G.cg[j] = j >= sel ? C.cw[j + 1] : C.cw[j];
    G.L.ld = C.wbase ^ ((g >> 2) ? csel : 0u) ^ swz<3>((uint32_t)(adjm_other(LO, (int)(g & 3u)) | (int)(t << LO)) << S.p0);
    const uint32_t wst = C.wbase ^ ((t >> 1) ? csel : 0u);
#pragma unroll
    for (int j = 0; j < 2; j++)
        G.L.st[j] = wst ^ swz<3>((uint32_t)(adjm_other(LO, (int)(2u * (t & 1u)) + j) | (int)((g >> 1) << LO)) << S.p0);
    return G;
}
// group i -> group i+1 in Gray-code order: toggle group-index bit ctz(i+1)
QCB_HD constexpr int adjm_gray_bit(int i)
{
    int b = 0;
    while (!(((i + 1) >> b) & 1)) b++;
    return b;
}
// A operands of the two un-apply MMAs for lane (g, t): row (amplitude a = g>>1, part = g&1) of the real form of U^dagger,
// columns Re(in_t) and Im(in_t)
QCB_HD void adjm_gate_operand(const GateMat<double> &Ud, uint32_t lane, double &a_re, double &a_im)
{
    const uint32_t g = lane >> 2, t = lane & 3u, a = g >> 1;
    const double ur = gate_re(Ud, (int)a, (int)t), ui = gate_im(Ud, (int)a, (int)t);
    a_re = (g & 1u) ? ui : ur;
    a_im = (g & 1u) ? ur : -ui;
}

// host side: the choice of `sel` per gate (the half-warp stores vary window bits LO and "other bit 1" plus warp-window bit
// sel: their bank vectors must be independent) ...
inline void fill_adj_mma(AdjMmaParams &M, int TB, int NW)
{
    int wbits = 0;
    while ((1 << (wbits + 1)) <= ((1 << (TB - kRegBits)) / NW)) wbits++;
    std::memset(M.sel, 0, sizeof M.sel);
    for (int s = 0; s < M.A.P.nstages; s++) {
        const StageDesc &S = M.A.P.stage[s];
        for (int slot = 0; slot < kSlots; slot++) {
            if (!((S.mask >> slot) & 1)) continue;
            const int LO = slot_lo(slot);
            const int ob1 = LO == 2 ? 1 : 3;
            const uint32_t v1 = bank_vector<3>(S.p0 + LO), v2 = bank_vector<3>(S.p0 + ob1);
            for (int k = 0; k < wbits; k++) {
                const uint32_t v = bank_vector<3>(S.tpos[k]);
                if (v != 0 && v != v1 && v != v2 && v != (v1 ^ v2)) { M.sel[s * kSlots + slot] = (uint8_t)k; break; }
            }
        }
    }
}
// ... and the threads' programs, prog[ordinal * 32 NW + thread], gates in processing order (stage by stage, slot by slot =
// the ordinals of A.slot_gate).  One record (host + device: the GPU builds the table with k_build_adj_mma_prog from the
// kernel parameters -- no host-to-device copy, which from pageable memory would synchronise the stream once per pass; the
// CPU replay and the `adjoint_prog_host` cross-check option build it on the host):
template <int TB, int NW> QCB_HD AdjmLaneProg adj_mma_prog_record(const AdjMmaParams &M, int s, int slot, uint32_t warp, uint32_t lane)
{
    constexpr int WPW = (1 << (TB - kRegBits)) / NW;
    constexpr int WB = Log2<WPW>::value;
    static_assert(WB >= 1 && WB <= 4, "2 to 16 windows per warp");
    const StageDesc &S = M.A.P.stage[s];
    const AdjmStage<TB, WB> C = adjm_stage_consts<TB, WB>(S, warp);
    const AdjmGate<TB, WB> G = adjm_gate_consts<TB, WB>(C, S, slot_lo(slot), M.sel[s * kSlots + slot], lane);
    AdjmLaneProg p;
    adjm_gate_operand(M.A.P.U[s * kSlots + slot], lane, p.a_re, p.a_im);
    p.ld = (uint16_t)G.L.ld;
    p.st[0] = (uint16_t)G.L.st[0];
    p.st[1] = (uint16_t)G.L.st[1];
    p.cg[0] = p.cg[1] = p.cg[2] = 0;
#pragma unroll
    for (int j = 0; j + 1 < WB; j++) p.cg[j] = (uint16_t)G.cg[j];
    p.flags = (S.mask >> (slot + 1)) == 0 ? 1 : 0; // last active slot of its stage
    p.pad = 0;
    return p;
}
template <int TB, int NW> inline void build_adj_mma_prog(const AdjMmaParams &M, AdjmLaneProg *prog)
{
    for (int s = 0; s < M.A.P.nstages; s++)
        for (int slot = 0; slot < kSlots; slot++) {
            if (!((M.A.P.stage[s].mask >> slot) & 1)) continue;
            const int ord = M.A.slot_gate[s * kSlots + slot];
            for (uint32_t t = 0; t < 32u * NW; t++) prog[(size_t)ord * (32 * NW) + t] = adj_mma_prog_record<TB, NW>(M, s, slot, t >> 5, t & 31u);
        }
}
#if defined(__CUDACC__)
// grid = (stage, slot) pairs = kMaxStages * kSlots blocks of 32 NW threads; inactive slots exit
template <int TB, int NW> __global__ void k_build_adj_mma_prog(const __grid_constant__ AdjMmaParams M, AdjmLaneProg *prog, const __grid_constant__ EnvIds ids,
                                                               int *id_row /* may be null */)
{
    if (id_row && blockIdx.x == 0 && threadIdx.x < kAdjMaxGates) id_row[threadIdx.x] = (int)threadIdx.x < ids.n ? ids.id[threadIdx.x] : -1;
    const int s = blockIdx.x / kSlots, slot = blockIdx.x % kSlots;
    if (s >= M.A.P.nstages || !((M.A.P.stage[s].mask >> slot) & 1)) return;
    const int ord = M.A.slot_gate[s * kSlots + slot];
    prog[(size_t)ord * (32 * NW) + threadIdx.x] = adj_mma_prog_record<TB, NW>(M, s, slot, threadIdx.x >> 5, threadIdx.x & 31u);
}
#endif

// raw[j * 32 + lane] = E[g][2t + j] (lane = 4g + t; rows m = (r', Re/Im of lambda), columns n = (c, Re/Im of psi)) ->
// M[r', c] in the layout of the environments, out32[2 (r' + 4c)] = Re, + 1 = Im:
// Re = E[2r'][2c] + E[2r'+1][2c+1], Im = E[2r'][2c+1] - E[2r'+1][2c]
QCB_HD void adjm_combine_raw(const double *raw, double *out32)
{
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            const int top = 4 * (2 * r) + c, bot = 4 * (2 * r + 1) + c; // lanes holding columns 2c, 2c+1 of rows 2r', 2r'+1
            out32[2 * (r + 4 * c)] = raw[top] + raw[32 + bot];
            out32[2 * (r + 4 * c) + 1] = raw[32 + top] - raw[bot];
        }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// State a thread carries from gate to gate (across stages and tiles): the MMA accumulators of the gate it finished last and
// the program of the gate it will work on next.  The accumulators are flushed in the middle of the NEXT gate, when the MMAs
// that produce them have long retired (an in-order warp that consumed them right away would stall at every gate end), and
// without touching the FP64 pipe, which the MMAs keep busy (ncu on an earlier version: a DADD waits tens of cycles for a
// slot): each lane adds its two raw values E[g][2t], E[g][2t+1] to a row in global memory that belongs to its warp alone
// with fire-and-forget reductions (red.global.add.f64, executed by the L2).  One writer per address and same-address
// program order make the sums reproducible; the rows are summed in fixed order and turned into the complex 4x4 matrix by
// k_env_reduce_raw / adjm_combine_raw afterwards.
struct AdjmCarry {
    double e0, e1; // pending accumulators E[g][2t], E[g][2t+1]
    double *row;   // their row: [2][32] doubles (column parity, lane)
    bool first;    // they are the row's first contribution (the CTA's first tile): stored, not added -- the rows need no zeroing
    uint4 prog[2]; // AdjmLaneProg of the next gate
};

__device__ __forceinline__ void adjm_flush(const AdjmCarry &C, uint32_t lane)
{
    if (C.first) {
        C.row[lane] = C.e0;
        C.row[32 + lane] = C.e1;
    } else {
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(C.row + lane), "d"(C.e0) : "memory");
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(C.row + 32 + lane), "d"(C.e1) : "memory");
    }
}

// one gate of the pass on the warp's windows; returns the program's flags
template <int TB, int NW>
__device__ __forceinline__ uint32_t adjoint_gate_mma(double2 *smP, double2 *smL, double *row, bool first_tile, uint32_t lane, AdjmCarry &carry,
                                                     const uint4 *__restrict__ next_prog)
{
    constexpr int WPW = (1 << (TB - kRegBits)) / NW; // windows per warp
    constexpr int NG = WPW / 2;                      // groups of 8 quads per gate
    static_assert(WPW >= 2 && WPW <= 16, "2 to 16 windows per warp");
    const uint32_t part = (lane >> 2) & 1u;
    double *dP = reinterpret_cast<double *>(smP) + part, *dL = reinterpret_cast<double *>(smL) + part;
    // unpack this gate's program, start fetching the next one
    const uint4 w0 = carry.prog[0], w1 = carry.prog[1];
    carry.prog[0] = __ldg(next_prog);
    carry.prog[1] = __ldg(next_prog + 1);
    const double a_re = __hiloint2double((int)w0.y, (int)w0.x), a_im = __hiloint2double((int)w0.w, (int)w0.z);
    const uint32_t ld = w1.x & 0xffffu, st0 = w1.x >> 16, st1 = w1.y & 0xffffu;
    const uint32_t cg[3] = {w1.y >> 16, w1.z & 0xffffu, w1.z >> 16};
    const uint32_t zero = w1.w >> 16; // AdjmLaneProg::pad: 0, but only the host knows
    double e0 = 0.0, e1 = 0.0;
    uint32_t gb = 0;
    double2 x = smP[ld], l = smL[ld];
#pragma unroll
    for (int i = 0; i < NG; i++) {
        const double2 xc = x, lc = l;
        const uint32_t gbc = gb;
        if (i + 1 < NG) { // next group's operands are in flight while this group's MMAs run
            gb ^= cg[adjm_gray_bit(i)];
            x = smP[gb ^ ld];
            l = smL[gb ^ ld];
        }
        double p0 = 0.0, p1 = 0.0, l0 = 0.0, l1 = 0.0;
        dmma(p0, p1, a_re, xc.x);
        dmma(l0, l1, a_re, lc.x);
        dmma(p0, p1, a_im, xc.y);
        dmma(l0, l1, a_im, lc.y);
        const uint32_t s0 = 2u * (gbc ^ st0), s1 = 2u * (gbc ^ st1);
        dP[s0] = p0; dP[s1] = p1;
        dL[s0] = l0; dL[s1] = l1;
        dmma(e0, e1, l0, p0);
        dmma(e0, e1, l1, p1);
        if (i == (NG > 1 ? 1 : 0)) {
            // the previous gate's accumulators.  The XOR with (bits of this group's result) & 0 ties the flush to a value of THIS
            // group at the cost of two integer instructions: without a true dependency ptxas may hoist it to the top of the
            // gate, where the in-order warp would wait for the previous gate's last MMAs after all
            const int tie = __double2hiint(l1) & (int)zero;
            carry.e0 = __hiloint2double(__double2hiint(carry.e0), __double2loint(carry.e0) ^ tie);
            carry.e1 = __hiloint2double(__double2hiint(carry.e1), __double2loint(carry.e1) ^ tie);
            adjm_flush(carry, lane);
        }
    }
    carry.e0 = e0; carry.e1 = e1;
    carry.row = row;
    carry.first = first_tile;
    return w1.w & 0xffffu;
}

// tile <-> shared memory with NT >= 2^(TB-4) threads: thread = (window id, part), each part moves 16 / PARTS elements of
// each state.  A thread stores to global memory exactly the shared-memory slots it fills on a load (the backward pass
// never permutes bits: st_lpos == ld_lpos, checked on the host), so a tile's write-back and the next tile's fill need no
// barrier between them, and the next tile's global loads are issued before the write-back starts (one state at a time:
// 32 registers in flight).
template <int TB, int NT> struct AdjmTileIo {
    static constexpr int TBITS = TB - kRegBits;
    static constexpr int PARTS = NT >> TBITS;
    static constexpr int NE = kRegAmps / PARTS;
    static_assert(PARTS >= 1 && PARTS <= kRegAmps, "thread count must be 1..16 times the number of windows");
    uint32_t e, prt;
    unsigned long long gthread; // the thread's share of the global index (without the tile's)
    __device__ __forceinline__ AdjmTileIo(const PassParams<double> &P, uint32_t tid)
    {
        const uint32_t geom = tid & ((1u << TBITS) - 1u);
        prt = tid >> TBITS;
        gthread = 0;
        e = 0;
#pragma unroll
        for (int k = 0; k < TBITS; k++) {
            const uint32_t bit = (geom >> k) & 1u;
            gthread |= (unsigned long long)bit << P.ld_phys[k];
            e |= bit << P.ld_lpos[k];
        }
        e = swz<3>(e);
    }
    __device__ __forceinline__ void fetch(const double2 *__restrict__ st, const PassParams<double> &P, unsigned long long tile, double2 (&r)[NE]) const
    {
        const unsigned long long g = cta_base(P, tile) | gthread;
#pragma unroll
        for (int i = 0; i < NE; i++) r[i] = st[g | P.ld_goff[i * PARTS + prt]];
    }
    __device__ __forceinline__ void fill(const PassParams<double> &P, double2 *sm, const double2 (&r)[NE]) const
    {
#pragma unroll
        for (int i = 0; i < NE; i++) sm[e ^ P.ld_eoff[i * PARTS + prt]] = r[i];
    }
    __device__ __forceinline__ void write_back(double2 *__restrict__ st, const PassParams<double> &P, unsigned long long tile, const double2 *sm) const
    {
        const unsigned long long g = cta_base(P, tile) | gthread;
#pragma unroll
        for (int i = 0; i < NE; i++) st[g | P.ld_goff[i * PARTS + prt]] = sm[e ^ P.ld_eoff[i * PARTS + prt]];
    }
};

constexpr int kAdjmRow = 64; // raw accumulators per warp and gate

template <int TB, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
adjoint_pass_mma_kernel(double2 *psi, double2 *lam, const __grid_constant__ PassParams<double> P, int ngates,
                        const uint4 *__restrict__ prog /* AdjmLaneProg [ngates][32 NW] */, unsigned long long ntiles,
                        double *rows /* [gridDim.x][NW][kAdjMaxGates][kAdjmRow], need not be initialised */)
{
    constexpr int NT = 32 * NW;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    double2 *smP = reinterpret_cast<double2 *>(qcb_smem_raw);
    double2 *smL = smP + (1 << TB);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    double *rows_warp = rows + ((size_t)blockIdx.x * NW + warp) * (kAdjMaxGates * kAdjmRow);
    const uint4 *my_prog = prog + 2 * tid;
    AdjmCarry carry;
    carry.e0 = carry.e1 = 0.0; // nothing pending: the first flush adds zeros to row 0 (whatever it holds: gate 0's own
    carry.row = rows_warp;     // first contribution is stored over it afterwards, same thread, same address)
    carry.first = false;
    carry.prog[0] = __ldg(my_prog);
    carry.prog[1] = __ldg(my_prog + 1);
    const AdjmTileIo<TB, NT> io(P, tid);
    double2 r[AdjmTileIo<TB, NT>::NE];
    unsigned long long tile = blockIdx.x; // (the host launches at most ntiles CTAs)
    io.fetch(psi, P, tile, r);
    io.fill(P, smP, r);
    io.fetch(lam, P, tile, r);
    io.fill(P, smL, r);
    while (true) {
        __syncthreads();
#pragma unroll 1
        for (int gi = 0; gi < ngates; gi++) {
            const int gn = gi + 1 < ngates ? gi + 1 : 0; // (the next tile runs the same program)
            const uint32_t flags = adjoint_gate_mma<TB, NW>(smP, smL, rows_warp + kAdjmRow * gi, tile == blockIdx.x, lane, carry, my_prog + (size_t)gn * (2 * NT));
            if (flags & 1u) __syncthreads(); // gates of the next stage regroup the tile's windows over the warps
            else __syncwarp();
        }
        // (the last gate of the pass ends a stage: the CTA barrier above orders the gate phase before the write-back)
        const unsigned long long next = tile + gridDim.x;
        if (next >= ntiles) {
            io.write_back(psi, P, tile, smP);
            io.write_back(lam, P, tile, smL);
            break;
        }
        io.fetch(psi, P, next, r); // in flight during the write-back
        io.write_back(psi, P, tile, smP);
        io.fill(P, smP, r);
        io.fetch(lam, P, next, r);
        io.write_back(lam, P, tile, smL);
        io.fill(P, smL, r);
        tile = next;
    }
    adjm_flush(carry, lane);
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// CPU replay of one stage, all warps, lane by lane, with a software model of the m8n8k4 fragment layout
// (A: lane 4g+t holds A[g][t]; B: lane 4g+t holds B[t][g]; C/D: lane 4g+t holds D[g][2t], D[g][2t+1]).
// conflicts (optional): [0] += shared-memory wavefronts issued, [1] += wavefronts a conflict-free layout would need.
// ---------------------------------------------------------------------------------------------------------------------
inline void adjm_soft_mma(double (&d0)[32], double (&d1)[32], const double (&a)[32], const double (&b)[32])
{
    double D[8][8];
    for (int m = 0; m < 8; m++)
        for (int n = 0; n < 8; n++) {
            double acc = (n & 1) ? d1[4 * m + (n >> 1)] : d0[4 * m + (n >> 1)];
            for (int k = 0; k < 4; k++) acc = std::fma(a[4 * m + k], b[4 * n + k], acc);
            D[m][n] = acc;
        }
    for (int ln = 0; ln < 32; ln++) {
        d0[ln] = D[ln >> 2][2 * (ln & 3)];
        d1[ln] = D[ln >> 2][2 * (ln & 3) + 1];
    }
}

// wavefronts of one warp-wide access of `bytes` (8 or 16) per lane at double-index addresses addr8[]
inline void adjm_count_wavefronts(const uint32_t (&addr8)[32], int bytes, long long *conflicts)
{
    if (!conflicts) return;
    const int lanes_per_req = 128 / bytes; // half warp for 8-byte, quarter warp for 16-byte accesses
    for (int first = 0; first < 32; first += lanes_per_req) {
        int worst = 0;
        for (int bank = 0; bank < 16; bank++) { // 16 bank pairs of 8 bytes
            uint32_t seen[32];
            int ns = 0;
            for (int ln = first; ln < first + lanes_per_req; ln++)
                for (int w = 0; w < bytes / 8; w++) {
                    const uint32_t a = addr8[ln] + (uint32_t)w;
                    if ((int)(a & 15u) != bank) continue;
                    bool dup = false;
                    for (int z = 0; z < ns; z++) dup = dup || seen[z] == a;
                    if (!dup) seen[ns++] = a;
                }
            worst = ns > worst ? ns : worst;
        }
        conflicts[0] += worst;
        conflicts[1] += 1;
    }
}

// all gates of a pass on one tile, driven by the same lane programs the GPU fetches
template <int TB, int NW>
inline void adjoint_tile_mma_host(const AdjmLaneProg *prog, int ngates, double2 *smP, double2 *smL, double *rows /*[NW][kAdjMaxGates][64] raw*/,
                                  long long *conflicts = nullptr)
{
    constexpr int WPW = (1 << (TB - kRegBits)) / NW;
    constexpr int NG = WPW / 2;
    static_assert(WPW >= 2 && WPW <= 16, "2 to 16 windows per warp");
    double *fP = reinterpret_cast<double *>(smP), *fL = reinterpret_cast<double *>(smL);
    for (int gi = 0; gi < ngates; gi++)
        for (uint32_t warp = 0; warp < (uint32_t)NW; warp++) {
            const AdjmLaneProg *L = prog + (size_t)gi * (32 * NW) + warp * 32;
            double a_re[32], a_im[32];
            for (int ln = 0; ln < 32; ln++) { a_re[ln] = L[ln].a_re; a_im[ln] = L[ln].a_im; }
            double e0[32] = {0}, e1[32] = {0};
            uint32_t gb = 0;
            for (int i = 0; i < NG; i++) {
                if (i > 0) gb ^= L[0].cg[adjm_gray_bit(i - 1)];
                double xr[32], xi[32], lr[32], li[32];
                uint32_t a8[32];
                for (int ln = 0; ln < 32; ln++) {
                    const uint32_t idx = gb ^ L[ln].ld;
                    xr[ln] = smP[idx].x; xi[ln] = smP[idx].y;
                    lr[ln] = smL[idx].x; li[ln] = smL[idx].y;
                    a8[ln] = 2 * idx;
                }
                adjm_count_wavefronts(a8, 16, conflicts);
                adjm_count_wavefronts(a8, 16, conflicts);
                double p0[32] = {0}, p1[32] = {0}, l0[32] = {0}, l1[32] = {0};
                adjm_soft_mma(p0, p1, a_re, xr);
                adjm_soft_mma(l0, l1, a_re, lr);
                adjm_soft_mma(p0, p1, a_im, xi);
                adjm_soft_mma(l0, l1, a_im, li);
                for (int j = 0; j < 2; j++) {
                    for (int ln = 0; ln < 32; ln++) {
                        const uint32_t part = ((uint32_t)ln >> 2) & 1u;
                        const uint32_t at = 2u * (gb ^ L[ln].st[j]) + part;
                        fP[at] = j ? p1[ln] : p0[ln];
                        fL[at] = j ? l1[ln] : l0[ln];
                        a8[ln] = at;
                    }
                    adjm_count_wavefronts(a8, 8, conflicts);
                    adjm_count_wavefronts(a8, 8, conflicts);
                }
                adjm_soft_mma(e0, e1, l0, p0);
                adjm_soft_mma(e0, e1, l1, p1);
            }
            double *row = rows + ((size_t)warp * kAdjMaxGates + gi) * 64;
            for (int ln = 0; ln < 32; ln++) {
                row[ln] += e0[ln];
                row[32 + ln] += e1[ln];
            }
        }
}

// env[r, c] = sum_r' conj(U[r, r']) M[r', c], in place on 32 doubles laid out [2 (r + 4 c)] (Re, Im); U32 = the forward gate
// in the same layout
inline void adjm_env_from_post(const double *U32, double *env32)
{
    double out[32];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            double re = 0, im = 0;
            for (int k = 0; k < 4; k++) {
                const double ur = U32[2 * (r + 4 * k)], ui = -U32[2 * (r + 4 * k) + 1];
                const double mr = env32[2 * (k + 4 * c)], mi = env32[2 * (k + 4 * c) + 1];
                re += ur * mr - ui * mi;
                im += ur * mi + ui * mr;
            }
            out[2 * (r + 4 * c)] = re;
            out[2 * (r + 4 * c) + 1] = im;
        }
    std::memcpy(env32, out, sizeof out);
}

} // namespace qcb
