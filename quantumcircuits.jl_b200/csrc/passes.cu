// passes.cu -- executes a program-ordered gate list on a device state: plans fused passes
// (schedule.hpp), fills the kernel-parameter blocks and launches tile_pass_kernel (tile.cuh).
// The reference's counterpart is the host loop `for gate: apply!(cache, psi', psi, gate)`
// of src/circuits.jl:19-31 / src/gradients.jl:54-71 with one kernel launch, one zero fill and
// one synchronize per gate (src/apply.jl:134-142).
#include <sstream>

#include <type_traits>

#include "adjoint_f32.cuh"
#include "adjoint_mma.cuh"
#include "kernels.cuh"
#include "qcb_internal.hpp"
#include "tile_tma.cuh"
#include "tile_async.cuh"
#include "small_circuit.cuh"

namespace qcb {


// src == nullptr: in place.  Otherwise the pass reads `src` (another state of the same size) and writes the state: every
// amplitude is read and written exactly once per pass, so the first pass of a circuit can double as the copy of its input.
// Arithmetic of a ComplexF64 pass (tile.cuh): passes with fewer than `f64_3m_min_gates` gates are bound by HBM, not by the
// FP64 pipe, and keep the ordinary form (0: fewer constants to fetch per gate); deeper ones run the 3M form -- as
// straight-line code when every stage is a full diamond (2), with conditional gates otherwise (1).
template <typename R> static int pass_form(const PassParams<R> &P)
{
    if (sizeof(R) != 8) return 0;
    int gates = 0;
    bool full = true;
    for (int i = 0; i < P.nstages; i++) {
        gates += __builtin_popcount(P.stage[i].mask);
        full = full && P.stage[i].mask == 15;
    }
    if (gates < ctx().f64_3m_min_gates) return 0;
    ctx().st_gates_3m += gates;
    return full ? 2 : 1;
}

template <typename R, int TB, int M3 = 0> static int launch_pass(qcb_state *s, const PassParams<R> &P, const void *src = nullptr)
{
    if constexpr (sizeof(R) == 8 && M3 == 0) {
        const int form = pass_form(P);
        if (form == 1) return launch_pass<R, TB, 1>(s, P, src);
        if (form == 2) return launch_pass<R, TB, 2>(s, P, src);
    }
    using V = typename Vec2<R>::type;
    constexpr int NT = 1 << (TB - kRegBits);
    // resident threads per SM the register budget is sized for: 512 (ComplexF64, 128 registers) / 1024 (ComplexF32, 64)
    constexpr int RES = sizeof(R) == 8 ? 512 : 1024;
    constexpr int MINB = (RES / NT) > 0 ? ((RES / NT) > 32 ? 32 : (RES / NT)) : 1;
    const unsigned long long nblocks = 1ull << (s->nloc - TB);
    // persistent form with asynchronous tile loads (tile_async.cuh): two tile buffers per CTA
    if constexpr ((sizeof(V) << (TB + 1)) <= (size_t)(64 << 10)) {
        if (ctx().pass_async) {
            constexpr int ARES = sizeof(R) == 8 ? 384 : 1024; // ComplexF64: 3 x 128 threads per SM, up to 168 registers
            constexpr int AMINB = (ARES / NT) > 0 ? ((ARES / NT) > 32 ? 32 : (ARES / NT)) : 1;
            auto akern = tile_pass_async_kernel<R, TB, AMINB, M3>;
            const size_t asmem = sizeof(V) << (TB + 1);
            static int ctas_per_sm = 0;
            if (!ctas_per_sm) {
                QCB_CUDA(cudaFuncSetAttribute(akern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem));
                QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, akern, NT, asmem));
                if (ctas_per_sm < 1) ctas_per_sm = 1;
            }
            const unsigned grid = (unsigned)std::min<unsigned long long>(nblocks, (unsigned long long)ctx().sm_count * ctas_per_sm);
            PassTimerScope timer;
            akern<<<grid, NT, asmem, ctx().stream>>>((const V *)(src ? src : s->d), (V *)s->d, P, nblocks);
            QCB_CUDA(cudaGetLastError());
            count_launch();
            ctx().st_async_passes += 1;
            return QCB_OK;
        }
    }
    auto kern = tile_pass_kernel<R, TB, MINB, M3>;
    const size_t smem = sizeof(V) << TB;
    static bool attr_set = false;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    PassTimerScope timer;
    kern<<<(unsigned)nblocks, NT, smem, ctx().stream>>>((const V *)(src ? src : s->d), (V *)s->d, P);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    return QCB_OK;
}

// the same pass on the TMA-fed persistent kernel (tile_tma.cuh); P was filled with the hardware swizzle (fill_params<double, -3>)
static int launch_pass_tma(qcb_state *s, const PassParams<double> &P, const TmaPassGeom &g, const void *src)
{
    auto kern = tile_pass_tma_kernel<3>;
    const size_t smem = tma_pass_smem_bytes();
    static bool attr_set = false;
    static int ctas_per_sm = 1;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kTmaStageThreads, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        attr_set = true;
    }
    CUtensorMap map_ld, map_st;
    const char *why = "";
    if (tma_pass_map(&map_st, s->d, g, &why)) return set_error(QCB_ECUDA, "tensor map: %s", why);
    if (src && src != s->d) {
        if (tma_pass_map(&map_ld, src, g, &why)) return set_error(QCB_ECUDA, "tensor map: %s", why);
    } else
        map_ld = map_st;
    const unsigned long long ntiles = 1ull << (s->nloc - kTmaTB);
    const unsigned grid = (unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)ctx().sm_count * ctas_per_sm);
    PassTimerScope timer;
    kern<<<grid, kTmaStageThreads, smem, ctx().stream>>>(map_ld, map_st, P, g.a2 - g.b1, ntiles);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    ctx().st_tma_passes += 1;
    return QCB_OK;
}

// does this pass go to the TMA kernel?  ("pass_tma": 0 never, 1 every eligible pass, 2 eligible passes with at most
// `pass_tma_max_stages` register stages, i.e. the HBM-bound ones)
static bool use_tma_pass(const qcb_state *s, const PassPlan &pl, TmaPassGeom &g)
{
    const Ctx &c = ctx();
    if (!c.pass_tma || s->dtype != QCB_C128) return false;
    if (c.pass_tma == 2 && (long)pl.stages.size() > c.pass_tma_max_stages) return false;
    return tma_pass_eligible(pl, s->nloc, g);
}

template <typename R, int TB, int M3 = 0> static int launch_pass_peer(qcb_state *s, const PeerSrc &src, void *dst, const PassParams<R> &P)
{
    if constexpr (sizeof(R) == 8 && M3 == 0) {
        const int form = pass_form(P);
        if (form == 1) return launch_pass_peer<R, TB, 1>(s, src, dst, P);
        if (form == 2) return launch_pass_peer<R, TB, 2>(s, src, dst, P);
    }
    using V = typename Vec2<R>::type;
    constexpr int NT = 1 << (TB - kRegBits);
    constexpr int RES = sizeof(R) == 8 ? 512 : 1024;
    constexpr int MINB = (RES / NT) > 0 ? ((RES / NT) > 32 ? 32 : (RES / NT)) : 1;
    auto kern = tile_pass_peer_kernel<R, TB, MINB, M3>;
    const size_t smem = sizeof(V) << TB;
    static bool attr_set = false;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    PassTimerScope timer;
    kern<<<(unsigned)(1ull << (s->nloc - TB)), NT, smem, ctx().stream>>>(src, (V *)dst, P);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    return QCB_OK;
}

template <typename R> static int launch_pass_peer_tb(qcb_state *s, int TB, const PeerSrc &src, void *dst, const PassParams<R> &P)
{
    switch (TB) {
    case 9: return launch_pass_peer<R, 9>(s, src, dst, P);
    case 10: return launch_pass_peer<R, 10>(s, src, dst, P);
    case 11: return launch_pass_peer<R, 11>(s, src, dst, P);
    case 12: return launch_pass_peer<R, 12>(s, src, dst, P);
    case 13: return launch_pass_peer<R, 13>(s, src, dst, P);
    default: return set_error(QCB_EINVAL, "unsupported tile size %d", TB);
    }
}

template <typename R> static int launch_pass_tb(qcb_state *s, int TB, const PassParams<R> &P, const void *src = nullptr)
{
    switch (TB) {
    case 9: return launch_pass<R, 9>(s, P, src);
    case 10: return launch_pass<R, 10>(s, P, src);
    case 11: return launch_pass<R, 11>(s, P, src);
    case 12: return launch_pass<R, 12>(s, P, src);
    case 13: return launch_pass<R, 13>(s, P, src);
    default: return set_error(QCB_EINVAL, "unsupported tile size %d", TB);
    }
}

template <typename R> static int run_simple(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint)
{
    using V = typename Vec2<R>::type;
    const unsigned long long nquads = 1ull << (s->nloc - 2);
    const unsigned blocks = (unsigned)std::min<unsigned long long>((nquads + 255) / 256, (unsigned long long)ctx().sm_count * 16);
    for (const GateOp &g : gates) {
        const int pb0 = s->phys[g.q], pb1 = s->phys[g.q + 1];
        if (pb0 >= s->nloc || pb1 >= s->nloc || pb1 != pb0 + 1)
            return set_error(QCB_EUNSUPPORTED, "simple path needs an adjacent local qubit pair");
        GateMat<R> U;
        const double *m = mats + 32 * (size_t)g.mat;
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                const int src = adjoint ? (c + 4 * r) : (r + 4 * c);
                gate_set(U, r, c, m[2 * src], adjoint ? -m[2 * src + 1] : m[2 * src + 1]);
            }
        k_apply_2q_simple<R><<<blocks, 256, 0, ctx().stream>>>((V *)s->d, nquads, pb0, U);
        QCB_CUDA(cudaGetLastError());
        count_launch();
    }
    ctx().st_passes += (double)gates.size();
    return QCB_OK;
}

SchedOptions current_sched_options(const qcb_state *s)
{
    Ctx &c = ctx();
    SchedOptions opt;
    opt.tile_bits = (int)std::max<long>(kMinTB, std::min<long>(kMaxTB, c.tile_bits));
    opt.low_bits = (int)std::max<long>(0, std::min<long>(c.low_bits, opt.tile_bits - kRegBits));
    opt.max_gates_per_pass = (int)std::max<long>(1, std::min<long>(c.max_gates_per_pass, 1 << 30));
    (void)s;
    return opt;
}

int launch_plan(qcb_state *s, const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool adjoint)
{
    try {
        if (s->dtype == QCB_C128) {
            static thread_local PassParams<double> P;
            TmaPassGeom g;
            if (use_tma_pass(s, pl, g)) {
                fill_params<double, -3>(pl, s->nloc, gates, mats, adjoint, P);
                return launch_pass_tma(s, P, g, nullptr);
            }
            fill_params<double>(pl, s->nloc, gates, mats, adjoint, P);
            return launch_pass_tb<double>(s, pl.TB, P);
        }
        static thread_local PassParams<float> P;
        fill_params<float>(pl, s->nloc, gates, mats, adjoint, P);
        return launch_pass_tb<float>(s, pl.TB, P);
    } catch (const std::exception &e) {
        return set_error(QCB_EINVAL, "pass parameters: %s", e.what());
    }
}

int launch_plan_peer(qcb_state *s, const PeerSrc &src, void *dst, const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool adjoint)
{
    try {
        if (s->dtype == QCB_C128) {
            static thread_local PassParams<double> P;
            fill_params<double>(pl, s->nloc, gates, mats, adjoint, P);
            return launch_pass_peer_tb<double>(s, pl.TB, src, dst, P);
        }
        static thread_local PassParams<float> P;
        fill_params<float>(pl, s->nloc, gates, mats, adjoint, P);
        return launch_pass_peer_tb<float>(s, pl.TB, src, dst, P);
    } catch (const std::exception &e) {
        return set_error(QCB_EINVAL, "pass parameters: %s", e.what());
    }
}

static std::string plan_key(const qcb_state *s, const std::vector<GateOp> &gates, const SchedOptions &o)
{
    // the whole key material, byte for byte (a hash could collide and silently reuse the plan of another gate list)
    std::string k;
    k.reserve(16 * 4 + s->phys.size() + gates.size() * 4);
    auto put = [&](long long v, int bytes) { for (int i = 0; i < bytes; i++) k.push_back((char)((v >> (8 * i)) & 0xff)); };
    put(s->nloc, 1); put(o.tile_bits, 1); put(o.low_bits, 1); put(o.max_stages, 1); put(o.max_gates_per_pass, 4);
    put((long long)s->phys.size(), 1);
    for (int p : s->phys) put(p, 1);
    put((long long)gates.size(), 4);
    for (const GateOp &g : gates) { put(g.q, 1); put(g.mat, 3); }
    return k;
}

template <typename R> static int run_fused(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint, const void *first_src)
{
    Ctx &c = ctx();
    const SchedOptions opt = current_sched_options(s);
    const std::string key = plan_key(s, gates, opt);
    auto it = c.plan_cache.find(key);
    if (it == c.plan_cache.end()) {
        std::vector<PassPlan> plans;
        std::vector<char> done;
        try {
            plans = plan_passes(s->nloc, s->phys, gates, opt, &done);
        } catch (const std::exception &e) {
            return set_error(QCB_EINVAL, "scheduler: %s", e.what());
        }
        for (char d : done)
            if (!d) return set_error(QCB_EINVAL, "gate list could not be scheduled");
        if (c.plan_cache.size() > 64) c.plan_cache.clear();
        it = c.plan_cache.emplace(key, std::move(plans)).first;
    }
    static thread_local PassParams<R> P; // ~13 KB, reused
    for (const PassPlan &pl : it->second) {
        TmaPassGeom g;
        bool tma_pass = false;
        try {
            if constexpr (std::is_same<R, double>::value) {
                tma_pass = use_tma_pass(s, pl, g);
                if (tma_pass) fill_params<double, -3>(pl, s->nloc, gates, mats, adjoint, P);
            }
            if (!tma_pass) fill_params<R>(pl, s->nloc, gates, mats, adjoint, P);
        } catch (const std::exception &e) {
            return set_error(QCB_EINVAL, "pass parameters: %s", e.what());
        }
        if (c.time_passes) {
            int fw = s->nloc;
            for (int b : pl.phys) if (b >= opt.low_bits) { fw = b; break; }
            unsigned long long masks = 0;
            for (size_t i = 0; i < pl.stages.size() && i < 16; i++)
                for (int k = 0; k < kSlots; k++)
                    if (pl.stages[i].gate[k] >= 0) masks |= 1ull << (4 * i + k);
            c.pass_meta.push_back(Ctx::PassMeta{(int)pl.stages.size(), pl.ngates, fw, tma_pass ? 1 : 0, masks});
        }
        if constexpr (std::is_same<R, double>::value) {
            if (tma_pass) QCB_TRY(launch_pass_tma(s, P, g, first_src));
        }
        if (!tma_pass) QCB_TRY(launch_pass_tb<R>(s, pl.TB, P, first_src));
        first_src = nullptr; // only the first pass reads the input state
        c.st_passes += 1;
        c.st_stages += (double)pl.stages.size();
    }
    return QCB_OK;
}

// ---- fused backward pass of the adjoint gradient sweep (adjoint.cuh) -------------------------------------
static __global__ void k_env_final(const double *__restrict__ partial, int nblocks, EnvIds ids, double *env_out)
{
    // block = (gate ordinal, element); fixed-order sum over the persistent CTAs' partials
    const int ord = blockIdx.x >> 5, el = blockIdx.x & 31;
    double s = 0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(size_t)b * (kAdjMaxGates * 32) + ord * 32 + el];
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        env_out[(size_t)ids.id[ord] * 32 + el] = t;
    }
}

// ---- deferred reduction (Ctx::AdjDefer): d_partial = [pass areas | gate-id table (ints, kAdjMaxGates per pass) | raw sums]
// every pass records which gates its accumulator rows belong to
static __global__ void k_env_note_ids(EnvIds ids, int *table_row)
{
    if (threadIdx.x < kAdjMaxGates) table_row[threadIdx.x] = (int)threadIdx.x < ids.n ? ids.id[threadIdx.x] : -1;
}
// kind 2: block = (pass, gate ordinal, element); fixed-order sum over the pass's CTAs
static __global__ void k_env_final_deferred(const double *__restrict__ areas, size_t stride, int nblocks, const int *__restrict__ table, double *env_out)
{
    const int el = blockIdx.x & 31, ord = (blockIdx.x >> 5) % kAdjMaxGates, pass = (blockIdx.x >> 5) / kAdjMaxGates;
    const int gate = table[pass * kAdjMaxGates + ord];
    if (gate < 0) return;
    const double *partial = areas + (size_t)pass * stride;
    double s = 0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(size_t)b * (kAdjMaxGates * 32) + ord * 32 + el];
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        env_out[(size_t)gate * 32 + el] = t;
    }
}

static size_t adj_defer_table_doubles() { return (((size_t)ctx().adj_defer.npasses * kAdjMaxGates * sizeof(int) + 15) / 16) * 2; } // keeps what follows 16-byte aligned
static int *adj_defer_table() { return reinterpret_cast<int *>(ctx().d_partial + (size_t)ctx().adj_defer.npasses * ctx().adj_defer.stride); }
static double *adj_defer_raw()
{
    const Ctx::AdjDefer &D = ctx().adj_defer;
    return ctx().d_partial + (size_t)D.npasses * D.stride + adj_defer_table_doubles();
}
// area of the current pass; the first pass of a sweep sizes the arena (nothing to preserve yet)
static int adj_defer_area(size_t doubles_per_pass, int nrows, int kind, size_t extra_doubles, double **area)
{
    Ctx::AdjDefer &D = ctx().adj_defer;
    if (D.pass == 0) {
        D.stride = doubles_per_pass; D.nrows = nrows; D.kind = kind;
        const size_t total = (size_t)D.npasses * D.stride + adj_defer_table_doubles() + (size_t)D.npasses * kAdjMaxGates * 64 + extra_doubles;
        if (total * sizeof(double) > ((size_t)512 << 20)) { D.on = false; return QCB_OK; } // too many passes: reduce pass by pass
        QCB_TRY(ensure_partial(total));
    } else if (D.stride != doubles_per_pass || D.nrows != nrows || D.kind != kind) {
        return set_error(QCB_EINVAL, "backward passes of one sweep differ in shape");
    }
    *area = ctx().d_partial + (size_t)D.pass * D.stride;
    return QCB_OK;
}
static int adj_defer_note(const EnvIds &ids)
{
    Ctx::AdjDefer &D = ctx().adj_defer;
    k_env_note_ids<<<1, 32, 0, ctx().stream>>>(ids, adj_defer_table() + D.pass * kAdjMaxGates);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    D.pass += 1;
    return QCB_OK;
}

template <typename R, int TB> static int launch_adjoint(qcb_state *psi, qcb_state *lam, const AdjParams<R> &A, const EnvIds &ids, double *d_env)
{
    using V = typename Vec2<R>::type;
    constexpr int NT = 1 << (TB - kRegBits);
    constexpr int NW = (NT + 31) / 32;
    // registers (<= 168 per thread) and shared memory (two tiles + accumulators) both allow this many CTAs per SM
    constexpr int MINB = TB >= 12 ? 1 : (TB == 11 ? 3 : (TB == 10 ? 6 : 8));
    auto kern = adjoint_pass_kernel<R, TB, MINB>;
    const size_t smem = 2 * (sizeof(V) << TB) + sizeof(double) * NW * kAdjMaxGates * 32;
    static bool attr_set = false;
    static int ctas_per_sm = 1;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        attr_set = true;
    }
    const unsigned long long ntiles = 1ull << (psi->nloc - TB);
    const unsigned grid = (unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)ctx().sm_count * ctas_per_sm);
    double *area = nullptr;
    if (ctx().adj_defer.on) QCB_TRY(adj_defer_area((size_t)grid * kAdjMaxGates * 32, (int)grid, 2, 0, &area));
    if (!ctx().adj_defer.on) { QCB_TRY(ensure_partial((size_t)grid * kAdjMaxGates * 32)); area = ctx().d_partial; }
    kern<<<grid, NT, smem, ctx().stream>>>((V *)psi->d, (V *)lam->d, A, ntiles, area);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    if (ctx().adj_defer.on) return adj_defer_note(ids);
    k_env_final<<<ids.n * 32, 128, 0, ctx().stream>>>(area, (int)grid, ids, d_env);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    return QCB_OK;
}

// ComplexF32: windows of both states in registers, FFMA2 outer products, warp transpose-reduction (adjoint_f32.cuh); same partial layout
template <int TB> static int launch_adjoint_f32(qcb_state *psi, qcb_state *lam, const AdjParams<float> &A, const EnvIds &ids, double *d_env)
{
    constexpr int NT = 1 << (TB - kRegBits);
    constexpr int NW = NT / 32;
    // two windows (64 registers) + 16 packed accumulators per thread: <= 128 registers; shared memory: two tiles + accumulators
    constexpr int MINB = TB >= 12 ? 2 : (TB == 11 ? 4 : 8);
    auto kern = adjoint_pass_f32_kernel<TB, MINB>;
    const size_t smem = 2 * (sizeof(float2) << TB) + sizeof(double) * NW * kAdjMaxGates * 32;
    static bool attr_set = false;
    static int ctas_per_sm = 1;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        attr_set = true;
    }
    const unsigned long long ntiles = 1ull << (psi->nloc - TB);
    const unsigned grid = (unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)ctx().sm_count * ctas_per_sm);
    double *area = nullptr;
    if (ctx().adj_defer.on) QCB_TRY(adj_defer_area((size_t)grid * kAdjMaxGates * 32, (int)grid, 2, 0, &area));
    if (ctx().adj_defer.on) {
        // the kernel itself notes which gates its rows belong to; the rows of all passes are summed after the sweep
        kern<<<grid, NT, smem, ctx().stream>>>((float2 *)psi->d, (float2 *)lam->d, A, ntiles, area, ids, adj_defer_table() + ctx().adj_defer.pass * kAdjMaxGates);
        QCB_CUDA(cudaGetLastError());
        count_launch();
        ctx().adj_defer.pass += 1;
        return QCB_OK;
    }
    QCB_TRY(ensure_partial((size_t)grid * kAdjMaxGates * 32));
    kern<<<grid, NT, smem, ctx().stream>>>((float2 *)psi->d, (float2 *)lam->d, A, ntiles, ctx().d_partial, ids, nullptr);
    QCB_CUDA(cudaGetLastError());
    k_env_final<<<ids.n * 32, 128, 0, ctx().stream>>>(ctx().d_partial, (int)grid, ids, d_env);
    QCB_CUDA(cudaGetLastError());
    count_launch(2);
    return QCB_OK;
}

// ComplexF64: the tensor-core form (adjoint_mma.cuh); env rows hold the post-gate matrix M, see run_adjoint_fused.
// Final reduction of its raw accumulator rows, fixed summation order.  Few rows (small states: a handful of CTAs): one
// block per gate sums and combines in one launch.  Many rows: one block per (gate, raw element) so that the whole GPU
// reads the rows, then a second tiny launch turns the 64 raw sums of each gate into its complex 4x4 matrix.
static __global__ void __launch_bounds__(256) k_env_reduce_raw(const double *__restrict__ rows, int nrows /* CTAs x warps */, EnvIds ids, double *env_out)
{
    __shared__ double part[4][kAdjmRow];
    __shared__ double raw[kAdjmRow];
    const int ord = blockIdx.x, el = threadIdx.x & (kAdjmRow - 1), q = threadIdx.x >> 6;
    double s = 0;
    for (int r = q; r < nrows; r += 4) s += rows[((size_t)r * kAdjMaxGates + ord) * kAdjmRow + el];
    part[q][el] = s;
    __syncthreads();
    if (threadIdx.x < kAdjmRow) raw[el] = (part[0][el] + part[1][el]) + (part[2][el] + part[3][el]);
    __syncthreads();
    if (threadIdx.x == 0) adjm_combine_raw(raw, env_out + (size_t)ids.id[ord] * 32);
}
static __global__ void k_env_reduce_raw_wide(const double *__restrict__ rows, int nrows, double *raw_out /* [ngates][64] */)
{
    const int ord = blockIdx.x >> 6, el = blockIdx.x & 63;
    double s = 0;
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) s += rows[((size_t)r * kAdjMaxGates + ord) * kAdjmRow + el];
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        raw_out[ord * kAdjmRow + el] = t;
    }
}
static __global__ void k_env_combine_raw(const double *__restrict__ raw, EnvIds ids, double *env_out)
{
    const int ord = threadIdx.x;
    if (ord < ids.n) adjm_combine_raw(raw + ord * kAdjmRow, env_out + (size_t)ids.id[ord] * 32);
}

// deferred forms: block = (pass, gate ordinal, raw element); then one thread per (pass, gate ordinal)
static __global__ void k_env_reduce_raw_deferred(const double *__restrict__ areas, size_t stride, int nrows, const int *__restrict__ table, double *raw_out)
{
    const int el = blockIdx.x & 63, slot = blockIdx.x >> 6, ord = slot % kAdjMaxGates, pass = slot / kAdjMaxGates;
    if (table[slot] < 0) return;
    const double *rows = areas + (size_t)pass * stride;
    double s = 0;
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) s += rows[((size_t)r * kAdjMaxGates + ord) * kAdjmRow + el];
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        raw_out[(size_t)slot * kAdjmRow + el] = t;
    }
}
static __global__ void k_env_combine_raw_deferred(const double *__restrict__ raw, const int *__restrict__ table, int nslots, double *env_out)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < nslots && table[slot] >= 0) adjm_combine_raw(raw + (size_t)slot * kAdjmRow, env_out + (size_t)table[slot] * 32);
}

// end of a sweep whose reductions were deferred
static int adj_defer_finish(double *d_env)
{
    Ctx::AdjDefer &D = ctx().adj_defer;
    if (!D.on || D.pass == 0) { D.on = false; return QCB_OK; }
    const int nslots = D.pass * kAdjMaxGates;
    if (D.kind == 1) {
        k_env_reduce_raw_deferred<<<nslots * kAdjmRow, 128, 0, ctx().stream>>>(ctx().d_partial, D.stride, D.nrows, adj_defer_table(), adj_defer_raw());
        QCB_CUDA(cudaGetLastError());
        k_env_combine_raw_deferred<<<(nslots + 63) / 64, 64, 0, ctx().stream>>>(adj_defer_raw(), adj_defer_table(), nslots, d_env);
        QCB_CUDA(cudaGetLastError());
        count_launch(2);
    } else {
        k_env_final_deferred<<<nslots * 32, 128, 0, ctx().stream>>>(ctx().d_partial, D.stride, D.nrows, adj_defer_table(), d_env);
        QCB_CUDA(cudaGetLastError());
        count_launch();
    }
    D.on = false;
    return QCB_OK;
}

template <int TB, int NW, int MINB_ = 0> static int launch_adjoint_mma(qcb_state *psi, qcb_state *lam, const AdjMmaParams &M, const EnvIds &ids, double *d_env)
{
    constexpr int NT = 32 * NW;
    // two tiles per CTA and at most 80 registers per thread: 6 CTAs = 24 warps per SM at TB = 10 (option "adjoint_ctas" = 5:
    // 96 registers, no spills, 20 warps)
    constexpr int MINB = MINB_ ? MINB_ : (TB >= 12 ? 1 : (TB == 11 ? 2 : (TB == 10 ? (NW == 8 ? 3 : 6) : 6)));
    auto kern = adjoint_pass_mma_kernel<TB, NW, MINB>;
    const size_t smem = 2 * (sizeof(double2) << TB);
    static bool attr_set = false;
    static int ctas_per_sm = 1;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        attr_set = true;
    }
    const unsigned long long ntiles = 1ull << (psi->nloc - TB);
    const unsigned grid = (unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)ctx().sm_count * ctas_per_sm);
    // device work area: [accumulator rows of every warp | their sums | the threads' per-gate programs]
    const size_t nrows = (size_t)grid * NW;
    const size_t rows_doubles = nrows * kAdjMaxGates * kAdjmRow, raw_doubles = (size_t)kAdjMaxGates * kAdjmRow;
    const size_t prog_doubles = (size_t)kAdjMaxGates * 16 * 32 * sizeof(AdjmLaneProg) / sizeof(double);
    // (the rows need no zeroing: first contributions are stored)
    double *d_rows = nullptr, *d_raw = nullptr, *d_prog = nullptr;
    int *id_row = nullptr;
    if (ctx().adj_defer.on) QCB_TRY(adj_defer_area(rows_doubles, (int)nrows, 1, prog_doubles, &d_rows));
    if (ctx().adj_defer.on) {
        const Ctx::AdjDefer &D = ctx().adj_defer;
        d_prog = adj_defer_raw() + (size_t)D.npasses * kAdjMaxGates * 64;
        id_row = adj_defer_table() + D.pass * kAdjMaxGates;
    } else {
        QCB_TRY(ensure_partial(rows_doubles + raw_doubles + prog_doubles));
        d_rows = ctx().d_partial; d_raw = d_rows + rows_doubles; d_prog = d_raw + raw_doubles;
    }
    if (ctx().adjoint_prog_host) {
        // cross-check path: the table built on the host (what the CPU replay runs) and uploaded
        static thread_local std::vector<AdjmLaneProg> prog;
        prog.resize((size_t)kAdjMaxGates * NT);
        build_adj_mma_prog<TB, NW>(M, prog.data());
        QCB_CUDA(cudaMemcpyAsync(d_prog, prog.data(), (size_t)M.A.ngates * NT * sizeof(AdjmLaneProg), cudaMemcpyHostToDevice, ctx().stream));
        if (id_row) {
            k_env_note_ids<<<1, 32, 0, ctx().stream>>>(ids, id_row);
            QCB_CUDA(cudaGetLastError());
            count_launch();
        }
    } else {
        k_build_adj_mma_prog<TB, NW><<<kMaxStages * kSlots, NT, 0, ctx().stream>>>(M, (AdjmLaneProg *)d_prog, ids, id_row);
        QCB_CUDA(cudaGetLastError());
        count_launch();
    }
    kern<<<grid, NT, smem, ctx().stream>>>((double2 *)psi->d, (double2 *)lam->d, M.A.P, M.A.ngates, (const uint4 *)d_prog, ntiles, d_rows);
    QCB_CUDA(cudaGetLastError());
    if (ctx().adj_defer.on) {
        count_launch();
        ctx().adj_defer.pass += 1;
        return QCB_OK;
    }
    if (nrows <= 128) {
        k_env_reduce_raw<<<ids.n, 256, 0, ctx().stream>>>(d_rows, (int)nrows, ids, d_env);
        QCB_CUDA(cudaGetLastError());
        count_launch(2);
    } else {
        k_env_reduce_raw_wide<<<ids.n * kAdjmRow, 128, 0, ctx().stream>>>(d_rows, (int)nrows, d_raw);
        QCB_CUDA(cudaGetLastError());
        k_env_combine_raw<<<1, 32, 0, ctx().stream>>>(d_raw, ids, d_env);
        QCB_CUDA(cudaGetLastError());
        count_launch(3);
    }
    return QCB_OK;
}

template <typename R> static int launch_adjoint_any(qcb_state *psi, qcb_state *lam, int TB, bool mma, AdjParams<R> &A, const EnvIds &ids, double *d_env);

template <> int launch_adjoint_any<float>(qcb_state *psi, qcb_state *lam, int TB, bool, AdjParams<float> &A, const EnvIds &ids, double *d_env)
{
    if (ctx().adjoint_f32) {
        for (int k = 0; k < TB; k++)
            if (A.P.ld_lpos[k] != A.P.st_lpos[k]) return set_error(QCB_EINVAL, "the backward pass does not permute bits");
        switch (TB) {
        case 9: return launch_adjoint_f32<9>(psi, lam, A, ids, d_env);
        case 10: return launch_adjoint_f32<10>(psi, lam, A, ids, d_env);
        case 11: return launch_adjoint_f32<11>(psi, lam, A, ids, d_env);
        case 12: return launch_adjoint_f32<12>(psi, lam, A, ids, d_env);
        default: return set_error(QCB_EINVAL, "unsupported tile size %d for the backward pass", TB);
        }
    }
    switch (TB) {
    case 9: return launch_adjoint<float, 9>(psi, lam, A, ids, d_env);
    case 10: return launch_adjoint<float, 10>(psi, lam, A, ids, d_env);
    case 11: return launch_adjoint<float, 11>(psi, lam, A, ids, d_env);
    case 12: return launch_adjoint<float, 12>(psi, lam, A, ids, d_env);
    default: return set_error(QCB_EINVAL, "unsupported tile size %d for the backward pass", TB);
    }
}

template <> int launch_adjoint_any<double>(qcb_state *psi, qcb_state *lam, int TB, bool mma, AdjParams<double> &A, const EnvIds &ids, double *d_env)
{
    if (mma) {
        static thread_local AdjMmaParams M;
        M.A = A;
        // warps per CTA: 4 keeps 16 windows per warp at TB = 10 (8 MMA groups per gate, 5 CTAs per SM); "adjoint_warps" overrides
        const long want = ctx().adjoint_warps;
        const int NW = TB == 12 ? 16 : (TB == 11 ? 8 : (TB == 9 ? 4 : (want == 4 || want == 8 ? (int)want : 4)));
        for (int k = 0; k < TB; k++)
            if (A.P.ld_lpos[k] != A.P.st_lpos[k]) return set_error(QCB_EINVAL, "the backward pass does not permute bits");
        fill_adj_mma(M, TB, NW);
        switch (TB * 100 + NW) {
        case 904: return launch_adjoint_mma<9, 4>(psi, lam, M, ids, d_env);
        case 1004: return ctx().adjoint_ctas == 5 ? launch_adjoint_mma<10, 4, 5>(psi, lam, M, ids, d_env) : launch_adjoint_mma<10, 4>(psi, lam, M, ids, d_env);
        case 1008: return launch_adjoint_mma<10, 8>(psi, lam, M, ids, d_env);
        case 1108: return launch_adjoint_mma<11, 8>(psi, lam, M, ids, d_env);
        case 1216: return launch_adjoint_mma<12, 16>(psi, lam, M, ids, d_env);
        default: return set_error(QCB_EINVAL, "unsupported tile size %d for the backward pass", TB);
        }
    }
    switch (TB) {
    case 9: return launch_adjoint<double, 9>(psi, lam, A, ids, d_env);
    case 10: return launch_adjoint<double, 10>(psi, lam, A, ids, d_env);
    case 11: return launch_adjoint<double, 11>(psi, lam, A, ids, d_env);
    case 12: return launch_adjoint<double, 12>(psi, lam, A, ids, d_env);
    default: return set_error(QCB_EINVAL, "unsupported tile size %d for the backward pass", TB);
    }
}

// one planned backward pass on (psi, lambda): un-apply its gates on both, environments into d_env rows (GateOp.mat)
template <typename R>
static int launch_adjoint_plan_t(qcb_state *psi, qcb_state *lam, const PassPlan &pl, const std::vector<GateOp> &rev, const double *mats, double *d_env, bool mma)
{
    Ctx &c = ctx();
    static thread_local AdjParams<R> A;
    try {
        fill_params<R>(pl, psi->nloc, rev, mats, true, A.P);
    } catch (const std::exception &e) {
        return set_error(QCB_EINVAL, "pass parameters: %s", e.what());
    }
    EnvIds ids;
    ids.n = 0;
    std::memset(A.slot_gate, 0, sizeof A.slot_gate);
    for (size_t st = 0; st < pl.stages.size(); st++)
        for (int slot = 0; slot < kSlots; slot++) {
            const int g = pl.stages[st].gate[slot];
            if (g < 0) continue;
            if (ids.n >= kAdjMaxGates) return set_error(QCB_EINVAL, "too many gates in a backward pass");
            A.slot_gate[st * kSlots + slot] = (uint8_t)ids.n;
            ids.id[ids.n++] = rev[g].mat;
        }
    A.ngates = ids.n;
    QCB_TRY(launch_adjoint_any<R>(psi, lam, pl.TB, mma, A, ids, d_env));
    c.st_passes += 1;
    c.st_stages += (double)pl.stages.size();
    return QCB_OK;
}

// (sharded gradients, dist.cu: the planner hands out one pass at a time)
bool adjoint_uses_mma(const qcb_state *psi) { return psi->dtype == QCB_C128 && ctx().adjoint_mma != 0; }
SchedOptions adjoint_sched_options(const qcb_state *psi)
{
    // two tiles per CTA: smaller tiles than the forward pass keep more CTAs per SM resident
    SchedOptions opt = current_sched_options(psi);
    opt.max_gates_per_pass = std::min(opt.max_gates_per_pass, kAdjMaxGates);
    opt.tile_bits = ctx().adjoint_tile_bits ? (int)ctx().adjoint_tile_bits : std::min(opt.tile_bits, psi->dtype == QCB_C128 ? 10 : 11);
    return opt;
}
int launch_adjoint_plan(qcb_state *psi, qcb_state *lam, const PassPlan &pl, const std::vector<GateOp> &rev, const double *mats, double *d_env)
{
    if (psi->dtype == QCB_C128) return launch_adjoint_plan_t<double>(psi, lam, pl, rev, mats, d_env, adjoint_uses_mma(psi));
    return launch_adjoint_plan_t<float>(psi, lam, pl, rev, mats, d_env, false);
}

template <typename R> static int run_adjoint_fused_t(qcb_state *psi, qcb_state *lam, const std::vector<GateOp> &rev, const double *mats, double *d_env, bool mma)
{
    Ctx &c = ctx();
    const SchedOptions opt = adjoint_sched_options(psi);
    // (cached like the forward plans: planning costs 0.2-0.4 ms, a fifth of a 20-qubit gradient)
    const std::string key = "adj:" + plan_key(psi, rev, opt);
    auto it = c.plan_cache.find(key);
    if (it == c.plan_cache.end()) {
        std::vector<PassPlan> plans;
        std::vector<char> done;
        try {
            plans = plan_passes(psi->nloc, psi->phys, rev, opt, &done);
        } catch (const std::exception &e) {
            return set_error(QCB_EINVAL, "scheduler: %s", e.what());
        }
        for (char d : done)
            if (!d) return set_error(QCB_EINVAL, "gate list could not be scheduled");
        if (c.plan_cache.size() > 64) c.plan_cache.clear();
        it = c.plan_cache.emplace(key, std::move(plans)).first;
    }
    Ctx::AdjDefer &D = c.adj_defer;
    D = Ctx::AdjDefer();
    D.on = c.adjoint_defer != 0 && it->second.size() > 0;
    D.npasses = (int)it->second.size();
    for (const PassPlan &pl : it->second) {
        const int rc = launch_adjoint_plan_t<R>(psi, lam, pl, rev, mats, d_env, mma);
        if (rc != QCB_OK) { D.on = false; return rc; }
    }
    return adj_defer_finish(d_env);
}

// rev: the circuit's gates in reverse order, GateOp.mat = original gate index (also the row of d_env, 32 doubles each).
// *env_post = false: row g holds env_g[r,c] = sum conj(lam_g[r]) psi_{g-1}[c];  true (ComplexF64 tensor-core pass): row g holds
// M_g[r,c] = sum conj(lam_{g-1}[r]) psi_{g-1}[c], to be turned into env_g = conj(U_g) M_g by the caller (adjm_env_from_post).
int run_adjoint_fused(qcb_state *psi, qcb_state *lam, const std::vector<GateOp> &rev, const double *mats, double *d_env, bool *env_post)
{
    const bool mma = psi->dtype == QCB_C128 && ctx().adjoint_mma != 0;
    if (env_post) *env_post = mma;
    if (psi->dtype == QCB_C128) return run_adjoint_fused_t<double>(psi, lam, rev, mats, d_env, mma);
    return run_adjoint_fused_t<float>(psi, lam, rev, mats, d_env, false);
}

// ---- small states: the whole gate list in one launch on a thread-block cluster (small_circuit.cuh) ------------------------
// index bits that go to the CTA rank (option small_gbits: -1 = by size, else at most that many)
static int small_gbits_for(int n)
{
    const int g = small_circuit_gbits(n);
    const long o = ctx().small_gbits;
    return o < 0 ? g : (int)std::min<long>(g, o);
}

static bool small_cluster_eligible(const qcb_state *s, const std::vector<GateOp> &gates)
{
    const Ctx &c = ctx();
    if (!c.small_cluster || c.force_simple || s->sharded || gates.size() < 2) return false;
    const int kmax = small_circuit_max_qubits((int)amp_bytes(s->dtype));
    if (s->nloc < 2 || s->nloc > (c.small_cluster == 1 ? kmax : std::min<long>(kmax, c.small_cluster))) return false;
    for (int q = 0; q < s->n; q++)
        if (s->phys[q] != q) return false;
    return true;
}

template <typename R> static int run_small_cluster(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint, const void *first_src)
{
    using V = typename Vec2<R>::type;
    Ctx &c = ctx();
    const int G = (int)gates.size();
    const size_t table = (size_t)G * sizeof(GateMat<R>), bytes = table + (size_t)G * sizeof(int);
    if (bytes > c.gates_cap) {
        if (c.gates_ev) cudaEventSynchronize(c.gates_ev);
        if (c.h_gates) cudaFreeHost(c.h_gates);
        if (c.d_gates) cudaFree(c.d_gates);
        c.h_gates = c.d_gates = nullptr; c.gates_cap = 0;
        const size_t cap = std::max<size_t>(bytes * 2, (size_t)256 << 10);
        QCB_CUDA(cudaMallocHost(&c.h_gates, cap));
        QCB_CUDA(cudaMalloc(&c.d_gates, cap));
        c.gates_cap = cap;
    }
    if (!c.gates_ev) QCB_CUDA(cudaEventCreateWithFlags(&c.gates_ev, cudaEventDisableTiming));
    else QCB_CUDA(cudaEventSynchronize(c.gates_ev)); // the previous upload has left the pinned buffer
    GateMat<R> *hg = static_cast<GateMat<R> *>(c.h_gates);
    int *hk = reinterpret_cast<int *>(static_cast<char *>(c.h_gates) + table);
    for (int g = 0; g < G; g++) {
        const double *m = mats + 32 * (size_t)gates[g].mat;
        for (int r = 0; r < 4; r++)
            for (int cc = 0; cc < 4; cc++) {
                const int src = adjoint ? (cc + 4 * r) : (r + 4 * cc);
                gate_set(hg[g], r, cc, m[2 * src], adjoint ? -m[2 * src + 1] : m[2 * src + 1]);
            }
        hk[g] = gates[g].q;
    }
    QCB_CUDA(cudaMemcpyAsync(c.d_gates, c.h_gates, bytes, cudaMemcpyHostToDevice, c.stream));
    QCB_CUDA(cudaEventRecord(c.gates_ev, c.stream));
    const int n = s->nloc, gbits = small_gbits_for(n);
    const size_t smem = small_circuit_smem<R>(n, gbits);
    auto kern = cluster_circuit_kernel<R>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1u << gbits);
    cfg.blockDim = dim3(kSmallThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1u << gbits; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const V *src = (const V *)(first_src ? first_src : s->d);
    V *dst = (V *)s->d;
    const GateMat<R> *dg = static_cast<const GateMat<R> *>(c.d_gates);
    const int *dk = reinterpret_cast<const int *>(static_cast<const char *>(c.d_gates) + table);
    PassTimerScope timer;
    QCB_CUDA(cudaLaunchKernelEx(&cfg, kern, src, dst, dg, dk, n, gbits, G));
    count_launch();
    c.st_passes += 1;
    c.st_cluster_circuits += 1;
    return QCB_OK;
}

// energy + environments of a brickwork circuit in one cluster launch (small_circuit.cuh: cluster_gradient_kernel) + the warp-per-gate
// finish; d_env: 32 doubles per gate (convention of k_adjoint_step), d_energy2: 2 doubles
bool small_gradient_eligible(const qcb_state *s, int ngates)
{
    const Ctx &c = ctx();
    if (!c.small_cluster || c.force_simple || s->sharded || ngates < 1) return false;
    if (s->nloc < 2 || s->nloc > kSmallGradMaxQubits || (c.small_cluster != 1 && s->nloc > c.small_cluster)) return false;
    for (int q = 0; q < s->n; q++)
        if (s->phys[q] != q) return false;
    return true;
}

template <typename R> static int run_small_gradient_t(const qcb_state *s, const std::vector<GateOp> &gates, const double *mats, const HamEval &H,
                                                      double *d_env, double *d_energy2)
{
    using V = typename Vec2<R>::type;
    Ctx &c = ctx();
    const int G = (int)gates.size(), n = s->nloc, gbits = small_gbits_for(n), NR = 1 << gbits;
    const size_t table = (size_t)G * sizeof(GateMat<R>);
    const size_t up = 2 * table + (((size_t)G * sizeof(int) + 15) / 16) * 16;            // U | U^dagger | k
    const size_t bytes = up + ((size_t)NR * G * 32 + (size_t)NR * 32) * sizeof(double);  // + Mpart | Epart (device only)
    if (bytes > c.gates_cap) {
        if (c.gates_ev) cudaEventSynchronize(c.gates_ev);
        if (c.h_gates) cudaFreeHost(c.h_gates);
        if (c.d_gates) cudaFree(c.d_gates);
        c.h_gates = c.d_gates = nullptr; c.gates_cap = 0;
        const size_t cap = std::max<size_t>(bytes * 2, (size_t)256 << 10);
        QCB_CUDA(cudaMallocHost(&c.h_gates, cap));
        QCB_CUDA(cudaMalloc(&c.d_gates, cap));
        c.gates_cap = cap;
    }
    if (!c.gates_ev) QCB_CUDA(cudaEventCreateWithFlags(&c.gates_ev, cudaEventDisableTiming));
    else QCB_CUDA(cudaEventSynchronize(c.gates_ev));
    GateMat<R> *hu = static_cast<GateMat<R> *>(c.h_gates), *hd = hu + G;
    int *hk = reinterpret_cast<int *>(static_cast<char *>(c.h_gates) + 2 * table);
    for (int g = 0; g < G; g++) {
        const double *m = mats + 32 * (size_t)gates[g].mat;
        for (int r = 0; r < 4; r++)
            for (int cc = 0; cc < 4; cc++) {
                gate_set(hu[g], r, cc, m[2 * (r + 4 * cc)], m[2 * (r + 4 * cc) + 1]);
                gate_set(hd[g], r, cc, m[2 * (cc + 4 * r)], -m[2 * (cc + 4 * r) + 1]); // conjugate transpose, src/gradients.jl:27-32
            }
        hk[g] = gates[g].q;
    }
    QCB_CUDA(cudaMemcpyAsync(c.d_gates, c.h_gates, up, cudaMemcpyHostToDevice, c.stream));
    QCB_CUDA(cudaEventRecord(c.gates_ev, c.stream));
    const GateMat<R> *du = static_cast<const GateMat<R> *>(c.d_gates), *dd = du + G;
    const int *dk = reinterpret_cast<const int *>(static_cast<const char *>(c.d_gates) + 2 * table);
    double *Mpart = reinterpret_cast<double *>(static_cast<char *>(c.d_gates) + up), *Epart = Mpart + (size_t)NR * G * 32;
    const size_t smem = small_gradient_smem<R>(n, gbits);
    auto kern = cluster_gradient_kernel<R>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)NR);
    cfg.blockDim = dim3(kSmallThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)NR; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QCB_CUDA(cudaLaunchKernelEx(&cfg, kern, (const V *)s->d, du, dd, dk, n, gbits, G, H, Mpart, Epart));
    k_small_env_finish<R><<<(unsigned)G, 32, 0, c.stream>>>(Mpart, Epart, du, G, NR, d_env, d_energy2);
    QCB_CUDA(cudaGetLastError());
    count_launch(2);
    c.st_cluster_circuits += 1;
    return QCB_OK;
}

int run_small_gradient(const qcb_state *s, const std::vector<GateOp> &gates, const double *mats, const HamEval &H, double *d_env, double *d_energy2)
{
    return s->dtype == QCB_C128 ? run_small_gradient_t<double>(s, gates, mats, H, d_env, d_energy2)
                                : run_small_gradient_t<float>(s, gates, mats, H, d_env, d_energy2);
}

// first_src != nullptr: the state's storage is uninitialised and the circuit's input is the device buffer `first_src`
// (same size and type): the first fused pass reads it directly; paths without such a pass copy it first.
int run_gate_list(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint, const void *first_src)
{
    Ctx &c = ctx();
    c.st_passes = c.st_stages = c.st_launches = c.st_swaps = c.st_fused_swaps = c.st_tma_passes = 0;
    c.st_gates = (double)gates.size();
    for (const GateOp &g : gates)
        if (g.q < 0 || g.q + 1 >= s->n)
            return set_error(QCB_EINVAL, "The gate cannot be applied as it sits outside the qubit space.");
    const bool simple = c.force_simple || s->nloc < kMinTB;
    if (first_src && (gates.empty() || s->sharded || simple)) {
        QCB_CUDA(cudaMemcpyAsync(s->d, first_src, s->bytes, cudaMemcpyDeviceToDevice, c.stream));
        first_src = nullptr;
    }
    if (gates.empty()) return QCB_OK;
    if (s->sharded) return dist_run_gate_list(s, gates, mats, adjoint);
    if (small_cluster_eligible(s, gates))
        return s->dtype == QCB_C128 ? run_small_cluster<double>(s, gates, mats, adjoint, first_src) : run_small_cluster<float>(s, gates, mats, adjoint, first_src);
    if (s->dtype == QCB_C128)
        return simple ? run_simple<double>(s, gates, mats, adjoint) : run_fused<double>(s, gates, mats, adjoint, first_src);
    return simple ? run_simple<float>(s, gates, mats, adjoint) : run_fused<float>(s, gates, mats, adjoint, first_src);
}

} // namespace qcb
