// CPU replay of the fused tile pass (tile.cuh) and the pass scheduler (schedule.hpp).
// TEST INFRASTRUCTURE: runs the very same __host__ __device__ phase functions thread by
// thread on the host, so that the index algebra, swizzle and scheduler are checked against
// the oracle in the CPU-only test suite.  Not part of the product library.
#include <cstdio>
#include <vector>
#include "../../quantumcircuits.jl_b200/csrc/dist_plan.hpp"
#include "../../quantumcircuits.jl_b200/csrc/tile_async.cuh"
#include "../../quantumcircuits.jl_b200/csrc/small_circuit.cuh"

using namespace qcb;

// ComplexF64 arithmetic of the replayed gate passes: 1 = the 3M form (what the GPU runs for passes of >= 5 gates: the
// straight-line kernel when every stage of the pass is a full diamond, the conditional one otherwise), 0 = the ordinary form
static int g_emu_m3 = 1;
extern "C" void emu_set_m3(int on) { g_emu_m3 = on; }
template <typename R> static int emu_pass_form(const PassParams<R> &P)
{
    if (!g_emu_m3) return 0;
    for (int i = 0; i < P.nstages; i++)
        if (P.stage[i].mask != 15) return 1;
    return 2;
}
template <typename R, int TB> static void emu_stage(const PassParams<R> &P, int form, int s, typename Vec2<R>::type *sm, uint32_t t)
{
    if (form == 2) tile_stage<R, TB, Vec2<R>::SWB, false, 2>(P, s, sm, t);
    else if (form == 1) tile_stage<R, TB, Vec2<R>::SWB, false, 1>(P, s, sm, t);
    else tile_stage<R, TB>(P, s, sm, t);
}

// > 0: replay the persistent form with asynchronous tile loads (tile_async.cuh) on a grid of that many CTAs: two tile
// buffers per CTA, the copies of tile k + 1 issued (here: done) before the stages of tile k
static int g_emu_async_grid = 0;
extern "C" void emu_set_async(int grid) { g_emu_async_grid = grid; }

template <typename R, int TB>
static void run_pass(typename Vec2<R>::type *psi, int nloc, const PassParams<R> &P)
{
    using V = typename Vec2<R>::type;
    const uint32_t NT = 1u << (TB - kRegBits);
    const unsigned long long nblocks = 1ull << (nloc - TB);
    if (g_emu_async_grid > 0) {
        const unsigned long long stride = (unsigned long long)g_emu_async_grid;
        const int form = emu_pass_form(P);
        std::vector<V> ring(2u << TB);
        for (unsigned long long cta = 0; cta < stride && cta < nblocks; cta++) {
            unsigned long long t = cta;
            for (uint32_t tid = 0; tid < NT; tid++) tile_load_async<R, TB>(psi, P, ring.data(), tid, t);
            uint32_t k = 0;
            for (; t < nblocks; t += stride, k ^= 1u) {
                V *sm = ring.data() + ((size_t)k << TB);
                if (t + stride < nblocks)
                    for (uint32_t tid = 0; tid < NT; tid++) tile_load_async<R, TB>(psi, P, ring.data() + ((size_t)(k ^ 1u) << TB), tid, t + stride);
                for (int s = 0; s < P.nstages; s++)
                    for (uint32_t tid = 0; tid < NT; tid++) emu_stage<R, TB>(P, form, s, sm, tid);
                for (uint32_t tid = 0; tid < NT; tid++) tile_store<R, TB>(psi, P, sm, tid, t);
            }
        }
        return;
    }
    std::vector<V> sm(1u << TB);
    for (unsigned long long b = 0; b < nblocks; b++) {
        for (uint32_t t = 0; t < NT; t++) tile_load<R, TB>(psi, P, sm.data(), t, b);
        const int form = emu_pass_form(P);
        for (int s = 0; s < P.nstages; s++)
            for (uint32_t t = 0; t < NT; t++) emu_stage<R, TB>(P, form, s, sm.data(), t);
        for (uint32_t t = 0; t < NT; t++) tile_store<R, TB>(psi, P, sm.data(), t, b);
    }
}

template <typename R>
static int run_all(void *psi_, int nloc, const std::vector<int> &phys, const std::vector<GateOp> &gates, const double *mats,
                   int tile_bits, int low_bits, int max_gates, int adjoint, int *stats)
{
    using V = typename Vec2<R>::type;
    SchedOptions opt;
    opt.tile_bits = tile_bits;
    opt.low_bits = low_bits;
    if (max_gates > 0) opt.max_gates_per_pass = max_gates;
    std::vector<char> done;
    std::vector<PassPlan> passes = plan_passes(nloc, phys, gates, opt, &done);
    for (char d : done)
        if (!d) return 2;
    int nst = 0;
    for (const PassPlan &pl : passes) {
        PassParams<R> *P = new PassParams<R>;
        fill_params<R>(pl, nloc, gates, mats, adjoint != 0, *P);
        nst += P->nstages;
        switch (pl.TB) {
        case 4: return 3;
#define CASE(T) case T: run_pass<R, T>((V *)psi_, nloc, *P); break;
        CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
        default: delete P; return 3;
        }
        delete P;
    }
    if (stats) { stats[0] = (int)passes.size(); stats[1] = nst; }
    return 0;
}

extern "C" int emu_apply_gates(int dtype, int nloc, int nlogical, const int *phys, void *psi, int ngates, const int *q,
                               const double *mats, int tile_bits, int low_bits, int max_gates, int adjoint, int *stats)
{
    std::vector<int> ph(phys, phys + nlogical);
    std::vector<GateOp> gates(ngates);
    for (int g = 0; g < ngates; g++) gates[g] = GateOp{q[g], g};
    try {
        if (dtype) return run_all<double>(psi, nloc, ph, gates, mats, tile_bits, low_bits, max_gates, adjoint, stats);
        return run_all<float>(psi, nloc, ph, gates, mats, tile_bits, low_bits, max_gates, adjoint, stats);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu: %s\n", e.what());
        return 1;
    }
}

// schedule statistics only (no state): passes, stages for a brickwork circuit
extern "C" int emu_plan_brickwork(int nbits, int nlayers, int tile_bits, int low_bits, int max_gates, int *stats)
{
    std::vector<int> pos = brickwork_positions(nbits, nlayers);
    std::vector<GateOp> gates(pos.size());
    for (size_t g = 0; g < pos.size(); g++) gates[g] = GateOp{pos[g], (int)g};
    std::vector<int> phys(nbits);
    for (int i = 0; i < nbits; i++) phys[i] = i;
    SchedOptions opt;
    opt.tile_bits = tile_bits;
    opt.low_bits = low_bits;
    if (max_gates > 0) opt.max_gates_per_pass = max_gates;
    std::vector<PassPlan> passes = plan_passes(nbits, phys, gates, opt);
    int nst = 0, ng = 0;
    for (auto &p : passes) { nst += (int)p.stages.size(); ng += p.ngates; }
    stats[0] = (int)passes.size(); stats[1] = nst; stats[2] = ng;
    return 0;
}

// ---- simulated ranks: P shards in one process, swap_top as a block transposition ----------------
template <typename R> struct EmuShards : ShardExecutor {
    using V = typename Vec2<R>::type;
    int nloc, g;
    std::vector<std::vector<V>> shard; // [rank][2^nloc]
    int npass = 0, nswap = 0, nfused = 0;
    bool fuse = false;    // defer swap_top and fuse it into the next pass (peer reads), like the GPU executor with IPC peers
    bool pending = false;
    template <int TB> void run_pass_impl(V *psi, const PassParams<R> &P) { ::run_pass<R, TB>(psi, nloc, P); }
    template <int TB> void run_pass_peer_impl(const PeerSrc &S, V *dst, const PassParams<R> &P)
    {
        const uint32_t NT = 1u << (TB - kRegBits);
        std::vector<V> sm(1u << TB);
        for (unsigned long long b = 0; b < (1ull << (nloc - TB)); b++) {
            for (uint32_t t = 0; t < NT; t++) tile_load_peer<R, TB>(S, P, sm.data(), t, b);
            const int form = emu_pass_form(P);
            for (int s = 0; s < P.nstages; s++)
                for (uint32_t t = 0; t < NT; t++) emu_stage<R, TB>(P, form, s, sm.data(), t);
            for (uint32_t t = 0; t < NT; t++) tile_store<R, TB>(dst, P, sm.data(), t, b);
        }
    }
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool adjoint) override
    {
        PassParams<R> *P = new PassParams<R>;
        fill_params<R>(pl, nloc, gates, mats, adjoint, *P);
        int rc = 0;
        if (pending) {
            // every rank reads its new-layout tiles from the peers' old shards and writes a fresh local buffer
            const int NR = (int)shard.size();
            std::vector<std::vector<V>> out(NR, std::vector<V>((size_t)1 << nloc));
            for (int me = 0; me < NR && !rc; me++) {
                PeerSrc S;
                S.shift = nloc - g;
                for (int r = 0; r < 8; r++) S.base[r] = r < NR ? (const void *)(shard[r].data() + ((size_t)me << S.shift)) : nullptr;
                switch (pl.TB) {
#define CASE(T) case T: run_pass_peer_impl<T>(S, out[me].data(), *P); break;
                CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
                default: rc = 3;
                }
            }
            shard.swap(out);
            pending = false;
            nfused++;
        } else {
            for (auto &sh : shard) {
                switch (pl.TB) {
#define CASE(T) case T: run_pass_impl<T>(sh.data(), *P); break;
                CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
                default: rc = 3;
                }
            }
        }
        delete P;
        npass++;
        return rc;
    }
    int exchange()
    {
        const int P = (int)shard.size();
        const size_t blk = (size_t)1 << (nloc - g);
        std::vector<std::vector<V>> out = shard;
        for (int r = 0; r < P; r++)
            for (int j = 0; j < P; j++) // block j of rank r <-> block r of rank j
                std::copy(shard[r].begin() + j * blk, shard[r].begin() + (j + 1) * blk, out[j].begin() + r * blk);
        shard.swap(out);
        return 0;
    }
    int swap_top() override
    {
        nswap++;
        if (fuse) { pending = true; return 0; }
        return exchange();
    }
    int flush() override
    {
        if (!pending) return 0;
        pending = false;
        return exchange();
    }
};

// psi: full canonical state of n qubits (complex128), split over P = 2^g simulated ranks.
extern "C" int emu_sharded_apply_gates(int n, int g, void *psi_, int ngates, const int *q, const double *mats, int tile_bits,
                                       int low_bits, int adjoint, int canonical_out, int *phys_out, int *stats)
{
    using V = double2;
    try {
        EmuShards<double> ex;
        ex.nloc = n - g; ex.g = g;
        ex.fuse = (canonical_out & 2) != 0; // bit 1 of the flag: fuse the swaps into the following pass
        canonical_out &= 1;
        const int P = 1 << g;
        const size_t per = (size_t)1 << ex.nloc;
        V *psi = (V *)psi_;
        ex.shard.resize(P);
        for (int r = 0; r < P; r++) ex.shard[r].assign(psi + r * per, psi + (r + 1) * per);
        ShardLayout L;
        L.n = n; L.nloc = ex.nloc; L.g = g;
        L.phys.resize(n);
        for (int i = 0; i < n; i++) L.phys[i] = i;
        std::vector<GateOp> gates(ngates);
        for (int i = 0; i < ngates; i++) gates[i] = GateOp{q[i], i};
        SchedOptions opt;
        opt.tile_bits = tile_bits; opt.low_bits = low_bits;
        double st[3] = {0, 0, 0};
        int rc = run_gates_sharded(ex, L, gates, mats, adjoint != 0, opt, st);
        if (rc) return 10 + (rc < 0 ? -rc : rc);
        if (canonical_out) {
            rc = canonicalize(ex, L, opt);
            if (rc) return 20 + rc;
        }
        for (int r = 0; r < P; r++) std::copy(ex.shard[r].begin(), ex.shard[r].end(), psi + r * per);
        for (int i = 0; i < n; i++) phys_out[i] = L.phys[i];
        if (stats) { stats[0] = (int)st[0]; stats[1] = (int)st[1]; stats[2] = ex.nswap; stats[3] = ex.npass; stats[4] = ex.nfused; }
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu sharded: %s\n", e.what());
        return 1;
    }
}

// planning statistics for a sharded brickwork circuit (no state)
struct NullExec : ShardExecutor {
    int npass = 0, nswap = 0, nperm = 0;
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &, const double *, bool) override { npass++; if (pl.stages.empty()) nperm++; return 0; }
    int swap_top() override { nswap++; return 0; }
};
extern "C" int emu_plan_sharded_brickwork(int nbits, int nlayers, int g, int tile_bits, int low_bits, int *stats)
{
    std::vector<int> pos = brickwork_positions(nbits, nlayers);
    std::vector<GateOp> gates(pos.size());
    for (size_t i = 0; i < pos.size(); i++) gates[i] = GateOp{pos[i], (int)i};
    ShardLayout L;
    L.n = nbits; L.nloc = nbits - g; L.g = g;
    L.phys.resize(nbits);
    for (int i = 0; i < nbits; i++) L.phys[i] = i;
    SchedOptions opt;
    opt.tile_bits = tile_bits; opt.low_bits = low_bits;
    NullExec ex;
    double st[3] = {0, 0, 0};
    struct NoFill : NullExec {};
    int rc = run_gates_sharded(ex, L, gates, nullptr, false, opt, st);
    if (rc) return rc;
    const int before = ex.npass;
    rc = canonicalize(ex, L, opt);
    stats[0] = before; stats[1] = (int)st[1]; stats[2] = ex.nswap; stats[3] = ex.npass - before; stats[4] = ex.nperm;
    return rc;
}

// ---- one real process per rank (tests: torch.distributed gloo, world_size 2): this process holds ONE shard,
// the exchange itself is delegated to the host framework through a callback ----------------------------
typedef int (*emu_swap_cb)(void *shard, int nloc, int g);
struct EmuOneRank : ShardExecutor {
    int nloc, g;
    double2 *shard;
    emu_swap_cb cb;
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool adjoint) override
    {
        PassParams<double> *P = new PassParams<double>;
        fill_params<double>(pl, nloc, gates, mats, adjoint, *P);
        switch (pl.TB) {
#define CASE(T) case T: ::run_pass<double, T>(shard, nloc, *P); break;
        CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
        default: delete P; return 3;
        }
        delete P;
        return 0;
    }
    int swap_top() override { return cb(shard, nloc, g); }
};

extern "C" int emu_rank_apply_gates(int n, int g, void *shard, int ngates, const int *q, const double *mats, int tile_bits,
                                    int low_bits, int canonical_out, emu_swap_cb cb, int *phys_out)
{
    try {
        EmuOneRank ex;
        ex.nloc = n - g; ex.g = g; ex.shard = (double2 *)shard; ex.cb = cb;
        ShardLayout L;
        L.n = n; L.nloc = n - g; L.g = g;
        L.phys.resize(n);
        for (int i = 0; i < n; i++) L.phys[i] = i;
        std::vector<GateOp> gates(ngates);
        for (int i = 0; i < ngates; i++) gates[i] = GateOp{q[i], i};
        SchedOptions opt;
        opt.tile_bits = tile_bits; opt.low_bits = low_bits;
        int rc = run_gates_sharded(ex, L, gates, mats, false, opt, nullptr);
        if (rc) return 10 + (rc < 0 ? -rc : rc);
        if (canonical_out) { rc = canonicalize(ex, L, opt); if (rc) return 20 + rc; }
        for (int i = 0; i < n; i++) phys_out[i] = L.phys[i];
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu rank: %s\n", e.what());
        return 1;
    }
}

// ---- fused backward pass of the adjoint sweep (adjoint.cuh) replayed on the CPU ---------------------------
#include "../../quantumcircuits.jl_b200/csrc/adjoint_mma.cuh"

template <typename R, int TB>
static void run_adjoint_pass(typename Vec2<R>::type *psi, typename Vec2<R>::type *lam, int nloc, const AdjParams<R> &A, double *acc)
{
    using V = typename Vec2<R>::type;
    const uint32_t NT = 1u << (TB - kRegBits);
    const unsigned long long ntiles = 1ull << (nloc - TB);
    std::vector<V> smP(1u << TB), smL(1u << TB);
    for (unsigned long long b = 0; b < ntiles; b++) {
        for (uint32_t t = 0; t < NT; t++) { tile_load<R, TB>(psi, A.P, smP.data(), t, b); tile_load<R, TB>(lam, A.P, smL.data(), t, b); }
        for (int s = 0; s < A.P.nstages; s++)
            for (uint32_t t = 0; t < NT; t++) adjoint_stage<R, TB>(A, s, smP.data(), smL.data(), acc, t);
        for (uint32_t t = 0; t < NT; t++) { tile_store<R, TB>(psi, A.P, smP.data(), t, b); tile_store<R, TB>(lam, A.P, smL.data(), t, b); }
    }
}

// psi, lam: states (modified: both get every gate un-applied, in reverse order); env_out: [ngates][32] indexed like the gate list
template <typename R>
static int adjoint_all(void *psi, void *lam, int n, int ngates, const int *q, const double *mats, int tile_bits, int low_bits, double *env_out)
{
    using V = typename Vec2<R>::type;
    std::vector<GateOp> rev(ngates);
    for (int i = 0; i < ngates; i++) rev[i] = GateOp{q[ngates - 1 - i], ngates - 1 - i};
    std::vector<int> phys(n);
    for (int i = 0; i < n; i++) phys[i] = i;
    SchedOptions opt;
    opt.tile_bits = tile_bits; opt.low_bits = low_bits; opt.max_gates_per_pass = kAdjMaxGates;
    std::vector<char> done;
    std::vector<PassPlan> plans = plan_passes(n, phys, rev, opt, &done);
    for (char d : done) if (!d) return 2;
    AdjParams<R> *A = new AdjParams<R>;
    for (const PassPlan &pl : plans) {
        fill_params<R>(pl, n, rev, mats, true, A->P);
        int ids[kAdjMaxGates], nid = 0;
        std::memset(A->slot_gate, 0, sizeof A->slot_gate);
        for (size_t st = 0; st < pl.stages.size(); st++)
            for (int slot = 0; slot < kSlots; slot++) {
                const int g = pl.stages[st].gate[slot];
                if (g < 0) continue;
                A->slot_gate[st * kSlots + slot] = (uint8_t)nid;
                ids[nid++] = rev[g].mat;
            }
        std::vector<double> acc(kAdjMaxGates * 32, 0.0);
        switch (pl.TB) {
#define CASE(T) case T: run_adjoint_pass<R, T>((V *)psi, (V *)lam, n, *A, acc.data()); break;
        CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
        default: delete A; return 3;
        }
        for (int k = 0; k < nid; k++)
            for (int i = 0; i < 32; i++) env_out[(size_t)ids[k] * 32 + i] = acc[k * 32 + i];
    }
    delete A;
    return 0;
}

extern "C" int emu_adjoint_sweep(int dtype, int n, void *psi, void *lam, int ngates, const int *q, const double *mats, int tile_bits,
                                 int low_bits, double *env_out)
{
    try {
        if (dtype) return adjoint_all<double>(psi, lam, n, ngates, q, mats, tile_bits, low_bits, env_out);
        return adjoint_all<float>(psi, lam, n, ngates, q, mats, tile_bits, low_bits, env_out);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu adjoint: %s\n", e.what());
        return 1;
    }
}

// ---- ComplexF32 backward pass with windows in registers and TF32-split reduction MMAs (adjoint_f32.cuh) -----------------
#include "../../quantumcircuits.jl_b200/csrc/adjoint_f32.cuh"

template <int TB>
static void run_adjoint_pass_f32(float2 *psi, float2 *lam, int nloc, const AdjParams<float> &A, double *acc_out)
{
    const uint32_t NT = 1u << (TB - kRegBits);
    const int NW = (int)NT / 32;
    const unsigned long long ntiles = 1ull << (nloc - TB);
    std::vector<float2> smP(1u << TB), smL(1u << TB);
    std::vector<double> acc((size_t)NW * kAdjMaxGates * 32, 0.0);
    for (unsigned long long b = 0; b < ntiles; b++) {
        for (uint32_t t = 0; t < NT; t++) { tile_load<float, TB>(psi, A.P, smP.data(), t, b); tile_load<float, TB>(lam, A.P, smL.data(), t, b); }
        for (int s = 0; s < A.P.nstages; s++) adjoint_stage_f32_host<TB>(A, s, smP.data(), smL.data(), acc.data());
        for (uint32_t t = 0; t < NT; t++) { tile_store<float, TB>(psi, A.P, smP.data(), t, b); tile_store<float, TB>(lam, A.P, smL.data(), t, b); }
    }
    for (int i = 0; i < kAdjMaxGates * 32; i++) {
        double t = 0;
        for (int w = 0; w < NW; w++) t += acc[(size_t)w * kAdjMaxGates * 32 + i];
        acc_out[i] = t;
    }
}

extern "C" int emu_adjoint_sweep_f32(int n, void *psi, void *lam, int ngates, const int *q, const double *mats, int tile_bits, int low_bits,
                                     double *env_out)
{
    try {
        std::vector<GateOp> rev(ngates);
        for (int i = 0; i < ngates; i++) rev[i] = GateOp{q[ngates - 1 - i], ngates - 1 - i};
        std::vector<int> phys(n);
        for (int i = 0; i < n; i++) phys[i] = i;
        SchedOptions opt;
        opt.tile_bits = tile_bits; opt.low_bits = low_bits; opt.max_gates_per_pass = kAdjMaxGates;
        std::vector<char> done;
        std::vector<PassPlan> plans = plan_passes(n, phys, rev, opt, &done);
        for (char d : done) if (!d) return 2;
        AdjParams<float> *A = new AdjParams<float>;
        int rc = 0;
        for (const PassPlan &pl : plans) {
            fill_params<float>(pl, n, rev, mats, true, A->P);
            int ids[kAdjMaxGates], nid = 0;
            std::memset(A->slot_gate, 0, sizeof A->slot_gate);
            for (size_t st = 0; st < pl.stages.size(); st++)
                for (int slot = 0; slot < kSlots; slot++) {
                    const int g = pl.stages[st].gate[slot];
                    if (g < 0) continue;
                    A->slot_gate[st * kSlots + slot] = (uint8_t)nid;
                    ids[nid++] = rev[g].mat;
                }
            A->ngates = nid;
            std::vector<double> acc(kAdjMaxGates * 32, 0.0);
            switch (pl.TB) {
#define CASE(T) case T: run_adjoint_pass_f32<T>((float2 *)psi, (float2 *)lam, n, *A, acc.data()); break;
            CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
            default: rc = 3;
            }
            if (rc) break;
            for (int k = 0; k < nid; k++)
                for (int i = 0; i < 32; i++) env_out[(size_t)ids[k] * 32 + i] = acc[k * 32 + i];
        }
        delete A;
        return rc;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu adjoint f32: %s\n", e.what());
        return 1;
    }
}

// ---- the tensor-core form of the same pass (adjoint_mma.cuh): software model of the MMA fragment layout ---------------
template <int TB, int NW>
static void run_adjoint_pass_mma(double2 *psi, double2 *lam, int nloc, const AdjMmaParams &M, double *acc_out, long long *conflicts)
{
    std::vector<AdjmLaneProg> prog((size_t)M.A.ngates * 32 * NW);
    build_adj_mma_prog<TB, NW>(M, prog.data());
    const uint32_t NT = 1u << (TB - kRegBits);
    const unsigned long long ntiles = 1ull << (nloc - TB);
    std::vector<double2> smP(1u << TB), smL(1u << TB);
    std::vector<double> rows((size_t)NW * kAdjMaxGates * 64, 0.0);
    for (unsigned long long b = 0; b < ntiles; b++) {
        for (uint32_t t = 0; t < NT; t++) { tile_load<double, TB>(psi, M.A.P, smP.data(), t, b); tile_load<double, TB>(lam, M.A.P, smL.data(), t, b); }
        adjoint_tile_mma_host<TB, NW>(prog.data(), M.A.ngates, smP.data(), smL.data(), rows.data(), conflicts);
        for (uint32_t t = 0; t < NT; t++) { tile_store<double, TB>(psi, M.A.P, smP.data(), t, b); tile_store<double, TB>(lam, M.A.P, smL.data(), t, b); }
    }
    // fixed-order sum of the warps' rows, then the complex 4x4 matrix (what k_env_reduce_raw does on the GPU)
    for (int gi = 0; gi < M.A.ngates; gi++) {
        double raw[64] = {0};
        for (int w = 0; w < NW; w++)
            for (int i = 0; i < 64; i++) raw[i] += rows[((size_t)w * kAdjMaxGates + gi) * 64 + i];
        adjm_combine_raw(raw, acc_out + 32 * gi);
    }
}

// ComplexF64 only.  env_out as in emu_adjoint_sweep (already converted from the post-gate cross-density matrix);
// conflicts[0] = shared-memory wavefronts of the stage loads/stores, conflicts[1] = the conflict-free count.
extern "C" int emu_adjoint_sweep_mma(int n, void *psi, void *lam, int ngates, const int *q, const double *mats, int tile_bits,
                                     int low_bits, int nwarps, double *env_out, long long *conflicts)
{
    try {
        std::vector<GateOp> rev(ngates);
        for (int i = 0; i < ngates; i++) rev[i] = GateOp{q[ngates - 1 - i], ngates - 1 - i};
        std::vector<int> phys(n);
        for (int i = 0; i < n; i++) phys[i] = i;
        SchedOptions opt;
        opt.tile_bits = tile_bits; opt.low_bits = low_bits; opt.max_gates_per_pass = kAdjMaxGates;
        std::vector<char> done;
        std::vector<PassPlan> plans = plan_passes(n, phys, rev, opt, &done);
        for (char d : done) if (!d) return 2;
        AdjMmaParams *M = new AdjMmaParams;
        int rc = 0;
        for (const PassPlan &pl : plans) {
            fill_params<double>(pl, n, rev, mats, true, M->A.P);
            int ids[kAdjMaxGates], nid = 0;
            std::memset(M->A.slot_gate, 0, sizeof M->A.slot_gate);
            for (size_t st = 0; st < pl.stages.size(); st++)
                for (int slot = 0; slot < kSlots; slot++) {
                    const int g = pl.stages[st].gate[slot];
                    if (g < 0) continue;
                    M->A.slot_gate[st * kSlots + slot] = (uint8_t)nid;
                    ids[nid++] = rev[g].mat;
                }
            M->A.ngates = nid;
            fill_adj_mma(*M, pl.TB, nwarps);
            std::vector<double> acc(kAdjMaxGates * 32, 0.0);
            const int key = pl.TB * 100 + nwarps;
            switch (key) {
#define CASE(T, W) case T * 100 + W: run_adjoint_pass_mma<T, W>((double2 *)psi, (double2 *)lam, n, *M, acc.data(), conflicts); break;
            CASE(6, 1) CASE(6, 2) CASE(7, 2) CASE(8, 2) CASE(8, 4) CASE(9, 4) CASE(9, 2) CASE(10, 4) CASE(10, 8) CASE(11, 8) CASE(12, 16)
#undef CASE
            default: rc = 3;
            }
            if (rc) break;
            for (int k = 0; k < nid; k++) {
                double *e = env_out + (size_t)ids[k] * 32;
                for (int i = 0; i < 32; i++) e[i] = acc[k * 32 + i];
                adjm_env_from_post(mats + 32 * (size_t)ids[k], e);
            }
        }
        delete M;
        return rc;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu adjoint mma: %s\n", e.what());
        return 1;
    }
}

// ---- tiled expectation value (measure_tile.cuh + measure_plan.hpp) replayed on the CPU ---------------------
#include "../../quantumcircuits.jl_b200/csrc/measure_plan.hpp"

template <typename R, int TB> static void emu_measure_pass(const typename Vec2<R>::type *psi, int n, const MeasParams<R> &P, double *acc)
{
    const int kind = P.H.kind;
    using V = typename Vec2<R>::type;
    const uint32_t NT = 1u << (TB - kRegBits);
    std::vector<V> sm(1u << TB);
    for (unsigned long long b = 0; b < (1ull << (n - TB)); b++) {
        unsigned long long base = 0;
        for (uint32_t t = 0; t < NT; t++) measure_tile_load<R, TB>(psi, P, sm.data(), t, b, base);
        for (int s = 0; s < P.nstages; s++)
            for (uint32_t t = 0; t < NT; t++) {
                if (kind == 0) measure_stage<R, TB, 0>(P, s, sm.data(), t, base, acc[0], acc[1]);
                else measure_stage<R, TB, 1>(P, s, sm.data(), t, base, acc[0], acc[1]);
            }
    }
}

extern "C" int emu_measure(int dtype, int n, const void *psi, int kind, double p0, double p1, double p2, int tile_bits, int low_bits,
                           double *out_E, int *npasses)
{
    HamEval H{kind, n, p0, p1, p2};
    double acc[2] = {0.0, 0.0};
    try {
        if (dtype) {
            auto passes = plan_measure<double>(n, tile_bits, low_bits, H);
            *npasses = (int)passes.size();
            for (auto &P : passes) switch (tile_bits) {
#define CASE(T) case T: emu_measure_pass<double, T>((const double2 *)psi, n, P, acc); break;
                CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
                default: return 3;
            }
        } else {
            auto passes = plan_measure<float>(n, tile_bits, low_bits, H);
            *npasses = (int)passes.size();
            for (auto &P : passes) switch (tile_bits) {
#define CASE(T) case T: emu_measure_pass<float, T>((const float2 *)psi, n, P, acc); break;
                CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
                default: return 3;
            }
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu measure: %s\n", e.what());
        return 1;
    }
    const double oc = kind == 0 ? -p0 * p2 : p1;
    *out_E = acc[0] + oc * acc[1];
    return 0;
}

// ---- lambda = H psi with the low-bit flips from a staged tile (kernels.cuh: ham_tile_load / ham_tile_rows, k_apply_ham_tiled) ----
template <typename R, int TBL>
static void emu_ham_tiled_run(typename Vec2<R>::type *lam, const typename Vec2<R>::type *psi, const HamEval &H, int grid, double *E2)
{
    using V = typename Vec2<R>::type;
    std::vector<V> sm((size_t)1 << TBL);
    const unsigned long long nchunks = 1ull << (H.n - TBL);
    const double oc = H.offdiag_coef();
    for (int cta = 0; cta < grid; cta++)
        for (unsigned long long ch = (unsigned long long)cta; ch < nchunks; ch += (unsigned long long)grid) {
            for (uint32_t t = 0; t < (uint32_t)kRedThreads; t++) ham_tile_load<R, TBL>(sm.data(), psi, ch << TBL, t);
            for (uint32_t t = 0; t < (uint32_t)kRedThreads; t++) ham_tile_rows<R, TBL>(lam, psi, sm.data(), ch << TBL, t, H, oc, E2[0], E2[1]);
        }
}
extern "C" int emu_apply_ham_tiled(int dtype, int n, const void *psi, int kind, double p0, double p1, double p2, int grid, void *lam, double *E2)
{
    HamEval H{kind, n, p0, p1, p2};
    E2[0] = E2[1] = 0.0;
    if (dtype) {
        if (n < 11) return 3;
        emu_ham_tiled_run<double, 11>((double2 *)lam, (const double2 *)psi, H, grid, E2);
    } else {
        if (n < 12) return 3;
        emu_ham_tiled_run<float, 12>((float2 *)lam, (const float2 *)psi, H, grid, E2);
    }
    return 0;
}

// ---- sharded expectation value with simulated ranks: local passes (rank bits via c_hi), one swap_top, top pass ----------
template <int TBX> static void emu_meas_dispatch(const double2 *psi, int nloc, const MeasParams<double> &P, double *acc)
{
    emu_measure_pass<double, TBX>(psi, nloc, P, acc);
}
extern "C" int emu_sharded_measure(int n, int g, const void *psi_, int kind, double p0, double p1, double p2, int tile_bits, int low_bits,
                                   double *out_E)
{
    try {
        EmuShards<double> ex;
        ex.nloc = n - g; ex.g = g;
        const int P = 1 << g, nloc = n - g;
        const size_t per = (size_t)1 << nloc;
        const double2 *psi = (const double2 *)psi_;
        ex.shard.resize(P);
        for (int r = 0; r < P; r++) ex.shard[r].assign(psi + r * per, psi + (r + 1) * per);
        HamEval H{kind, n, p0, p1, p2};
        double acc[2] = {0.0, 0.0};
        auto run = [&](const MeasParams<double> &MP, const double2 *sh) -> int {
            switch (tile_bits) {
#define CASE(T) case T: emu_meas_dispatch<T>(sh, nloc, MP, acc); return 0;
            CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
            default: return 3;
            }
        };
        for (int r = 0; r < P; r++)
            for (auto &MP : plan_measure<double>(nloc, tile_bits, low_bits, H, (unsigned long long)r << nloc))
                if (run(MP, ex.shard[r].data())) return 3;
        ex.exchange();
        for (int r = 0; r < P; r++)
            if (run(plan_measure_top<double>(nloc, g, tile_bits, low_bits, H, (unsigned long long)r << nloc), ex.shard[r].data())) return 3;
        const double oc = kind == 0 ? -p0 * p2 : p1;
        *out_E = acc[0] + oc * acc[1];
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu sharded measure: %s\n", e.what());
        return 1;
    }
}

// ---- sharded adjoint gradients with simulated ranks: lambda = H psi in two sweeps around one qubit swap (kernels.cuh:
// ham_sharded_amp), then the backward sweep on the (psi, lambda) pair through the sharded planner (what
// gradients_sharded_impl / PairShardExec do on the GPUs with NCCL swaps) -------------------------------------------------
struct EmuPairShards : ShardExecutor {
    int nloc = 0, g = 0;
    std::vector<std::vector<double2>> psi, lam; // [rank][2^nloc]
    std::vector<double> env;                    // [gate][32], summed over the ranks
    int nswap = 0, npass = 0;
    bool forward = false; // forward circuit: plain gate passes and swaps on psi alone
    static void exchange(std::vector<std::vector<double2>> &sh, int nloc, int g)
    {
        const int P = (int)sh.size();
        const size_t blk = (size_t)1 << (nloc - g);
        std::vector<std::vector<double2>> out = sh;
        for (int r = 0; r < P; r++)
            for (int j = 0; j < P; j++) std::copy(sh[r].begin() + j * blk, sh[r].begin() + (j + 1) * blk, out[j].begin() + r * blk);
        sh.swap(out);
    }
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool) override
    {
        npass++;
        if (pl.stages.empty() || forward) { // bit permutation of both states / forward gate pass on psi
            PassParams<double> *P = new PassParams<double>;
            fill_params<double>(pl, nloc, gates, mats, !forward, *P);
            for (size_t r = 0; r < psi.size(); r++)
                switch (pl.TB) {
#define CASE(T) case T: ::run_pass<double, T>(psi[r].data(), nloc, *P); if (!forward) ::run_pass<double, T>(lam[r].data(), nloc, *P); break;
                CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
                default: delete P; return 3;
                }
            delete P;
            return 0;
        }
        AdjParams<double> *A = new AdjParams<double>;
        fill_params<double>(pl, nloc, gates, mats, true, A->P);
        int ids[kAdjMaxGates], nid = 0;
        std::memset(A->slot_gate, 0, sizeof A->slot_gate);
        for (size_t st = 0; st < pl.stages.size(); st++)
            for (int slot = 0; slot < kSlots; slot++) {
                const int gi = pl.stages[st].gate[slot];
                if (gi < 0) continue;
                A->slot_gate[st * kSlots + slot] = (uint8_t)nid;
                ids[nid++] = gates[gi].mat;
            }
        A->ngates = nid;
        int rc = 0;
        for (size_t r = 0; r < psi.size() && !rc; r++) {
            std::vector<double> acc(kAdjMaxGates * 32, 0.0);
            switch (pl.TB) {
#define CASE(T) case T: run_adjoint_pass<double, T>(psi[r].data(), lam[r].data(), nloc, *A, acc.data()); break;
            CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
            default: rc = 3;
            }
            for (int k = 0; k < nid; k++)
                for (int i = 0; i < 32; i++) env[(size_t)ids[k] * 32 + i] += acc[k * 32 + i];
        }
        delete A;
        return rc;
    }
    int swap_top() override
    {
        nswap++;
        exchange(psi, nloc, g);
        if (!forward) exchange(lam, nloc, g);
        return 0;
    }
};

// psi_: full canonical input state psi_0 (complex128); the forward circuit runs sharded first and leaves whatever layout it
// ends in.  env_out[G][32]: sum conj(lam_g[r]) psi_{g-1}[c]; E_out[2]: <psi|H|psi>.
extern "C" int emu_sharded_gradient_env(int n, int g, const void *psi_, int kind, double p0, double p1, double p2, int ngates, const int *q,
                                        const double *mats, int tile_bits, int low_bits, double *env_out, double *E_out, int *stats)
{
    try {
        EmuPairShards ex;
        ex.nloc = n - g; ex.g = g;
        const int P = 1 << g;
        const size_t per = (size_t)1 << ex.nloc;
        const double2 *psi = (const double2 *)psi_;
        ex.psi.resize(P); ex.lam.resize(P);
        for (int r = 0; r < P; r++) { ex.psi[r].assign(psi + r * per, psi + (r + 1) * per); ex.lam[r].assign(per, double2{0.0, 0.0}); }
        ex.env.assign((size_t)ngates * 32, 0.0);
        ShardLayout L;
        L.n = n; L.nloc = ex.nloc; L.g = g;
        L.phys.resize(n);
        for (int i = 0; i < n; i++) L.phys[i] = i;
        SchedOptions fopt;
        fopt.tile_bits = tile_bits; fopt.low_bits = low_bits;
        {
            std::vector<GateOp> fwd(ngates);
            for (int i = 0; i < ngates; i++) fwd[i] = GateOp{q[i], i};
            ex.forward = true;
            const int rc = run_gates_sharded(ex, L, fwd, mats, false, fopt, nullptr);
            ex.forward = false;
            if (rc) return 30 + (rc < 0 ? -rc : rc);
            if (stats) stats[2] = L.canonical() ? 1 : 0;
        }
        HamEval H{kind, n, p0, p1, p2};
        double er = 0, ei = 0;
        auto sweep = [&](unsigned long long mask, bool first) {
            for (int r = 0; r < P; r++) {
                ShardMap M;
                std::memset(&M, 0, sizeof M);
                M.n = n; M.nloc = ex.nloc; M.hi = (unsigned long long)r << ex.nloc;
                for (int i = 0; i < n; i++) M.phys[i] = (unsigned char)L.phys[i];
                // (out of place on the host: every amplitude reads psi only, lambda[x] is its own)
                for (unsigned long long x = 0; x < per; x++) ham_sharded_amp(ex.lam[r].data(), ex.psi[r].data(), x, M, H, mask, first, er, ei);
            }
        };
        unsigned long long local_mask = 0;
        for (int i = 0; i < n; i++)
            if (L.phys[i] < L.nloc) local_mask |= 1ull << i;
        sweep(local_mask, true);
        if (g > 0) {
            ex.swap_top();
            swap_top_layout(L);
            sweep(((1ull << n) - 1ull) & ~local_mask, false);
        }
        E_out[0] = er; E_out[1] = ei;
        std::vector<GateOp> rev(ngates);
        for (int i = 0; i < ngates; i++) rev[i] = GateOp{q[ngates - 1 - i], ngates - 1 - i};
        SchedOptions opt;
        opt.tile_bits = tile_bits; opt.low_bits = low_bits; opt.max_gates_per_pass = kAdjMaxGates;
        const int rc = run_gates_sharded(ex, L, rev, mats, true, opt, nullptr);
        if (rc) return 10 + (rc < 0 ? -rc : rc);
        std::copy(ex.env.begin(), ex.env.end(), env_out);
        if (stats) { stats[0] = ex.npass; stats[1] = ex.nswap; }
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu sharded gradient: %s\n", e.what());
        return 1;
    }
}

// ---- gates_hd.cuh on the host: the per-gate contraction the GPU runs in k_contract_env ----------------------------------
#include "../../quantumcircuits.jl_b200/csrc/gates_hd.cuh"
extern "C" int emu_contract_env_gate(const double *angles15, const double *env32, int post, double *out15, double *u32_out)
{
    for (int layer = 1; layer <= 4; layer++) hd::contract_env_gate(angles15, env32, post, out15, layer); // as the GPU threads do
    double all[15];
    hd::contract_env_gate(angles15, env32, post, all, 0);
    for (int k = 0; k < 15; k++)
        if (all[k] != out15[k]) return 1;
    if (u32_out) {
        hd::GeneralGate G;
        hd::general_gate_setup(angles15, G);
        const hd::M4 U = hd::general_unitary_of(G);
        for (int i = 0; i < 16; i++) { u32_out[2 * i] = U.m[i].re; u32_out[2 * i + 1] = U.m[i].im; }
    }
    return 0;
}

// ---- the same sharded gradient with ONE real process per rank (tests: torch.distributed gloo): this process holds its
// shards of psi and lambda, qubit swaps go through the host framework's all-to-all (callback, once per state), the
// per-rank environments / energy parts are returned for an all-reduce by the caller --------------------------------------
struct EmuOneRankPair : ShardExecutor {
    int nloc = 0, g = 0;
    std::vector<double2> psi, lam;
    std::vector<double> env;
    emu_swap_cb cb = nullptr;
    bool forward = false;
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool) override
    {
        if (pl.stages.empty() || forward) {
            PassParams<double> *P = new PassParams<double>;
            fill_params<double>(pl, nloc, gates, mats, !forward, *P);
            switch (pl.TB) {
#define CASE(T) case T: ::run_pass<double, T>(psi.data(), nloc, *P); if (!forward) ::run_pass<double, T>(lam.data(), nloc, *P); break;
            CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
            default: delete P; return 3;
            }
            delete P;
            return 0;
        }
        AdjParams<double> *A = new AdjParams<double>;
        fill_params<double>(pl, nloc, gates, mats, true, A->P);
        int ids[kAdjMaxGates], nid = 0;
        std::memset(A->slot_gate, 0, sizeof A->slot_gate);
        for (size_t st = 0; st < pl.stages.size(); st++)
            for (int slot = 0; slot < kSlots; slot++) {
                const int gi = pl.stages[st].gate[slot];
                if (gi < 0) continue;
                A->slot_gate[st * kSlots + slot] = (uint8_t)nid;
                ids[nid++] = gates[gi].mat;
            }
        A->ngates = nid;
        std::vector<double> acc(kAdjMaxGates * 32, 0.0);
        int rc = 0;
        switch (pl.TB) {
#define CASE(T) case T: run_adjoint_pass<double, T>(psi.data(), lam.data(), nloc, *A, acc.data()); break;
        CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
        default: rc = 3;
        }
        for (int k = 0; k < nid; k++)
            for (int i = 0; i < 32; i++) env[(size_t)ids[k] * 32 + i] += acc[k * 32 + i];
        delete A;
        return rc;
    }
    int swap_top() override
    {
        int rc = cb(psi.data(), nloc, g);
        if (!rc && !forward) rc = cb(lam.data(), nloc, g);
        return rc;
    }
};

// shard: this rank's canonical slice of psi_0 (not modified).  env_out[G][32], E_out[2]: this rank's partial sums.
extern "C" int emu_rank_gradient_env(int n, int g, int rank, const void *shard, int kind, double p0, double p1, double p2, int ngates, const int *q,
                                     const double *mats, int tile_bits, int low_bits, emu_swap_cb cb, double *env_out, double *E_out)
{
    try {
        EmuOneRankPair ex;
        ex.nloc = n - g; ex.g = g; ex.cb = cb;
        const size_t per = (size_t)1 << ex.nloc;
        ex.psi.assign((const double2 *)shard, (const double2 *)shard + per);
        ex.lam.assign(per, double2{0.0, 0.0});
        ex.env.assign((size_t)ngates * 32, 0.0);
        ShardLayout L;
        L.n = n; L.nloc = ex.nloc; L.g = g;
        L.phys.resize(n);
        for (int i = 0; i < n; i++) L.phys[i] = i;
        SchedOptions fopt;
        fopt.tile_bits = tile_bits; fopt.low_bits = low_bits;
        std::vector<GateOp> fwd(ngates);
        for (int i = 0; i < ngates; i++) fwd[i] = GateOp{q[i], i};
        ex.forward = true;
        int rc = run_gates_sharded(ex, L, fwd, mats, false, fopt, nullptr);
        ex.forward = false;
        if (rc) return 30 + (rc < 0 ? -rc : rc);
        HamEval H{kind, n, p0, p1, p2};
        double er = 0, ei = 0;
        auto sweep = [&](unsigned long long mask, bool first) {
            ShardMap M;
            std::memset(&M, 0, sizeof M);
            M.n = n; M.nloc = ex.nloc; M.hi = (unsigned long long)rank << ex.nloc;
            for (int i = 0; i < n; i++) M.phys[i] = (unsigned char)L.phys[i];
            for (unsigned long long x = 0; x < per; x++) ham_sharded_amp(ex.lam.data(), ex.psi.data(), x, M, H, mask, first, er, ei);
        };
        unsigned long long local_mask = 0;
        for (int i = 0; i < n; i++)
            if (L.phys[i] < L.nloc) local_mask |= 1ull << i;
        sweep(local_mask, true);
        if (g > 0) {
            rc = ex.swap_top();
            if (rc) return 40 + rc;
            swap_top_layout(L);
            sweep(((1ull << n) - 1ull) & ~local_mask, false);
        }
        E_out[0] = er; E_out[1] = ei;
        std::vector<GateOp> rev(ngates);
        for (int i = 0; i < ngates; i++) rev[i] = GateOp{q[ngates - 1 - i], ngates - 1 - i};
        SchedOptions opt;
        opt.tile_bits = tile_bits; opt.low_bits = low_bits; opt.max_gates_per_pass = kAdjMaxGates;
        rc = run_gates_sharded(ex, L, rev, mats, true, opt, nullptr);
        if (rc) return 10 + (rc < 0 ? -rc : rc);
        std::copy(ex.env.begin(), ex.env.end(), env_out);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu rank gradient: %s\n", e.what());
        return 1;
    }
}

// ---- persistent form of the tiled expectation value (measure_tile_persistent_kernel) replayed on the CPU: `grid` CTAs walk
// the tiles with their tile-independent constants hoisted; also the sharded top-bits pass through c_hi is reachable ---------
template <typename R, int TB>
static void emu_measure_pass_persistent(const typename Vec2<R>::type *psi, int n, const MeasParams<R> &P, int grid, double *acc)
{
    using V = typename Vec2<R>::type;
    const int kind = P.H.kind;
    const uint32_t NT = 1u << (TB - kRegBits);
    const unsigned long long ntiles = 1ull << (n - TB);
    std::vector<V> sm(1u << TB);
    for (int b = 0; b < grid; b++) {
        std::vector<MeasThread<R, TB>> T(NT);
        for (uint32_t t = 0; t < NT; t++) T[t] = measure_thread_consts<R, TB>(P, t);
        for (unsigned long long tile = (unsigned long long)b; tile < ntiles; tile += (unsigned long long)grid) {
            const unsigned long long base = measure_tile_base(P, tile);
            for (uint32_t t = 0; t < NT; t++) measure_tile_load_pre<R, TB>(psi, P, sm.data(), T[t], base);
            for (int s = 0; s < kMeasMaxStages; s++) {
                if (s >= P.nstages) continue;
                for (uint32_t t = 0; t < NT; t++) {
                    if (kind == 0) measure_stage_pre<R, TB, 0>(P, s, sm.data(), T[t].es[s], base | P.c_hi | T[t].cth[s], acc[0], acc[1]);
                    else measure_stage_pre<R, TB, 1>(P, s, sm.data(), T[t].es[s], base | P.c_hi | T[t].cth[s], acc[0], acc[1]);
                }
            }
        }
    }
}

extern "C" int emu_measure_persistent(int dtype, int n, const void *psi, int kind, double p0, double p1, double p2, int tile_bits, int low_bits,
                                      int grid, double *out_E)
{
    HamEval H{kind, n, p0, p1, p2};
    double acc[2] = {0.0, 0.0};
    try {
        if (dtype) {
            for (auto &P : plan_measure<double>(n, tile_bits, low_bits, H)) switch (tile_bits) {
#define CASE(T) case T: emu_measure_pass_persistent<double, T>((const double2 *)psi, n, P, grid, acc); break;
                CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
                default: return 3;
            }
        } else {
            for (auto &P : plan_measure<float>(n, tile_bits, low_bits, H)) switch (tile_bits) {
#define CASE(T) case T: emu_measure_pass_persistent<float, T>((const float2 *)psi, n, P, grid, acc); break;
                CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
                default: return 3;
            }
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu measure persistent: %s\n", e.what());
        return 1;
    }
    const double oc = kind == 0 ? -p0 * p2 : p1;
    *out_E = acc[0] + oc * acc[1];
    return 0;
}

// ---- TMA-fed persistent expectation value (flip_tile.cuh): the tile loads are emulated (what the two TMA boxes of a tile
// put into shared memory, 128-byte swizzle included), the consumer code is the device code ------------------------------
#include "../../quantumcircuits.jl_b200/csrc/flip_tile.cuh"

template <typename R, int KIND>
static void emu_flip_run(const typename Vec2<R>::type *psi, const FlipParams &P, int grid, double *acc)
{
    std::vector<unsigned char> tile(kFlipTileBytes + 1024);
    unsigned char *t0 = reinterpret_cast<unsigned char *>(((uintptr_t)tile.data() + 1023) & ~(uintptr_t)1023);
    for (int b = 0; b < grid; b++)
        for (unsigned long long j = b; j < P.n_items; j += grid) {
            const FlipItem it = flip_decode(P, j);
            flip_emulate_tile_load<R>(psi, P, it, t0);
            for (uint32_t tid = 0; tid < (uint32_t)kFlipConsumers; tid++) flip_tile_consume<R, KIND>(t0, tid, P, it, acc[0], acc[1]);
        }
}

// nloc local qubits of an n_total-qubit register (c_hi = the index bits above nloc: simulated rank << nloc)
extern "C" int emu_flip_measure(int dtype, int nloc, int n_total, const void *psi, int kind, double p0, double p1, double p2,
                                unsigned long long c_hi, int grid, double *out2 /* diagonal sum, flip sum */, int *nlevels)
{
    HamEval H{kind, n_total, p0, p1, p2};
    FlipParams P;
    double acc[2] = {0.0, 0.0};
    if (dtype) {
        if (!flip_plan<double>(P, nloc, H, c_hi)) return 2;
        flip_fill_diag_table<double>(P);
        if (kind == 0) emu_flip_run<double, 0>((const double2 *)psi, P, grid, acc);
        else emu_flip_run<double, 1>((const double2 *)psi, P, grid, acc);
    } else {
        if (!flip_plan<float>(P, nloc, H, c_hi)) return 2;
        flip_fill_diag_table<float>(P);
        if (kind == 0) emu_flip_run<float, 0>((const float2 *)psi, P, grid, acc);
        else emu_flip_run<float, 1>((const float2 *)psi, P, grid, acc);
    }
    out2[0] = acc[0];
    out2[1] = acc[1];
    *nlevels = P.nlevels;
    return 0;
}

// ---- the fused gate pass on the TMA-fed kernel (tile_tma.cuh): the box load / store is emulated (ascending tile bits, 128-byte
// hardware swizzle), the stage code is the device code with the stage programs fill_params<double, -3> builds ----------------
#include "../../quantumcircuits.jl_b200/csrc/tile_tma.cuh"

extern "C" int emu_apply_gates_tma(int nloc, void *psi_, int ngates, const int *q, const double *mats, int low_bits, int max_gates, int adjoint, int *stats)
{
    try {
        std::vector<int> phys(nloc);
        for (int i = 0; i < nloc; i++) phys[i] = i;
        std::vector<GateOp> gates(ngates);
        for (int g = 0; g < ngates; g++) gates[g] = GateOp{q[g], g};
        SchedOptions opt;
        opt.tile_bits = kTmaTB;
        opt.low_bits = low_bits;
        if (max_gates > 0) opt.max_gates_per_pass = max_gates;
        std::vector<char> done;
        std::vector<PassPlan> passes = plan_passes(nloc, phys, gates, opt, &done);
        for (char d : done)
            if (!d) return 2;
        double2 *psi = (double2 *)psi_;
        std::vector<double2> sm(1u << kTmaTB);
        PassParams<double> *P = new PassParams<double>;
        int ntma = 0;
        for (const PassPlan &pl : passes) {
            TmaPassGeom g;
            if (!tma_pass_eligible(pl, nloc, g)) { delete P; return 4; } // every pass of a canonical state must be eligible
            ntma++;
            fill_params<double, -3>(pl, nloc, gates, mats, adjoint != 0, *P);
            const int gap_bits = g.a2 - g.b1;
            for (unsigned long long t = 0; t < (1ull << (nloc - kTmaTB)); t++) {
                const unsigned long long base = ((t & ((1ull << gap_bits) - 1ull)) << g.b1) | ((t >> gap_bits) << g.b2);
                auto gidx = [&](uint32_t e) { // box order: [0, b1) then [a2, b2)
                    const unsigned long long lo = e & ((1u << g.b1) - 1u), hi = e >> g.b1;
                    return base | lo | (hi << g.a2);
                };
                for (uint32_t e = 0; e < (1u << kTmaTB); e++) sm[swz<-3>(e)] = psi[gidx(e)];
                for (int s = 0; s < P->nstages; s++)
                    for (uint32_t tid = 0; tid < (uint32_t)kTmaStageThreads; tid++) tile_stage<double, kTmaTB, -3>(*P, s, sm.data(), tid);
                for (uint32_t e = 0; e < (1u << kTmaTB); e++) psi[gidx(e)] = sm[swz<-3>(e)];
            }
        }
        delete P;
        if (stats) { stats[0] = (int)passes.size(); stats[1] = ntma; }
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "emu tma: %s\n", e.what());
        return 1;
    }
}

// ---- cluster-resident small-circuit kernel (small_circuit.cuh): replay with simulated CTAs, same index algebra and arithmetic
template <typename R>
static int emu_small_circuit_impl(void *psi_, int n, int gbits, int ngates, const int *q, const double *mats, int adjoint)
{
    using V = typename Vec2<R>::type;
    const int L = n - gbits;
    if (L < 2) return 2;
    const unsigned nloc = 1u << L, NR = 1u << gbits;
    V *psi = (V *)psi_;
    std::vector<std::vector<V>> A(NR, std::vector<V>(nloc)), B(NR, std::vector<V>(nloc));
    for (unsigned r = 0; r < NR; r++)
        for (unsigned i = 0; i < nloc; i++) A[r][i] = psi[((size_t)r << L) + i];
    for (int g = 0; g < ngates; g++) {
        GateMat<R> U;
        const double *m = mats + 32 * (size_t)g;
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                const int src = adjoint ? (c + 4 * r) : (r + 4 * c);
                gate_set(U, r, c, m[2 * src], adjoint ? -m[2 * src + 1] : m[2 * src + 1]);
            }
        const int k = q[g];
        if (k + 1 < L) {
            for (unsigned rank = 0; rank < NR; rank++)
                for (unsigned qd = 0; qd < (nloc >> 2); qd++) {
                    const unsigned base = small_quad_base(qd, k);
                    V *a = A[rank].data();
                    apply_quad(a[base], a[base | (1u << k)], a[base | (2u << k)], a[base | (3u << k)], U);
                }
        } else {
            for (unsigned rank = 0; rank < NR; rank++)
                for (unsigned a = 0; a < nloc; a++) {
                    unsigned owner[4], loc[4];
                    const int r = small_remote_item(rank, L, a, k, owner, loc);
                    V x[4];
                    for (int c = 0; c < 4; c++) {
                        if (owner[c] >= NR) return 3;
                        x[c] = A[owner[c]][loc[c]];
                    }
                    B[rank][a] = small_row<R, V>(U, r, x);
                }
            A.swap(B);
        }
    }
    for (unsigned r = 0; r < NR; r++)
        for (unsigned i = 0; i < nloc; i++) psi[((size_t)r << L) + i] = A[r][i];
    return 0;
}
extern "C" int emu_small_circuit(int dtype, int n, int gbits, void *psi, int ngates, const int *q, const double *mats, int adjoint)
{
    if (gbits < 0) gbits = small_circuit_gbits(n);
    return dtype ? emu_small_circuit_impl<double>(psi, n, gbits, ngates, q, mats, adjoint) : emu_small_circuit_impl<float>(psi, n, gbits, ngates, q, mats, adjoint);
}
