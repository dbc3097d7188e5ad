// passes.cu -- executes a program-ordered gate list on a device state: plans fused passes
// (schedule.hpp), fills the kernel-parameter blocks and launches tile_pass_kernel (tile.cuh).
// The reference's counterpart is the host loop `for gate: apply!(cache, psi', psi, gate)`
// of src/circuits.jl:19-31 / src/gradients.jl:54-71 with one kernel launch, one zero fill and
// one synchronize per gate (src/apply.jl:134-142).
#include <sstream>

#include "kernels.cuh"
#include "qcb_internal.hpp"

namespace qcb {


template <typename R, int TB> static int launch_pass(qcb_state *s, const PassParams<R> &P)
{
    using V = typename Vec2<R>::type;
    constexpr int NT = 1 << (TB - kRegBits);
    constexpr int MINB = (512 / NT) > 0 ? (512 / NT) : 1;
    auto kern = tile_pass_kernel<R, TB, MINB>;
    const size_t smem = sizeof(V) << TB;
    static bool attr_set = false;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const unsigned long long nblocks = 1ull << (s->nloc - TB);
    kern<<<(unsigned)nblocks, NT, smem, ctx().stream>>>((const V *)s->d, (V *)s->d, P);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    return QCB_OK;
}

template <typename R> static int launch_pass_tb(qcb_state *s, int TB, const PassParams<R> &P)
{
    switch (TB) {
    case 9: return launch_pass<R, 9>(s, P);
    case 10: return launch_pass<R, 10>(s, P);
    case 11: return launch_pass<R, 11>(s, P);
    case 12: return launch_pass<R, 12>(s, P);
    case 13: return launch_pass<R, 13>(s, P);
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

static std::string plan_key(const qcb_state *s, const std::vector<GateOp> &gates, const SchedOptions &o)
{
    // FNV-1a over everything the plan depends on
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { h ^= v; h *= 1099511628211ull; };
    mix((unsigned long long)s->nloc); mix((unsigned long long)o.tile_bits); mix((unsigned long long)o.low_bits);
    mix((unsigned long long)o.max_gates_per_pass); mix((unsigned long long)o.max_stages);
    for (int p : s->phys) mix((unsigned long long)p + 17);
    for (const GateOp &g : gates) { mix((unsigned long long)g.q + 3); mix((unsigned long long)g.mat + 5); }
    std::ostringstream ss;
    ss << gates.size() << ":" << h;
    return ss.str();
}

template <typename R> static int run_fused(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint)
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
        try {
            fill_params<R>(pl, s->nloc, gates, mats, adjoint, P);
        } catch (const std::exception &e) {
            return set_error(QCB_EINVAL, "pass parameters: %s", e.what());
        }
        QCB_TRY(launch_pass_tb<R>(s, pl.TB, P));
        c.st_passes += 1;
        c.st_stages += (double)pl.stages.size();
    }
    return QCB_OK;
}

int run_gate_list(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint)
{
    Ctx &c = ctx();
    c.st_passes = c.st_stages = c.st_launches = c.st_swaps = 0;
    c.st_gates = (double)gates.size();
    if (gates.empty()) return QCB_OK;
    for (const GateOp &g : gates)
        if (g.q < 0 || g.q + 1 >= s->n)
            return set_error(QCB_EINVAL, "The gate cannot be applied as it sits outside the qubit space.");
    if (s->sharded) return dist_run_gate_list(s, gates, mats, adjoint);
    const bool simple = c.force_simple || s->nloc < kMinTB;
    if (s->dtype == QCB_C128)
        return simple ? run_simple<double>(s, gates, mats, adjoint) : run_fused<double>(s, gates, mats, adjoint);
    return simple ? run_simple<float>(s, gates, mats, adjoint) : run_fused<float>(s, gates, mats, adjoint);
}

} // namespace qcb
