// CPU replay of the fused tile pass (tile.cuh) and the pass scheduler (schedule.hpp).
// TEST INFRASTRUCTURE: runs the very same __host__ __device__ phase functions thread by
// thread on the host, so that the index algebra, swizzle and scheduler are checked against
// the oracle in the CPU-only test suite.  Not part of the product library.
#include <cstdio>
#include <vector>
#include "../../quantumcircuits.jl_b200/csrc/schedule.hpp"

using namespace qcb;

template <typename R, int TB>
static void run_pass(typename Vec2<R>::type *psi, int nloc, const PassParams<R> &P)
{
    using V = typename Vec2<R>::type;
    const uint32_t NT = 1u << (TB - kRegBits);
    const unsigned long long nblocks = 1ull << (nloc - TB);
    std::vector<V> sm(1u << TB);
    for (unsigned long long b = 0; b < nblocks; b++) {
        for (uint32_t t = 0; t < NT; t++) tile_load<R, TB>(psi, P, sm.data(), t, b);
        for (int s = 0; s < P.nstages; s++)
            for (uint32_t t = 0; t < NT; t++) tile_stage<R, TB>(P, s, sm.data(), t);
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
