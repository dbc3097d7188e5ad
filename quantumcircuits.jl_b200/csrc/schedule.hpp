// schedule.hpp -- host-side pass scheduler.
//
// Turns a program-ordered list of adjacent 2-qubit gates (the reference's host loops:
// apply!(cache, psi', psi, circuit) src/circuits.jl:19-31, propagate_forwards!
// src/gradients.jl:54-71, apply(psi, gates) src/apply.jl:5-18) into a short list of
// fused passes for tile.cuh.  Where the reference launches one kernel per gate, a pass
// applies every gate of a light-cone trapezoid that fits the tile.
//
// Host only, no CUDA dependency: the CPU emulator in tests/emu uses it unchanged.
#pragma once
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "tile.cuh"

namespace qcb {

struct GateOp {
    int q;   // logical low qubit (0-based); the gate acts on logical qubits q and q+1
    int mat; // index of its 4x4 matrix in the caller's matrix array
};

struct StagePlan {
    int p0;
    int gate[kSlots]; // index into the gate list, or -1
};

struct PassPlan {
    int TB = 0;
    std::vector<int> phys;    // tile physical bits, ascending (memory order)
    std::vector<int> ld_lpos; // local position of phys[k] when loading
    std::vector<int> st_lpos; // local position of phys[k] when storing
    std::vector<StagePlan> stages;
    int ngates = 0;
};

struct SchedOptions {
    int tile_bits = 11;
    int low_bits = 3;            // lowest physical bits always in the tile (coalescing)
    int max_stages = kMaxStages;
    int max_gates_per_pass = 1 << 30;
};

// ---------------------------------------------------------------------------------
// stage packing: ordered gates on local adjacent pairs -> register stages
// ---------------------------------------------------------------------------------
inline std::vector<StagePlan> pack_stages(const std::vector<int> &lpos_of_gate, // local low position per collected gate
                                          const std::vector<int> &gate_ids, int TB, int max_stages,
                                          std::vector<char> &scheduled /*out, per collected gate*/)
{
    const int m = (int)lpos_of_gate.size();
    std::vector<char> done(m, 0);
    scheduled.assign(m, 0);
    std::vector<StagePlan> stages;
    int remaining = m;
    while (remaining > 0 && (int)stages.size() < max_stages) {
        int best_cnt = -1;
        StagePlan best{};
        std::vector<int> best_sel;
        for (int p0 = 0; p0 + kRegBits <= TB; p0++) {
            std::vector<char> d = done;
            StagePlan sp;
            sp.p0 = p0;
            std::vector<int> sel;
            int cnt = 0;
            for (int slot = 0; slot < kSlots; slot++) {
                sp.gate[slot] = -1;
                const int want = p0 + slot_lo(slot);
                // first not-yet-done gate at `want` that no earlier pending gate overlaps
                unsigned long long blocked = 0;
                for (int i = 0; i < m; i++) {
                    if (d[i]) continue;
                    const int p = lpos_of_gate[i];
                    const unsigned long long bits = 3ull << p;
                    if (p == want && !(blocked & bits)) {
                        sp.gate[slot] = gate_ids[i];
                        d[i] = 1;
                        sel.push_back(i);
                        cnt++;
                        break;
                    }
                    blocked |= bits;
                }
            }
            if (cnt > best_cnt) { best_cnt = cnt; best = sp; best_sel = sel; }
        }
        if (best_cnt <= 0) break; // cannot happen for adjacent pairs inside the tile
        for (int i : best_sel) { done[i] = 1; scheduled[i] = 1; }
        remaining -= best_cnt;
        stages.push_back(best);
    }
    return stages;
}

// ---------------------------------------------------------------------------------
// pass planning
// ---------------------------------------------------------------------------------
// nloc      : number of local physical index bits (n on one GPU, n - log2(P) when sharded)
// phys[q]   : physical bit of logical qubit q (>= nloc means "global", not addressable here)
// gates     : program order
// returns the passes; `done_out` flags the gates that were scheduled (all of them unless a gate
// touches a global qubit, in which case planning stops at the blocked frontier).
inline std::vector<PassPlan> plan_passes(int nloc, const std::vector<int> &phys, const std::vector<GateOp> &gates,
                                         const SchedOptions &opt, std::vector<char> *done_out = nullptr)
{
    const int n = (int)phys.size();
    const int TB = std::min(opt.tile_bits, nloc);
    if (TB < kRegBits) throw std::invalid_argument("tile smaller than the register window");
    const int c = std::min(opt.low_bits, TB);
    std::vector<int> logical_of_phys(nloc, -1);
    for (int q = 0; q < n; q++)
        if (phys[q] < nloc) logical_of_phys[phys[q]] = q;

    // per-qubit queues
    std::vector<std::vector<int>> queue(n);
    for (int g = 0; g < (int)gates.size(); g++) {
        if (gates[g].q < 0 || gates[g].q + 1 >= n) throw std::invalid_argument("gate outside the qubit space");
        queue[gates[g].q].push_back(g);
        queue[gates[g].q + 1].push_back(g);
    }
    std::vector<int> head(n, 0);
    std::vector<char> done(gates.size(), 0);
    size_t ndone = 0;
    std::vector<PassPlan> passes;

    while (ndone < gates.size()) {
        int best_cnt = 0;
        PassPlan best_plan;
        std::vector<int> best_sched;
        for (int a = 0; a < n; a++) {
            if (phys[a] >= nloc) continue;
            // window of consecutive local logical qubits starting at a
            std::vector<int> window;
            std::vector<char> in_tile(nloc, 0);
            int passive_low = c; // low physical bits not (yet) covered by the window
            int q = a;
            while (q < n && phys[q] < nloc) {
                const int covered = phys[q] < c ? 1 : 0;
                if ((int)window.size() + 1 + passive_low - covered > TB) break;
                window.push_back(q);
                in_tile[phys[q]] = 1;
                passive_low -= covered;
                q++;
            }
            if (window.size() < 2) continue;
            // passive bits: uncovered low bits, then padding with the next lowest free physical bits
            std::vector<int> passive;
            for (int b = 0; b < c; b++)
                if (!in_tile[b]) { passive.push_back(b); in_tile[b] = 1; }
            for (int b = c; (int)(window.size() + passive.size()) < TB && b < nloc; b++)
                if (!in_tile[b]) { passive.push_back(b); in_tile[b] = 1; }
            if ((int)(window.size() + passive.size()) != TB) continue;
            std::sort(passive.begin(), passive.end());
            // local order.  If every passive bit sits below every window bit and the window is in
            // ascending physical order this is the identity; in general passive bits come first.
            std::vector<int> lpos_of_phys(nloc, -1);
            int lp = 0;
            for (int b : passive) lpos_of_phys[b] = lp++;
            for (int w : window) lpos_of_phys[phys[w]] = lp++;
            std::vector<int> lpos_of_logical(n, -1);
            for (int b = 0; b < nloc; b++)
                if (lpos_of_phys[b] >= 0 && logical_of_phys[b] >= 0) lpos_of_logical[logical_of_phys[b]] = lpos_of_phys[b];

            // collect ready gates, round by round
            std::vector<int> h = head;
            std::vector<int> coll, coll_lpos;
            bool progress = true;
            while (progress && (int)coll.size() < opt.max_gates_per_pass) {
                progress = false;
                for (int x = 0; x + 1 < n && (int)coll.size() < opt.max_gates_per_pass; x++) {
                    const int l0 = lpos_of_logical[x], l1 = lpos_of_logical[x + 1];
                    if (l0 < 0 || l1 != l0 + 1) continue;
                    if (h[x] >= (int)queue[x].size() || h[x + 1] >= (int)queue[x + 1].size()) continue;
                    const int g = queue[x][h[x]];
                    if (g != queue[x + 1][h[x + 1]] || gates[g].q != x) continue;
                    coll.push_back(g);
                    coll_lpos.push_back(l0);
                    h[x]++;
                    h[x + 1]++;
                    progress = true;
                }
            }
            if ((int)coll.size() <= best_cnt) continue;
            std::vector<char> sched;
            std::vector<StagePlan> stages = pack_stages(coll_lpos, coll, TB, opt.max_stages, sched);
            int cnt = 0;
            for (char s : sched) cnt += s;
            if (cnt <= best_cnt) continue;
            best_cnt = cnt;
            best_plan = PassPlan();
            best_plan.TB = TB;
            best_plan.stages = stages;
            best_plan.ngates = cnt;
            for (int b = 0; b < nloc; b++)
                if (lpos_of_phys[b] >= 0) {
                    best_plan.phys.push_back(b);
                    best_plan.ld_lpos.push_back(lpos_of_phys[b]);
                }
            best_plan.st_lpos = best_plan.ld_lpos;
            best_sched.clear();
            for (size_t i = 0; i < coll.size(); i++)
                if (sched[i]) best_sched.push_back(coll[i]);
        }
        if (best_cnt == 0) break; // frontier blocked by global qubits
        // commit: advance heads in program order
        std::sort(best_sched.begin(), best_sched.end());
        for (int g : best_sched) {
            done[g] = 1;
            head[gates[g].q]++;
            head[gates[g].q + 1]++;
        }
        ndone += best_sched.size();
        passes.push_back(std::move(best_plan));
    }
    if (done_out) *done_out = done;
    return passes;
}

// ---------------------------------------------------------------------------------
// PassPlan -> kernel parameter block
// ---------------------------------------------------------------------------------
// register window [p0, p0+4): thread bit k -> local position (outside the window), swizzled register offsets.
// The first SWB thread bits get positions with linearly independent bank vectors, which makes the lanes of a quarter
// (half) warp hit distinct bank groups under the XOR swizzle.
template <int SWB> inline void fill_stage_desc(StageDesc &S, int p0, int TB)
{
    const int TBITS = std::min(TB, kMaxTB) - kRegBits;
    S.p0 = (uint8_t)p0;
    S.mask = 0;
    std::vector<int> freepos;
    for (int p = 0; p < TB; p++)
        if (p < p0 || p >= p0 + kRegBits) freepos.push_back(p);
    // greedy: the first SWB thread bits get positions whose bank vectors are linearly independent over GF(2)
    std::vector<int> order;
    std::vector<char> used(TB, 0);
    std::vector<uint32_t> basis; // reduced basis of the chosen vectors
    const int nlane = SWB < 0 ? -SWB : SWB; // thread bits that select the bank group of a quarter / half warp
    for (int p : freepos) {
        if ((int)order.size() >= nlane) break;
        uint32_t v = bank_vector<SWB>(p);
        for (uint32_t bvec : basis) v = std::min(v, v ^ bvec);
        if (v) { basis.push_back(v); order.push_back(p); used[p] = 1; }
    }
    for (int p : freepos)
        if (!used[p]) order.push_back(p);
    for (int k = 0; k < TBITS; k++) S.tpos[k] = (uint8_t)order[k];
    for (int r = 0; r < kRegAmps; r++) S.roff[r] = (uint16_t)swz<SWB>((uint32_t)r << p0);
}

template <typename R, int SWB = Vec2<R>::SWB>
inline void fill_params(const PassPlan &plan, int nloc, const std::vector<GateOp> &gates, const double *mats /*32 doubles per matrix*/,
                        bool adjoint, PassParams<R> &P)
{
    const int TB = plan.TB;
    const int TBITS = TB - kRegBits;
    std::memset(&P, 0, sizeof(P));
    P.nstages = (int)plan.stages.size();
    if (P.nstages > kMaxStages) throw std::invalid_argument("too many stages");
    // CTA index -> non-tile bits, ascending runs
    std::vector<char> in_tile(nloc, 0);
    for (int b : plan.phys) in_tile[b] = 1;
    P.nseg = 0;
    for (int b = 0; b < nloc;) {
        if (in_tile[b]) { b++; continue; }
        int e = b;
        while (e < nloc && !in_tile[e]) e++;
        if (P.nseg >= kMaxSeg) throw std::invalid_argument("tile too fragmented");
        P.seg_start[P.nseg] = (uint8_t)b;
        P.seg_len[P.nseg] = (uint8_t)(e - b);
        P.nseg++;
        b = e;
    }
    for (int k = 0; k < TB; k++) {
        P.ld_phys[k] = (uint8_t)plan.phys[k];
        P.ld_lpos[k] = (uint8_t)plan.ld_lpos[k];
        P.st_lpos[k] = (uint8_t)plan.st_lpos[k];
    }
    for (int i = 0; i < kRegAmps; i++) {
        unsigned long long g = 0;
        uint32_t el = 0, es = 0;
        for (int j = 0; j < kRegBits; j++)
            if ((i >> j) & 1) {
                g |= 1ull << plan.phys[TBITS + j];
                el |= 1u << plan.ld_lpos[TBITS + j];
                es |= 1u << plan.st_lpos[TBITS + j];
            }
        P.ld_goff[i] = g;
        P.ld_eoff[i] = (uint16_t)swz<SWB>(el);
        P.st_eoff[i] = (uint16_t)swz<SWB>(es);
    }
    for (int s = 0; s < P.nstages; s++) {
        const StagePlan &sp = plan.stages[s];
        StageDesc &S = P.stage[s];
        fill_stage_desc<SWB>(S, sp.p0, TB);
        for (int slot = 0; slot < kSlots; slot++) {
            const int g = sp.gate[slot];
            if (g < 0) continue;
            S.mask |= (uint8_t)(1u << slot);
            const double *m = mats + 32 * (size_t)gates[g].mat;
            GateMat<R> &U = P.U[s * kSlots + slot];
            for (int r = 0; r < 4; r++)
                for (int cc = 0; cc < 4; cc++) {
                    if (!adjoint) gate_set(U, r, cc, m[2 * (r + 4 * cc)], m[2 * (r + 4 * cc) + 1]);
                    else gate_set(U, r, cc, m[2 * (cc + 4 * r)], -m[2 * (cc + 4 * r) + 1]); // conjugate transpose, src/gradients.jl:27-32
                }
        }
    }
}

// ---------------------------------------------------------------------------------
// circuit geometry (src/circuits.jl:9-12), 0-based low qubit of every gate in order
// ---------------------------------------------------------------------------------
inline std::vector<int> brickwork_positions(int nbits, int nlayers)
{
    std::vector<int> q;
    for (int l = 1; l <= nlayers; l++)
        for (int j = 1 + (l - 1) % 2; j <= nbits - 1; j += 2) q.push_back(j - 1);
    return q;
}

} // namespace qcb
