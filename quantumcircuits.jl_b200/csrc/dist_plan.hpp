// dist_plan.hpp -- host logic of the sharded (multi-GPU) path: which qubits live in the rank bits,
// when to exchange them, and the bit-permuting passes that prepare an exchange.
//
// No reference counterpart: the reference holds one array in one memory.
// Layout: every rank holds 2^nloc amplitudes; physical index bits [0, nloc) are local, bit nloc+k is
// bit k of the rank.  `phys[q]` maps logical qubit q to its physical bit.  The only inter-rank
// primitive is swap_top: exchange the g rank bits with the top g local bits, an all-to-all of
// contiguous 2^(nloc-g)-amplitude blocks.  Everything else is local tile passes (tile.cuh), which can
// also permute index bits in place (ld_lpos != st_lpos) at the cost of one sweep.
//
// Host only; the CPU replay in tests/emu drives the same code with simulated ranks.
#pragma once
#include "schedule.hpp"

namespace qcb {

struct ShardLayout {
    int n = 0, nloc = 0, g = 0;
    std::vector<int> phys; // logical -> physical
    std::vector<int> logical_of_phys() const
    {
        std::vector<int> l(n, -1);
        for (int q = 0; q < n; q++) l[phys[q]] = q;
        return l;
    }
    bool canonical() const
    {
        for (int q = 0; q < n; q++)
            if (phys[q] != q) return false;
        return true;
    }
};

// executes planned work on every shard (GPU: kernels + NCCL; tests: loops over simulated ranks)
struct ShardExecutor {
    virtual ~ShardExecutor() {}
    virtual int run_pass(const PassPlan &plan, const std::vector<GateOp> &gates, const double *mats, bool adjoint) = 0;
    // exchange rank bit k with local bit nloc-g+k, k = 0..g-1.  An executor may defer the data movement and fuse it
    // into the loads of the next run_pass (peer-memory reads); flush() forces a deferred exchange to happen.
    virtual int swap_top() = 0;
    virtual int flush() { return 0; }
};

// gates executable (dependency-closed, program order) while the qubits flagged in `blocked` are unavailable
inline int count_executable(int n, const std::vector<GateOp> &gates, std::vector<char> dead)
{
    int cnt = 0;
    for (const GateOp &g : gates) {
        if (dead[g.q] || dead[g.q + 1]) dead[g.q] = dead[g.q + 1] = 1;
        else cnt++;
    }
    return cnt;
}

// which g local logical qubits should move into the rank bits next: the contiguous block whose absence
// blocks the fewest pending gates (for a brickwork chain this alternates between the two ends).
inline std::vector<int> choose_new_globals(const ShardLayout &L, const std::vector<GateOp> &pending)
{
    int best = -1, best_s = -1;
    for (int s = 0; s + L.g <= L.n; s++) {
        bool ok = true;
        for (int q = s; q < s + L.g; q++)
            if (L.phys[q] >= L.nloc) ok = false; // must all be local now: swap_top exchanges all g bits at once
        if (!ok) continue;
        std::vector<char> dead(L.n, 0);
        for (int q = s; q < s + L.g; q++) dead[q] = 1;
        const int cnt = count_executable(L.n, pending, dead);
        if (cnt > best) { best = cnt; best_s = s; }
    }
    std::vector<int> v;
    if (best_s >= 0)
        for (int q = best_s; q < best_s + L.g; q++) v.push_back(q);
    return v;
}

// Bit-permuting passes that realise new_phys from L.phys (local bits only).  Each pass permutes the bits of
// one tile: the new occupant of physical bit b is the logical qubit that sat at old physical bit src(b).
inline std::vector<PassPlan> plan_permutation(const ShardLayout &L, const std::vector<int> &new_phys, const SchedOptions &opt)
{
    const int TB = std::min(opt.tile_bits, L.nloc);
    const int c = std::min(opt.low_bits, TB);
    std::vector<int> cur = L.phys;
    std::vector<PassPlan> out;
    for (int guard = 0; guard < 64; guard++) {
        // physical bits whose occupant must change
        std::vector<int> lop(L.n, -1), lnew(L.n, -1);
        for (int q = 0; q < L.n; q++) { lop[cur[q]] = q; lnew[new_phys[q]] = q; }
        std::vector<char> in_tile(L.nloc, 0);
        int used = 0;
        for (int b = 0; b < c; b++) { in_tile[b] = 1; used++; }
        // add whole cycles while they fit
        std::vector<char> visited(L.nloc, 0), moving(L.nloc, 0);
        bool any = false;
        for (int b = 0; b < L.nloc; b++) {
            if (visited[b] || lop[b] == lnew[b]) continue;
            std::vector<int> cyc;
            int x = b;
            while (!visited[x]) { // follow: the qubit now at x must go to new_phys[...]
                visited[x] = 1;
                cyc.push_back(x);
                x = new_phys[lop[x]];
                if (x >= L.nloc) throw std::invalid_argument("permutation moves a qubit into the rank bits");
            }
            int extra = 0;
            for (int y : cyc) extra += in_tile[y] ? 0 : 1;
            if (used + extra > TB) continue;
            for (int y : cyc) { in_tile[y] = 1; moving[y] = 1; }
            used += extra;
            any = true;
        }
        if (!any) {
            bool same = true;
            for (int q = 0; q < L.n; q++) same = same && (cur[q] == new_phys[q]);
            if (same) break;
            throw std::invalid_argument("permutation cycle does not fit one tile");
        }
        for (int b = c; used < TB && b < L.nloc; b++)
            if (!in_tile[b]) { in_tile[b] = 1; used++; }
        PassPlan pl;
        pl.TB = TB;
        for (int b = 0; b < L.nloc; b++)
            if (in_tile[b]) pl.phys.push_back(b);
        std::vector<int> idx(L.nloc, -1);
        for (int k = 0; k < TB; k++) idx[pl.phys[k]] = k;
        pl.ld_lpos.resize(TB);
        pl.st_lpos.resize(TB);
        std::vector<int> next = cur;
        for (int k = 0; k < TB; k++) {
            pl.ld_lpos[k] = k;
            const int b = pl.phys[k];
            // the qubit that must end at b -- only cycles fully inside the tile are moved in this pass
            const int q = lnew[b];
            const int from = moving[b] ? cur[q] : b;
            pl.st_lpos[k] = idx[from];
            if (from != b) next[q] = b;
        }
        cur = next;
        out.push_back(std::move(pl));
    }
    return out;
}

inline void swap_top_layout(ShardLayout &L)
{
    for (int q = 0; q < L.n; q++) {
        const int p = L.phys[q];
        if (p >= L.nloc) L.phys[q] = p - L.g;              // rank bit k -> local bit nloc-g+k
        else if (p >= L.nloc - L.g) L.phys[q] = p + L.g;   // and back
    }
}

// Make the logical qubits V (|V| = g, all local) the rank bits: permute them to the top local bits, swap_top.
inline int make_global(ShardExecutor &ex, ShardLayout &L, const std::vector<int> &V, const SchedOptions &opt)
{
    std::vector<int> np = L.phys;
    std::vector<int> lop = L.logical_of_phys();
    for (int k = 0; k < L.g; k++) {
        const int q = V[k], target = L.nloc - L.g + k;
        if (np[q] == target) continue;
        int q2 = -1;
        for (int x = 0; x < L.n; x++)
            if (np[x] == target) q2 = x;
        std::swap(np[q], np[q2]);
    }
    std::vector<GateOp> none;
    for (const PassPlan &pl : plan_permutation(L, np, opt)) {
        const int rc = ex.run_pass(pl, none, nullptr, false);
        if (rc) return rc;
    }
    L.phys = np;
    const int rc = ex.swap_top();
    if (rc) return rc;
    swap_top_layout(L);
    return 0;
}

// Run a program-ordered gate list on a (possibly sharded) state.  stats: [passes, stages, swaps]
inline int run_gates_sharded(ShardExecutor &ex, ShardLayout &L, const std::vector<GateOp> &gates, const double *mats, bool adjoint,
                             const SchedOptions &opt, double *stats)
{
    std::vector<GateOp> pending = gates;
    for (int guard = 0; !pending.empty(); guard++) {
        if (guard > 4096) return -1;
        std::vector<char> done;
        std::vector<PassPlan> plans = plan_passes(L.nloc, L.phys, pending, opt, &done);
        for (const PassPlan &pl : plans) {
            const int rc = ex.run_pass(pl, pending, mats, adjoint);
            if (rc) return rc;
            if (stats) { stats[0] += 1; stats[1] += (double)pl.stages.size(); }
        }
        std::vector<GateOp> rest;
        for (size_t i = 0; i < pending.size(); i++)
            if (!done[i]) rest.push_back(pending[i]);
        if (rest.empty()) break;
        if (L.g == 0) return -2;
        const std::vector<int> V = choose_new_globals(L, rest);
        if (V.empty()) return -3;
        const int rc = make_global(ex, L, V, opt);
        if (rc) return rc;
        if (stats) stats[2] += 1;
        pending.swap(rest);
    }
    return ex.flush();
}

// Bring the state back to the canonical layout (phys[q] = q): the reference's index convention.
inline int canonicalize(ShardExecutor &ex, ShardLayout &L, const SchedOptions &opt)
{
    if (L.canonical()) return 0;
    std::vector<GateOp> none;
    if (L.g > 0) {
        bool top_global = true;
        for (int k = 0; k < L.g; k++) top_global = top_global && (L.phys[L.n - L.g + k] == L.nloc + k);
        if (!top_global) {
            bool all_local = true;
            for (int k = 0; k < L.g; k++) all_local = all_local && (L.phys[L.n - L.g + k] < L.nloc);
            if (!all_local) {
                // some of the wanted qubits are global but in the wrong rank bits: bring all rank bits down first
                const int rc = ex.swap_top();
                if (rc) return rc;
                swap_top_layout(L);
            }
            std::vector<int> V;
            for (int k = 0; k < L.g; k++) V.push_back(L.n - L.g + k);
            const int rc = make_global(ex, L, V, opt);
            if (rc) return rc;
        }
    }
    std::vector<int> ident(L.n);
    for (int q = 0; q < L.n; q++) ident[q] = q;
    for (const PassPlan &pl : plan_permutation(L, ident, opt)) {
        const int rc = ex.run_pass(pl, none, nullptr, false);
        if (rc) return rc;
    }
    L.phys = ident;
    return ex.flush();
}

} // namespace qcb
