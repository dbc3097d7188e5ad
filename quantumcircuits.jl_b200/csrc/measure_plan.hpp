// measure_plan.hpp -- host planning of the tiled expectation-value passes (measure_tile.cuh).
#pragma once
#include <vector>

#include "measure_tile.cuh"
#include "schedule.hpp"

namespace qcb {

// Parameter blocks of all passes over the local index bits [0, n) of a canonical state (physical bit == logical qubit).
// For a sharded state n = nloc, H.n is the full qubit count and c_hi = rank << nloc carries the rank bits into the
// diagonal term and the East control bits; the flips of the qubits in the rank bits are planned by plan_measure_top.
template <typename R> inline std::vector<MeasParams<R>> plan_measure(int n, int TB, int low_bits, const HamEval &H, unsigned long long c_hi = 0)
{
    constexpr int SWB = Vec2<R>::SWB;
    std::vector<MeasParams<R>> out;
    const int c = std::min(low_bits, TB - kRegBits);
    int next = 0; // next qubit whose flip term is still to be covered
    bool first = true;
    while (next < n) {
        MeasParams<R> P;
        std::memset(&P, 0, sizeof P);
        P.H = H;
        P.c_hi = c_hi;
        // tile bits: pass 0 = [0, TB); later passes = low bits [0, c) + window [w0, w0 + TB - c) (shifted down at the top)
        std::vector<int> phys;
        int w0, wlen;
        if (first) {
            for (int b = 0; b < TB; b++) phys.push_back(b);
            w0 = 0; wlen = TB;
        } else {
            wlen = TB - c;
            w0 = std::min(next, n - wlen);
            for (int b = 0; b < c; b++) phys.push_back(b);
            for (int b = w0; b < w0 + wlen; b++) phys.push_back(b);
        }
        std::vector<char> in_tile(n, 0);
        for (int b : phys) in_tile[b] = 1;
        for (int b = 0; b < n;) {
            if (in_tile[b]) { b++; continue; }
            int e = b;
            while (e < n && !in_tile[e]) e++;
            P.seg_start[P.nseg] = (uint8_t)b;
            P.seg_len[P.nseg] = (uint8_t)(e - b);
            P.nseg++;
            b = e;
        }
        const int TBITS = TB - kRegBits;
        for (int k = 0; k < TB; k++) { P.ld_phys[k] = (uint8_t)phys[k]; P.ld_lpos[k] = (uint8_t)k; }
        for (int i = 0; i < kRegAmps; i++) {
            unsigned long long g = 0;
            uint32_t el = 0;
            for (int j = 0; j < kRegBits; j++)
                if ((i >> j) & 1) { g |= 1ull << phys[TBITS + j]; el |= 1u << (TBITS + j); }
            P.ld_goff[i] = g;
            P.ld_eoff[i] = (uint16_t)swz<SWB>(el);
        }
        // stages: 4-bit windows over the local positions that hold the flip qubits of this pass
        const int lp_begin = first ? 0 : c, lp_end = TB; // local positions of the window bits
        std::vector<char> want(TB, 0);
        for (int lp = lp_begin; lp < lp_end; lp++)
            if (phys[lp] >= next) want[lp] = 1;
        int covered_to = next;
        for (int lp = lp_begin; lp < lp_end && P.nstages < kMeasMaxStages;) {
            if (!want[lp]) { lp++; continue; }
            const int p0 = std::min(lp, TB - kRegBits);
            MeasStage &S = P.stage[P.nstages];
            fill_stage_desc<SWB>(S.sd, p0, TB);
            S.active = 0;
            S.diag = (first && P.nstages == 0) ? 1 : 0;
            for (int b = 0; b < kRegBits; b++) {
                S.wphys[b] = (uint8_t)phys[p0 + b];
                S.ctl[b] = (int8_t)(phys[p0 + b] - 1); // canonical layout: qubit i-1 sits one index bit below (-1 for qubit 0)
                if (want[p0 + b]) { S.active |= (uint8_t)(1u << b); want[p0 + b] = 0; covered_to = std::max(covered_to, phys[p0 + b] + 1); }
            }
            P.nstages++;
            lp = p0 + kRegBits;
        }
        // the window is contiguous from `next`, so everything below covered_to is done
        next = covered_to;
        first = false;
        out.push_back(P);
    }
    return out;
}

// Sharded states after swap_top: the g qubits that lived in the rank bits (logical nloc+k) now sit at the top local bits
// nloc-g+k, and the former top local qubits (logical nloc-g+k) are in rank bit k.  One pass adds the flips of those g
// qubits; the East control of logical nloc+k is logical nloc+k-1: the previous top local bit for k >= 1, rank bit g-1
// (index bit nloc+g-1 of c_hi) for k = 0.
template <typename R> inline MeasParams<R> plan_measure_top(int nloc, int g, int TB, int low_bits, const HamEval &H, unsigned long long c_hi)
{
    constexpr int SWB = Vec2<R>::SWB;
    MeasParams<R> P;
    std::memset(&P, 0, sizeof P);
    P.H = H;
    P.c_hi = c_hi;
    const int c = std::min(low_bits, TB - kRegBits);
    std::vector<int> phys;
    for (int b = 0; b < c; b++) phys.push_back(b);
    for (int b = nloc - (TB - c); b < nloc; b++) phys.push_back(b);
    std::vector<char> in_tile(nloc, 0);
    for (int b : phys) in_tile[b] = 1;
    for (int b = 0; b < nloc;) {
        if (in_tile[b]) { b++; continue; }
        int e = b;
        while (e < nloc && !in_tile[e]) e++;
        P.seg_start[P.nseg] = (uint8_t)b;
        P.seg_len[P.nseg] = (uint8_t)(e - b);
        P.nseg++;
        b = e;
    }
    const int TBITS = TB - kRegBits;
    for (int k = 0; k < TB; k++) { P.ld_phys[k] = (uint8_t)phys[k]; P.ld_lpos[k] = (uint8_t)k; }
    for (int i = 0; i < kRegAmps; i++) {
        unsigned long long gg = 0;
        uint32_t el = 0;
        for (int j = 0; j < kRegBits; j++)
            if ((i >> j) & 1) { gg |= 1ull << phys[TBITS + j]; el |= 1u << (TBITS + j); }
        P.ld_goff[i] = gg;
        P.ld_eoff[i] = (uint16_t)swz<SWB>(el);
    }
    MeasStage &S = P.stage[0];
    const int p0 = TB - kRegBits; // window = the top 4 local bits
    fill_stage_desc<SWB>(S.sd, p0, TB);
    S.active = 0;
    S.diag = 0;
    for (int b = 0; b < kRegBits; b++) {
        const int pb = phys[p0 + b];
        S.wphys[b] = (uint8_t)pb;
        S.ctl[b] = -1;
        const int k = pb - (nloc - g); // which former rank-bit qubit sits here
        if (k >= 0) {
            S.active |= (uint8_t)(1u << b);
            S.ctl[b] = (int8_t)(k >= 1 ? pb - 1 : nloc + g - 1);
        }
    }
    P.nstages = 1;
    return P;
}

} // namespace qcb
