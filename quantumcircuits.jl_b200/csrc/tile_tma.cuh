// tile_tma.cuh -- the fused gate pass (tile.cuh) with Blackwell-native tile movement: a persistent CTA streams its
// tiles through a two-slot shared-memory ring with TMA (cp.async.bulk.tensor loads AND stores, mbarrier phases), so that
// the load of tile k+1 and the write-back of tile k-1 run under the register stages of tile k.  tile_pass_kernel
// (one tile per CTA, LDG -> registers -> STS) overlaps these phases only across the CTAs of an SM.
//
// Same arithmetic, same stage programs (PassParams); only the shared-memory layout differs: the tile is written by the TMA
// unit as 128-byte rows with the hardware 128-byte swizzle (swz<-3>).  Eligible passes: ComplexF64, 2^11-amplitude tiles
// whose index bits are [0, b1) with b1 >= 3 plus at most one more run [a2, b2), in ascending local order, not permuting
// (every pass of an unsharded state); everything else stays on tile_pass_kernel.
//
// Roles: warp 4 lane 0 = producer (loads, and the stores: it waits for "tile computed", issues the store, waits until the
// store has read the slot, then refills it); warps 0..3 = the 128 stage threads of tile.cuh.
#pragma once
#include "tile.cuh"
#include "tma.cuh"

namespace qcb {

constexpr int kTmaTB = 11;
constexpr int kTmaTileBytes = 16 << kTmaTB; // 32 KiB
constexpr int kTmaSlots = 2;
constexpr int kTmaStageThreads = 1 << (kTmaTB - kRegBits); // 128

struct TmaPassGeom {
    int b1 = 0;        // first run of tile bits: [0, b1)
    int a2 = 0, b2 = 0; // second run [a2, b2) (a2 == b2: none)
    int n = 0;
};

// host: can this plan run on the TMA kernel?  (fills the geometry)
inline bool tma_pass_eligible(const PassPlan &pl, int nloc, TmaPassGeom &g)
{
    if (pl.TB != kTmaTB || (int)pl.phys.size() != kTmaTB || nloc < kTmaTB) return false;
    for (int k = 0; k < kTmaTB; k++)
        if (pl.ld_lpos[k] != k || pl.st_lpos[k] != k) return false;
    if (pl.phys[0] != 0) return false;
    int k = 0;
    while (k < kTmaTB && pl.phys[k] == k) k++;
    g.b1 = k;
    if (g.b1 < 3) return false;
    g.a2 = g.b2 = nloc;
    if (k < kTmaTB) {
        g.a2 = pl.phys[k];
        for (int j = k; j < kTmaTB; j++)
            if (pl.phys[j] != g.a2 + (j - k)) return false;
        g.b2 = g.a2 + (kTmaTB - k);
    } else {
        g.a2 = g.b2 = g.b1;
    }
    if (g.b1 - 3 > 8 || g.b2 - g.a2 > 8) return false; // box extents <= 256
    g.n = nloc;
    return true;
}

#if defined(__CUDACC__)
// 5-D view of the state: (128-byte row, rest of run 1, gap, run 2, rest); a tile is one box (1, 2^(b1-3), 1, 2^(b2-a2), 1)
inline int tma_pass_map(CUtensorMap *out, const void *base, const TmaPassGeom &g, const char **why)
{
    TmaDims d;
    d.rank = 5;
    d.dims[0] = 16;
    d.dims[1] = 1ull << (g.b1 - 3);        d.strides[0] = 128;
    d.dims[2] = 1ull << (g.a2 - g.b1);     d.strides[1] = 16ull << g.b1;
    d.dims[3] = 1ull << (g.b2 - g.a2);     d.strides[2] = 16ull << g.a2;
    d.dims[4] = 1ull << (g.n - g.b2);      d.strides[3] = g.b2 < g.n ? (16ull << g.b2) : (16ull << (g.n - 1));
    d.box[0] = 16; d.box[1] = (unsigned)d.dims[1]; d.box[2] = 1; d.box[3] = (unsigned)d.dims[3]; d.box[4] = 1;
    d.strided_rows = g.b1 < 4; // run 1 shorter than 256 bytes
    return tma_encode(out, base, true, d, why);
}

// register budget: the stage body alone fills the 128 registers tile_pass_kernel gets; the ring bookkeeping on top of it
// spilled inside the stage loop under the launch-bound-derived cap (104 bytes of stack).  3 CTAs x 160 threads x 136 <= 64 K.
template <int MINB>
__global__ void __maxnreg__(136)
tile_pass_tma_kernel(const __grid_constant__ CUtensorMap map_ld, const __grid_constant__ CUtensorMap map_st,
                     const __grid_constant__ PassParams<double> P, int gap_bits, unsigned long long ntiles)
{
    extern __shared__ unsigned char tma_smem_raw[];
    unsigned char *ring = reinterpret_cast<unsigned char *>(((uintptr_t)tma_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + kTmaSlots * kTmaTileBytes);
    uint64_t *done = full + kTmaSlots;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kTmaSlots; i++) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&done[i], kTmaStageThreads);
        }
        tma::fence_barrier_init();
    }
    __syncthreads();
    const unsigned long long gap_mask = (1ull << gap_bits) - 1ull;
    if (warp == kTmaStageThreads / 32) {
        if (lane != 0) return;
        tma::prefetch_map(&map_ld);
        tma::prefetch_map(&map_st);
        uint32_t k = 0;
        const unsigned long long back = (unsigned long long)kTmaSlots * gridDim.x; // the tile a slot held before: t - back
        for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x, k++) {
            const uint32_t slot = k % kTmaSlots;
            unsigned char *buf = ring + (size_t)slot * kTmaTileBytes;
            if (k >= kTmaSlots) {
                // the slot holds tile k - 2: wait until its stages are done, write it back, wait until the store has read it
                tma::mbar_wait(&done[slot], ((k / kTmaSlots) - 1) & 1u);
                const unsigned long long p = t - back;
                tma::store_5d(&map_st, buf, 0, 0, (int)(p & gap_mask), 0, (int)(p >> gap_bits));
                tma::store_commit();
                tma::store_wait_read<0>();
            }
            tma::mbar_expect_tx(&full[slot], kTmaTileBytes);
            tma::load_5d(buf, &map_ld, &full[slot], 0, 0, (int)(t & gap_mask), 0, (int)(t >> gap_bits));
        }
        // drain: the last min(k, 2) tiles
        const uint32_t total = k;
        for (uint32_t j = total > kTmaSlots ? total - kTmaSlots : 0; j < total; j++) {
            const uint32_t slot = j % kTmaSlots;
            tma::mbar_wait(&done[slot], (j / kTmaSlots) & 1u);
            const unsigned long long p = blockIdx.x + (unsigned long long)j * gridDim.x;
            tma::store_5d(&map_st, ring + (size_t)slot * kTmaTileBytes, 0, 0, (int)(p & gap_mask), 0, (int)(p >> gap_bits));
            tma::store_commit();
        }
        tma::store_wait_all<0>();
        return;
    }
    uint32_t k = 0;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x, k++) {
        const uint32_t slot = k % kTmaSlots;
        double2 *sm = reinterpret_cast<double2 *>(ring + (size_t)slot * kTmaTileBytes);
        tma::mbar_wait(&full[slot], (k / kTmaSlots) & 1u);
#pragma unroll 1
        for (int s = 0; s < P.nstages; s++) {
            if (s) tma::named_bar_sync(1, kTmaStageThreads);
            tile_stage<double, kTmaTB, -3, true>(P, s, sm, tid);
        }
        tma::fence_proxy_async(); // this thread's writes -> visible to the TMA store
        tma::mbar_arrive(&done[slot]);
    }
}

inline size_t tma_pass_smem_bytes() { return (size_t)kTmaSlots * kTmaTileBytes + 1024 + 2 * kTmaSlots * sizeof(uint64_t); }
#endif

} // namespace qcb
