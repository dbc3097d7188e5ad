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
// Roles: the 128 stage threads of tile.cuh; thread 0 also issues the TMA stores and loads between two tiles.
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

// No producer warp: a CTA's warps are allocated in fours, so a fifth warp costs as many registers as four more stage warps (the
// first form, 4 + 1 warps at 136 registers, ran 2 CTAs per SM).  Thread 0 does the TMA work between two tiles: after the CTA
// barrier that ends tile k it issues the store of the slot, waits until the store has READ the slot (a few hundred ns; the other
// slot's tile is already there for the other warps to start on) and refills it with tile k + 2.  Every thread waits for every
// tile in order, so the parity waits need no further bookkeeping.
template <int MINB>
__global__ void __launch_bounds__(kTmaStageThreads, MINB)
tile_pass_tma_kernel(const __grid_constant__ CUtensorMap map_ld, const __grid_constant__ CUtensorMap map_st,
                     const __grid_constant__ PassParams<double> P, int gap_bits, unsigned long long ntiles)
{
    extern __shared__ unsigned char tma_smem_raw[];
    // (pointer arithmetic on the shared array itself: a round trip through uintptr_t makes every tile access a generic LD / ST
    // with a 64-bit address instead of LDS / STS)
    unsigned char *ring = tma_smem_raw + ((1024u - (tma::smem_u32(tma_smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + kTmaSlots * kTmaTileBytes);
    const uint32_t tid = threadIdx.x;
    const unsigned long long gap_mask = (1ull << gap_bits) - 1ull, stride = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < kTmaSlots; i++) tma::mbar_init(&full[i], 1);
        tma::fence_barrier_init();
        tma::prefetch_map(&map_ld);
        tma::prefetch_map(&map_st);
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t i = 0; i < (uint32_t)kTmaSlots; i++) {
            const unsigned long long t = blockIdx.x + i * stride;
            if (t < ntiles) {
                tma::mbar_expect_tx(&full[i], kTmaTileBytes);
                tma::load_5d(ring + (size_t)i * kTmaTileBytes, &map_ld, &full[i], 0, 0, (int)(t & gap_mask), 0, (int)(t >> gap_bits));
            }
        }
    }
    uint32_t k = 0;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += stride, k++) {
        const uint32_t slot = k % kTmaSlots;
        unsigned char *buf = ring + (size_t)slot * kTmaTileBytes;
        double2 *sm = reinterpret_cast<double2 *>(buf);
        tma::mbar_wait(&full[slot], (k / kTmaSlots) & 1u);
#pragma unroll 1
        for (int s = 0; s < P.nstages; s++) {
            if (s) __syncthreads();
            tile_stage<double, kTmaTB, -3, true>(P, s, sm, tid);
        }
        tma::fence_proxy_async(); // this thread's writes -> visible to the TMA store
        __syncthreads();
        if (tid == 0) {
            tma::store_5d(&map_st, buf, 0, 0, (int)(t & gap_mask), 0, (int)(t >> gap_bits));
            tma::store_commit();
            const unsigned long long tn = t + (unsigned long long)kTmaSlots * stride;
            if (tn < ntiles) {
                tma::store_wait_read<0>(); // the store has read the slot: it may be refilled
                tma::mbar_expect_tx(&full[slot], kTmaTileBytes);
                tma::load_5d(buf, &map_ld, &full[slot], 0, 0, (int)(tn & gap_mask), 0, (int)(tn >> gap_bits));
            }
        }
    }
    if (tid == 0) tma::store_wait_all<0>();
}

inline size_t tma_pass_smem_bytes() { return (size_t)kTmaSlots * kTmaTileBytes + 1024 + kTmaSlots * sizeof(uint64_t); }
#endif

} // namespace qcb
