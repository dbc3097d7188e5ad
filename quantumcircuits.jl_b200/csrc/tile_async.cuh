// tile_async.cuh -- the fused gate pass (tile.cuh) as a persistent kernel whose tile loads are asynchronous copies issued by
// the stage threads themselves: cp.async (LDGSTS), 16 bytes per copy, global -> shared memory without passing through
// registers, destination = the software-swizzled element the synchronous form (tile_load) stores to.  A CTA owns two tile
// buffers: while it runs the register stages of tile k and writes it back, the copies of tile k + 1 are in flight, so the
// load latency is under the arithmetic of the same CTA instead of only under that of the other CTAs of the SM
// (tile_pass_kernel), and bytes stay in flight on every SM all the time.  Compared with the TMA form (tile_tma.cuh): any pass
// is eligible (same index algebra as tile_load, bit-permuting passes included), the swizzle is the conflict-free one of
// tile.cuh, and nothing goes through the copy engine's 128-byte-row rate.
//
// Same arithmetic, same stage programs and parameter block as tile_pass_kernel (bit-identical results).
#pragma once
#include "tile.cuh"

namespace qcb {

// the load phase of tile.cuh with asynchronous copies (host replay: plain copies)
template <typename R, int TB>
QCB_HD void tile_load_async(const typename Vec2<R>::type *__restrict__ src, const PassParams<R> &P,
                            typename Vec2<R>::type *sm, uint32_t tid, unsigned long long bid)
{
    constexpr int SWB = Vec2<R>::SWB;
    constexpr int TBITS = TB - kRegBits;
    unsigned long long g = cta_base(P, bid);
    uint32_t e = 0;
#pragma unroll
    for (int k = 0; k < TBITS; k++) {
        const uint32_t bit = (tid >> k) & 1u;
        g |= (unsigned long long)bit << P.ld_phys[k];
        e |= bit << P.ld_lpos[k];
    }
    e = swz<SWB>(e);
#pragma unroll
    for (int i = 0; i < kRegAmps; i++) {
#if defined(__CUDA_ARCH__)
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + (e ^ P.ld_eoff[i]));
        if constexpr (sizeof(typename Vec2<R>::type) == 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (g | P.ld_goff[i])) : "memory");
        else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + (g | P.ld_goff[i])) : "memory");
#else
        sm[e ^ P.ld_eoff[i]] = src[g | P.ld_goff[i]];
#endif
    }
}

#if defined(__CUDACC__)
template <typename R, int TB, int MINB, int M3>
__global__ void __launch_bounds__(1 << (TB - kRegBits), MINB)
tile_pass_async_kernel(const typename Vec2<R>::type *src, typename Vec2<R>::type *dst, const __grid_constant__ PassParams<R> P,
                       unsigned long long ntiles)
{
    using V = typename Vec2<R>::type;
    extern __shared__ __align__(16) unsigned char qcb_smem_raw[];
    V *ring = reinterpret_cast<V *>(qcb_smem_raw);
    const uint32_t tid = threadIdx.x;
    const unsigned long long stride = gridDim.x;
    unsigned long long t = blockIdx.x;
    if (t < ntiles) tile_load_async<R, TB>(src, P, ring, tid, t);
    asm volatile("cp.async.commit_group;" ::: "memory");
    uint32_t k = 0;
    for (; t < ntiles; t += stride, k ^= 1u) {
        V *sm = ring + ((size_t)k << TB);
        const unsigned long long tn = t + stride;
        // the other buffer was read for the last time by the write-back of tile k - 1, before the barrier that ended it
        if (tn < ntiles) tile_load_async<R, TB>(src, P, ring + ((size_t)(k ^ 1u) << TB), tid, tn);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory"); // this thread's copies of tile t have landed ...
        __syncthreads();                                     // ... and everybody else's
#pragma unroll 1
        for (int s = 0; s < P.nstages; s++) {
            tile_stage<R, TB, Vec2<R>::SWB, true, M3>(P, s, sm, tid);
            __syncthreads();
        }
        tile_store<R, TB>(dst, P, sm, tid, t);
        __syncthreads();
    }
}
#endif

} // namespace qcb
