// tma.cuh -- the Blackwell/Hopper asynchronous-copy plumbing the tile kernels share: mbarrier phases, TMA tensor
// loads and stores (cp.async.bulk.tensor -> SASS UTMALDG / UTMASTG), and the host-side tensor-map encoder (driver entry
// point fetched at run time, so libqcb200.so does not link libcuda).
//
// Shared-memory layout all TMA users of this library rely on: boxes whose innermost extent is exactly 128 bytes,
// CU_TENSOR_MAP_SWIZZLE_128B, destination aligned to 1024 bytes.  Row r of the box (128 bytes) then holds its eight
// 16-byte chunks at positions  chunk ^ (r & 7)  -- `tma_swz_unit` below restates that for a linear 16-byte-unit index.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda.h>
#include <cuda_runtime.h>
#endif

#include "tile.cuh"

namespace qcb {

// unit = 16 bytes; u = (row << 3) | chunk  ->  position of that unit in shared memory (in units)
QCB_HD constexpr uint32_t tma_swz_unit(uint32_t u) { return u ^ ((u >> 3) & 7u); }

#if defined(__CUDACC__)
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA stores), and vice versa
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void store_5d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void store_2d(const CUtensorMap *map, const void *src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent store groups of this thread have finished READING shared memory (the buffer may be reused)
template <int N> __device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) { asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory"); }
// named barrier over a subset of the CTA's warps
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

} // namespace tma

// ---- host: tensor maps --------------------------------------------------------------------------------------------
// A state seen as a tensor of 128-byte rows.  dims[0] is the row (16 doubles / 32 floats); dims[k], k >= 1, are
// power-of-two extents with byte strides `strides[k - 1]`.  box[k] <= 256.
struct TmaDims {
    int rank = 0;
    unsigned long long dims[5] = {1, 1, 1, 1, 1};
    unsigned long long strides[4] = {128, 128, 128, 128}; // bytes, for dims 1..rank-1
    unsigned box[5] = {1, 1, 1, 1, 1};
    // the 128-byte rows of a box are NOT adjacent in memory (strided window tiles): a 256-byte L2 promotion would pull a
    // second, unwanted line out of HBM for every row.  Contiguous boxes keep the 256-byte promotion.
    bool strided_rows = false;
};

// L2 promotion for strided boxes: 0 none, 1 = 128 bytes (default), 2 = 256 bytes (option "tma_l2_promotion")
inline int &tma_strided_promotion()
{
    static int v = 1;
    return v;
}

// returns 0 on success; on failure a short reason in *why (static string)
inline int tma_encode(CUtensorMap *out, const void *base, bool f64, const TmaDims &d, const char **why)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            if (why) *why = "cuTensorMapEncodeTiled is not available from this driver";
            return 1;
        }
        fn = (EncodeFn)p;
    }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < 5; i++) { gd[i] = d.dims[i]; bx[i] = d.box[i]; es[i] = 1; }
    for (int i = 0; i < 4; i++) gs[i] = d.strides[i];
    const int pv = d.strided_rows ? tma_strided_promotion() : 2;
    const CUtensorMapL2promotion promo = pv == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (pv == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
    const CUresult r = fn(out, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)d.rank, const_cast<void *>(base), gd, gs, bx,
                          es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (why) *why = "cuTensorMapEncodeTiled rejected the tensor description";
        return 2;
    }
    return 0;
}
#endif

} // namespace qcb
