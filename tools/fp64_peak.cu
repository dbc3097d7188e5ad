// fp64_peak.cu -- measures the FP64 / FP32 FMA issue roofs of this GPU, and whether FP64 tensor
// (DMMA m8n8k4) work runs concurrently with the FP64 FMA pipe.  SURVEY.md section 8d asks for a measured
// FMA-bound microkernel before any compute-roof fraction is quoted for the fused gate passes.
// Usage: fp64_peak [iters]   -> one JSON line.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

template <typename T, int ILP> __global__ void __launch_bounds__(256) k_fma(T *out, int iters, T a, T b)
{
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m8n8k4: D(8x8) += A(8x4) B(4x8), one instruction = 256 FMA per warp
template <int ILP> __global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a, double b)
{
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half of the warps do DFMA, the other half DMMA
template <int ILP> __global__ void __launch_bounds__(256) k_mixed(double *out, int iters, double a, double b)
{
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double c0[ILP], c1[ILP];
#pragma unroll
        for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x; c1[i] = i; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    } else {
        double acc[ILP];
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = (double)(threadIdx.x + i);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < ILP; i++) s += acc[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed FP32 (FFMA2): operands from registers vs from the kernel-parameter bank (uniform registers)
template <int ILP, bool UNIFORM> __global__ void __launch_bounds__(256) k_ffma2(float2 *out, int iters, float2 a, float2 b)
{
    float2 acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
    float2 ra = a, rb = b;
    if (!UNIFORM) { ra.x += (float)threadIdx.x * 1e-30f; rb.y += (float)threadIdx.x * 1e-30f; } // force per-thread registers
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = __ffma2_rn(acc[i], ra, rb);
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ILP; i++) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy tensor path in single precision: mma.sync m16n8k8 TF32 (1024 FMA per warp instruction), FP32 accumulate
template <int ILP> __global__ void __launch_bounds__(256) k_tf32(float *out, int iters, unsigned a, unsigned b)
{
    float c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1.f; c[i][3] = 2.f; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a), "r"(b), "r"(a), "r"(b), "r"(b), "r"(a));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mma.sync m16n8k16 BF16 (2048 FMA per warp instruction), FP32 accumulate
template <int ILP> __global__ void __launch_bounds__(256) k_bf16(float *out, int iters, unsigned a, unsigned b)
{
    float c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1.f; c[i][3] = 2.f; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a), "r"(b), "r"(a), "r"(b), "r"(b), "r"(a));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// every warp interleaves FFMA2 with TF32 MMAs (8 FFMA2 per MMA, the ratio of the ComplexF32 backward pass): do the two
// pipes overlap?  Reports the FFMA2 rate with the MMAs in the stream.
template <int ILP> __global__ void __launch_bounds__(256) k_ffma2_tf32(float2 *out, int iters, float2 a, float2 b, unsigned ua, unsigned ub)
{
    float2 acc[ILP];
    float c[2][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
#pragma unroll
    for (int i = 0; i < 2; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1.f; c[i][3] = 2.f; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            acc[i] = __ffma2_rn(acc[i], a, b);
            if ((i & 7) == 7)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(c[(i >> 3) & 1][0]), "+f"(c[(i >> 3) & 1][1]), "+f"(c[(i >> 3) & 1][2]), "+f"(c[(i >> 3) & 1][3])
                             : "r"(ua), "r"(ub), "r"(ua), "r"(ub), "r"(ub), "r"(ua));
        }
    }
    float2 s = make_float2(c[0][0] + c[1][1], c[0][2] + c[1][3]);
#pragma unroll
    for (int i = 0; i < ILP; i++) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 4096;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    constexpr int ILP = 16;
    const double nthreads = (double)blocks * threads;
    float ms_d = time_ms([&] { k_fma<double, ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float ms_f = time_ms([&] { k_fma<float, ILP><<<blocks, threads>>>((float *)out, iters, 1.0000001f, 1e-9f); });
    float ms_m = time_ms([&] { k_dmma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float ms_x = time_ms([&] { k_mixed<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float ms_p2r = time_ms([&] { k_ffma2<ILP, false><<<blocks, threads>>>((float2 *)out, iters, make_float2(1.0000001f, 0.9999999f), make_float2(1e-9f, 1e-9f)); });
    float ms_p2u = time_ms([&] { k_ffma2<ILP, true><<<blocks, threads>>>((float2 *)out, iters, make_float2(1.0000001f, 0.9999999f), make_float2(1e-9f, 1e-9f)); });
    float ms_t = time_ms([&] { k_tf32<8><<<blocks, threads>>>((float *)out, iters, 0x3f800000u, 0x3f000000u); });
    float ms_b = time_ms([&] { k_bf16<8><<<blocks, threads>>>((float *)out, iters, 0x3f803f80u, 0x3f003f00u); });
    float ms_ft = time_ms([&] { k_ffma2_tf32<ILP><<<blocks, threads>>>((float2 *)out, iters, make_float2(1.0000001f, 0.9999999f), make_float2(1e-9f, 1e-9f), 0x3f800000u, 0x3f000000u); });
    CK(cudaGetLastError());
    const double tf32_tflops = 2.0 * (nthreads / 32) * iters * 8 * 1024.0 / (ms_t * 1e-3) / 1e12;
    const double bf16_tflops = 2.0 * (nthreads / 32) * iters * 8 * 2048.0 / (ms_b * 1e-3) / 1e12;
    const double ffma2_with_tf32_tflops = 2.0 * 2.0 * nthreads * iters * ILP / (ms_ft * 1e-3) / 1e12;
    const double fma_p2r = 2.0 * nthreads * iters * ILP / (ms_p2r * 1e-3), fma_p2u = 2.0 * nthreads * iters * ILP / (ms_p2u * 1e-3);
    const double fma_d = nthreads * iters * ILP / (ms_d * 1e-3);
    const double fma_f = nthreads * iters * ILP / (ms_f * 1e-3);
    const double fma_m = (nthreads / 32) * iters * ILP * 256.0 / (ms_m * 1e-3);
    const double fma_x = ((nthreads / 2) * iters * ILP + (nthreads / 64) * iters * ILP * 256.0) / (ms_x * 1e-3);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"ffma_tflops\": %.2f, \"dmma_tflops\": %.2f, "
           "\"mixed_dfma_dmma_tflops\": %.2f, \"ffma2_reg_operand_tflops\": %.2f, \"ffma2_uniform_operand_tflops\": %.2f, "
           "\"mma_sync_tf32_tflops\": %.1f, \"mma_sync_bf16_tflops\": %.1f, \"ffma2_tflops_with_1_tf32_mma_per_8_ffma2\": %.2f, "
           "\"dfma_per_clk_per_sm_at_max_clock\": %.1f, \"ms\": [%.3f, %.3f, %.3f, %.3f]}\n",
           p.name, sms, 2 * fma_d / 1e12, 2 * fma_f / 1e12, 2 * fma_m / 1e12, 2 * fma_x / 1e12, 2 * fma_p2r / 1e12, 2 * fma_p2u / 1e12,
           tf32_tflops, bf16_tflops, ffma2_with_tf32_tflops,
           fma_d / sms / (p.clockRate * 1e3), ms_d, ms_f, ms_m, ms_x);
    return 0;
}
