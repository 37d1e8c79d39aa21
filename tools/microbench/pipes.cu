// Throughput microbenchmark of the CUDA-core pipes that bound the softmax / depthwise-conv inner loops:
// FFMA vs FFMA2 (packed fp32x2), MUFU.EX2, LDS.128.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters) {
    __shared__ float4 sm[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float a[8]; f32x2 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((f32x2)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i]); }
    const float m = 0.999f; const f32x2 m2 = ((f32x2)__float_as_uint(m) << 32) | __float_as_uint(m);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(m)); }
            if (MODE == 1) { for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, p[i]); }
            if (MODE == 2) { for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (MODE == 3) { for (int i = 0; i < 8; ++i) { float4 v = sm[(threadIdx.x + i * 33 + u + it) & 511]; acc += v.x; } }
        }
    }
    float s = acc; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double ops_per_thread_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 2000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 148.0 * 8 * 256 * iters * ops_per_thread_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-10s %8.3f ms  %8.1f G lane-ops/s  = %6.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, ms, lane_ops / ms / 1e6,
           lane_ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main() { run<0>("FFMA", 64); run<1>("FFMA2", 64); run<2>("MUFU.EX2", 64); run<3>("LDS.128", 64); return 0; }
