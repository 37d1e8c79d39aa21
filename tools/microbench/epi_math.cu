// Throughput of candidate epilogue activations (SwooshL on fp32 accumulators -> packed fp16), per SM:
// how many elements/clk the CUDA-core pipes sustain with the 8 epilogue warps of the GEMM (2 per SMSP)
// and with full occupancy.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_math epi_math.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
constexpr float L2E = 1.4426950408889634f, C = 4.0f, K0 = -(0.08f * 4.0f + 0.035f);
// V0: one MUFU + degree-5 log1p polynomial (current)
__device__ __forceinline__ float v0(float x) {
    const float z = fmaf(x, L2E, -C * L2E);
    const float t = ex2(-fabsf(z));
    float p = fmaf(t, 0.031377589387161245f, -0.1341354334221127f);
    p = fmaf(t, p, 0.2878262894239249f); p = fmaf(t, p, -0.491347927069251f); p = fmaf(t, p, 0.9994349844843187f);
    float r = fmaf(t, p, K0 - 0.42f * C);
    r = fmaf(x, 0.42f, r);
    return fmaf(fabsf(z), 0.5f / L2E, r);
}
// V1: two MUFU (ex2 + lg2)
__device__ __forceinline__ float v1(float x) {
    const float z = fmaf(x, L2E, -C * L2E);
    const float t = ex2(-fabsf(z));
    const float l = lg2(1.0f + t);
    float r = fmaf(x, 0.42f, K0 - 0.42f * C);
    r = fmaf(fabsf(z), 0.5f / L2E, r);
    return fmaf(l, 0.6931471805599453f, r);
}
// V2: V0 on packed pairs (FFMA2)
__device__ __forceinline__ void v2(float x0, float x1, float& o0, float& o1) {
    const f32x2 x = pack2(x0, x1);
    const f32x2 z = fma2(x, pack2(L2E, L2E), pack2(-C * L2E, -C * L2E));
    float z0, z1; unpack2(z, z0, z1);
    const float t0 = ex2(-fabsf(z0)), t1 = ex2(-fabsf(z1));
    const f32x2 t = pack2(t0, t1);
    f32x2 p = fma2(t, pack2(0.031377589387161245f, 0.031377589387161245f), pack2(-0.1341354334221127f, -0.1341354334221127f));
    p = fma2(t, p, pack2(0.2878262894239249f, 0.2878262894239249f));
    p = fma2(t, p, pack2(-0.491347927069251f, -0.491347927069251f));
    p = fma2(t, p, pack2(0.9994349844843187f, 0.9994349844843187f));
    f32x2 r = fma2(t, p, pack2(K0 - 0.42f * C, K0 - 0.42f * C));
    r = fma2(x, pack2(0.42f, 0.42f), r);
    r = fma2(pack2(fabsf(z0), fabsf(z1)), pack2(0.5f / L2E, 0.5f / L2E), r);
    unpack2(r, o0, o1);
}
// V3: one MUFU, degree-3 polynomial in t for log1p(t)/t is not accurate enough; instead exploit that for
// z << 0 (the common case: x - 4 < -2) log1p(t) ~ t - t^2/2 + t^3/3: degree-3 (max error 1e-3 at t = 1 -> reject).
// Kept only as an instruction-count probe.
__device__ __forceinline__ float v3(float x) {
    const float z = fmaf(x, L2E, -C * L2E);
    const float t = ex2(-fabsf(z));
    float p = fmaf(t, 0.10f, -0.40f);
    p = fmaf(t, p, 0.99f);
    float r = fmaf(t, p, K0 - 0.42f * C);
    r = fmaf(x, 0.42f, r);
    return fmaf(fabsf(z), 0.5f / L2E, r);
}
template <int MODE> __global__ void __launch_bounds__(256) k(const float* __restrict__ in, uint32_t* __restrict__ out, int iters, float bias) {
    float a[32];
    for (int i = 0; i < 32; ++i) a[i] = in[(threadIdx.x * 32 + i) & 1023];
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(a[i], 1.0001f, bias + it);      // acc * rscale + bias
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v0(v[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v1(v[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) v2(v[i], v[i + 1], v[i], v[i + 1]);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v3(v[i]);
        } else if (MODE == 4) {      // half the elements with V0, half with V1: balances MUFU and FMA pipes
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (i & 1) ? v1(v[i]) : v0(v[i]);
        } else if (MODE == 5) {      // 3 of 4 with V0
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (i & 3) == 3 ? v1(v[i]) : v0(v[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) acc ^= pack_h2(v[i], v[i + 1]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, int blocks_per_sm) {
    float* in; uint32_t* out; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * blocks_per_sm, 256>>>(in, out, 10, -300.f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * blocks_per_sm, 256>>>(in, out, iters, -2000.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double elems = 148.0 * blocks_per_sm * 256 * (double)iters * 32;
    printf("%-28s warps/SM=%2d  %8.3f ms  %6.2f elements/ns/SM-equivalent -> %6.2f elem/clk/SM @1.9GHz\n", name, blocks_per_sm * 8, ms,
           elems / (ms * 1e6) / 148, elems / (ms * 1e6) / 148 / 1.9);
    cudaFree(in); cudaFree(out);
}
int main() {
    for (int b : {1, 4}) {
        if (b == 1) { run<0>("V0 1 MUFU + poly5", 1); run<1>("V1 2 MUFU", 1); run<2>("V2 FFMA2 poly5", 1); run<3>("V3 poly3 probe", 1); run<4>("V4 mix 1:1", 1); run<5>("V5 mix 3:1", 1); }
        else { run<0>("V0 1 MUFU + poly5", 4); run<1>("V1 2 MUFU", 4); run<2>("V2 FFMA2 poly5", 4); run<3>("V3 poly3 probe", 4); run<4>("V4 mix 1:1", 4); run<5>("V5 mix 3:1", 4); }
    }
    return 0;
}
