// Issue rate of the FMA forms a depthwise convolution / epilogue can be built from, per SM sub-partition:
// cycles per warp instruction with 1, 2 and 4 warps per scheduler, 16 independent accumulators per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rates fma_rates.cu && ./fma_rates
// Forms: FFMA 3-register, FFMA with a constant-bank multiplicand (warp-uniform weight), FFMA with an immediate
// addend, FFMA2 (fp32x2) 3-register, FFMA2 with a constant-bank operand, HFMA2 3-register, HFMA2 constant-bank.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

__constant__ float cw[64];
__constant__ uint32_t chw[64];
__constant__ unsigned long long cw2[64];

constexpr int ACC = 16, ITERS = 256;

template <int MODE>
__global__ void bench(float* out, long long* cyc, const float* in) {
    float a[ACC], x[ACC];
    f32x2 a2[ACC], x2[ACC];
    uint32_t ah[ACC], xh[ACC];
    const float w = in[threadIdx.x & 31];
    const f32x2 w2 = pack2(w, w + 1.0f);
    const uint32_t wh = __float_as_uint(w);
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        a[i] = in[i] + threadIdx.x; x[i] = in[i + 16] * 0.001f;
        a2[i] = pack2(a[i], a[i] + 1.f); x2[i] = pack2(x[i], x[i] * 2.f);
        ah[i] = __float_as_uint(a[i]); xh[i] = __float_as_uint(x[i]);
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int i = 0; i < ACC; ++i) {
                if (MODE == 0) a[i] = ffma(x[(i + k) & 15], w, a[i]);                 // 3 registers
                if (MODE == 1) a[i] = ffma(x[(i + k) & 15], cw[k], a[i]);             // constant-bank multiplicand
                if (MODE == 2) a[i] = ffma(a[i], x[(i + k) & 15], 0.2878262894239249f);  // immediate addend
                if (MODE == 3) a2[i] = fma2(x2[(i + k) & 15], w2, a2[i]);             // fp32x2, 3 register pairs
                if (MODE == 4) { f32x2 c; asm volatile("ld.const.b64 %0, [%1];" : "=l"(c) : "l"(&cw2[k])); a2[i] = fma2(x2[(i + k) & 15], c, a2[i]); }
                if (MODE == 5) ah[i] = hfma2(xh[(i + k) & 15], wh, ah[i]);            // half2, 3 registers
                if (MODE == 6) ah[i] = hfma2(xh[(i + k) & 15], chw[k], ah[i]);        // half2, constant-bank
                if (MODE == 7) a[i] = ffma(x[(i + k) & 15], x[(i + k + 1) & 15], a[i]);  // 3 distinct vector registers, no reuse
                if (MODE == 8) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i]));            // MUFU.EX2
                if (MODE == 9) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(ah[i]) : "r"(ah[i]));            // 2 x MUFU.EX2.F16 ?
                if (MODE == 10) { asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(ah[i]) : "r"(ah[i]));          // MUFU pair + 3 HFMA2
                                  xh[i] = hfma2(xh[i], wh, ah[i]); xh[(i + 5) & 15] = hfma2(xh[(i + 5) & 15], wh, xh[i]);
                                  xh[(i + 9) & 15] = hfma2(xh[(i + 9) & 15], wh, xh[i]); }
                if (MODE == 11) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i]));          // MUFU + 3 FFMA
                                  x[i] = ffma(x[i], w, a[i]); x[(i + 5) & 15] = ffma(x[(i + 5) & 15], w, x[i]);
                                  x[(i + 9) & 15] = ffma(x[(i + 9) & 15], w, x[i]); }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        float p, q; unpack2(a2[i], p, q);
        s += a[i] + p + q + __uint_as_float(ah[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, float* out, long long* cyc, const float* in) {
    for (int warps : {4, 8, 16}) {           // per SM: 1, 2, 4 per scheduler
        bench<MODE><<<1, warps * 32>>>(out, cyc, in);
        cudaDeviceSynchronize();
        bench<MODE><<<1, warps * 32>>>(out, cyc, in);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / (ITERS * 4.0 * ACC) / (warps / 4.0);
        printf("%-44s warps/sched %d: %.2f clk per warp instruction\n", name, warps / 4, per);
    }
}

int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 1 << 16); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 64);
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.01f * i;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cw, h, sizeof h);
    cudaMemcpyToSymbol(chw, h, sizeof h);
    unsigned long long h2[64]; for (int i = 0; i < 64; ++i) { float p[2] = {h[i], h[i]}; memcpy(&h2[i], p, 8); }
    cudaMemcpyToSymbol(cw2, h2, sizeof h2);
    run<0>("FFMA  x, w(reg), acc", out, cyc, in);
    run<7>("FFMA  x, y, acc (3 distinct)", out, cyc, in);
    run<1>("FFMA  x, c[bank], acc", out, cyc, in);
    run<2>("FFMA  acc, x, imm", out, cyc, in);
    run<3>("FFMA2 x2, w2(reg), acc2", out, cyc, in);
    run<4>("FFMA2 x2, c[bank].64, acc2", out, cyc, in);
    run<5>("HFMA2 xh, wh(reg), acch", out, cyc, in);
    run<6>("HFMA2 xh, c[bank], acch", out, cyc, in);
    run<8>("MUFU.EX2 f32 (per ex2 instruction)", out, cyc, in);
    run<9>("ex2.approx.f16x2 (per PTX instruction)", out, cyc, in);
    run<10>("ex2.f16x2 + 3 HFMA2 (per group of 4)", out, cyc, in);
    run<11>("MUFU.EX2 f32 + 3 FFMA (per group of 4)", out, cyc, in);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
