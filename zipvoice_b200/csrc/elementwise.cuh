// Memory-bound kernels of the TTSZipformer forward and of the sampler: fused, coalesced,
// 16-byte vectorised, warp-shuffle reductions.  Activations are (N, L, C) bf16, channel
// contiguous; the ODE state and velocities are fp32.
#pragma once
#include "ptx.cuh"

namespace zvb {

__device__ __forceinline__ void unpack8(const uint4& w, float* v) {
    v[0] = bf16_lo(w.x); v[1] = bf16_hi(w.x); v[2] = bf16_lo(w.y); v[3] = bf16_hi(w.y);
    v[4] = bf16_lo(w.z); v[5] = bf16_hi(w.z); v[6] = bf16_lo(w.w); v[7] = bf16_hi(w.w);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                      pack_bf16(v[6], v[7]));
}

// ---------------------------------------------------------------------------------------
// The residual stream is fp32 (N, L, C); kernels that end a module also emit the bf16 shadow copy
// the next tensor-core kernel consumes (and the copy with the time embedding added, reference:
// modules/zipformer.py:532-534).
__device__ __forceinline__ void load8f(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8f(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// BiasNorm + bypass (reference: modules/scaling.py:358-363, modules/zipformer.py:634-637,
// 803-804):  y = x * rsqrt(mean((x-b)^2)) * exp(log_scale);  out = orig + (y-orig)*scale.
// One warp per row, C <= 1024, C % 8 == 0.  Outputs (each nullable): fp32 stream `out`, bf16
// shadow `out_b`, bf16 `out_t` = out + temb[row / rows_per_group].
__global__ void __launch_bounds__(256)
biasnorm_bypass_kernel(const float* __restrict__ src, const float* __restrict__ orig,
                       float* __restrict__ out, __nv_bfloat16* __restrict__ out_b,
                       __nv_bfloat16* __restrict__ out_t, const float* __restrict__ temb, int rows_per_group,
                       const float* __restrict__ nbias, const float* __restrict__ log_scale,
                       const float* __restrict__ bscale, long long rows, int C) {
    const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float x[4][8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (c < C) {
            load8f(src + row * C + c, x[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = x[k][i] - __ldg(nbias + c + i);
                ss = fmaf(d, d, ss);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = rsqrtf(ss / static_cast<float>(C)) * __expf(__ldg(log_scale));
    const long long grp = out_t != nullptr ? row / rows_per_group : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (c < C) {
            float o[8], y[8];
            load8f(orig + row * C + c, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = o[i] + (x[k][i] * scale - o[i]) * __ldg(bscale + c + i);
            if (out != nullptr) store8f(out + row * C + c, y);
            if (out_b != nullptr) *reinterpret_cast<uint4*>(out_b + row * C + c) = pack8(y);
            if (out_t != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] += __ldg(temb + grp * C + c + i);
                *reinterpret_cast<uint4*>(out_t + row * C + c) = pack8(y);
            }
        }
    }
}

// Stack entry: bf16 shadow of the fp32 stream and its time-embedded copy (each nullable):
// xb = bf16(x), xt = bf16(x + temb[row / L])   (reference: modules/zipformer.py:532-534)
__global__ void __launch_bounds__(256)
stream_prep_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, __nv_bfloat16* __restrict__ xt,
                   const float* __restrict__ temb, int rows_per_group, long long rows, int C) {
    const int cv = C >> 3;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= rows * cv) return;
    const long long row = idx / cv;
    const int c = static_cast<int>(idx - row * cv) * 8;
    float v[8];
    load8f(x + row * C + c, v);
    if (xb != nullptr) *reinterpret_cast<uint4*>(xb + row * C + c) = pack8(v);
    if (xt != nullptr) {
        const long long grp = row / rows_per_group;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += __ldg(temb + grp * C + c + i);
        *reinterpret_cast<uint4*>(xt + row * C + c) = pack8(v);
    }
}

// SimpleDownsample (reference: modules/zipformer.py:887-913): weighted sum over groups of ds
// frames, right-padded by repeating frame L-1.  w = softmax(bias) precomputed on the host.
__global__ void __launch_bounds__(256)
downsample_kernel(const float* __restrict__ src, float* __restrict__ out, int N, int L,
                  int Ld, int ds, float w0, float w1, float w2, float w3, int C) {
    const int cv = C >> 3;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ld * cv) return;
    const int c = static_cast<int>(idx % cv) * 8;
    const long long rl = idx / cv;
    const int ld = static_cast<int>(rl % Ld);
    const int n = static_cast<int>(rl / Ld);
    const float w[4] = {w0, w1, w2, w3};
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ds; ++k) {
        int l = ld * ds + k;
        l = l < L ? l : L - 1;
        float v[8];
        load8f(src + (static_cast<long long>(n) * L + l) * C + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], w[k], acc[i]);
    }
    store8f(out + (static_cast<long long>(n) * Ld + ld) * C + c, acc);
}

// SimpleUpsample + truncate + out_combiner bypass (reference: modules/zipformer.py:866-870,
// 925-935):  out[n,l] = orig[n,l] + (y[n, l/ds] - orig[n,l]) * scale   (+ optional bf16 shadow)
__global__ void __launch_bounds__(256)
upsample_combine_kernel(const float* __restrict__ orig, const float* __restrict__ y,
                        float* __restrict__ out, __nv_bfloat16* __restrict__ out_b,
                        const float* __restrict__ scale, int N, int L, int Ld, int ds, int C) {
    const int cv = C >> 3;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * L * cv) return;
    const int c = static_cast<int>(idx % cv) * 8;
    const long long rl = idx / cv;
    const int l = static_cast<int>(rl % L);
    const int n = static_cast<int>(rl / L);
    float o[8], v[8];
    load8f(orig + rl * C + c, o);
    load8f(y + (static_cast<long long>(n) * Ld + l / ds) * C + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = o[i] + (v[i] - o[i]) * __ldg(scale + c + i);
    store8f(out + rl * C + c, v);
    if (out_b != nullptr) *reinterpret_cast<uint4*>(out_b + rl * C + c) = pack8(v);
}

// Depthwise Conv1d over time (cross-correlation, zero padding K/2) + bias + SwooshR
// (reference: modules/zipformer.py:1672-1678 with scaling.py:1200-1206).  The input is the
// already GLU-gated and key-masked tensor.  Block = 64 channels x 128 frames staged in shared
// memory; thread = one channel pair x 16 consecutive frames with the 16+K-1 input window held
// in registers as packed fp32x2, so every staged input is read from shared memory once and one
// FFMA2 advances both channels.
constexpr int DW_TT = 128;     // frames per block
constexpr int DW_OT = 16;      // outputs per thread
template <int K>
__global__ void __launch_bounds__(256)
dwconv_swooshr_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                      const float* __restrict__ wt /*[K][C]*/, const float* __restrict__ bias, int L,
                      int C) {
    constexpr int HALF = K / 2, WIN = DW_TT + K - 1, NW = DW_OT + K - 1;
    __shared__ uint32_t tile[WIN][32];
    __shared__ float2 wsm[K][32];
    const int c0 = blockIdx.x * 64;
    const int t0 = blockIdx.y * DW_TT;
    const int n = blockIdx.z;
    const int cp = threadIdx.x & 31;
    const int tg = threadIdx.x >> 5;
    const int cvalid = C - c0;            // channels valid in this block (multiple of 8)
    const __nv_bfloat16* xn = x + static_cast<long long>(n) * L * C;
    for (int idx = threadIdx.x; idx < WIN * 8; idx += 256) {
        const int rr = idx >> 3, q = idx & 7;                 // 8 x 16B per 64-channel row
        const int t = t0 - HALF + rr;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < L && q * 8 < cvalid)
            v = *reinterpret_cast<const uint4*>(xn + static_cast<long long>(t) * C + c0 + q * 8);
        *reinterpret_cast<uint4*>(&tile[rr][q * 4]) = v;
    }
    for (int idx = threadIdx.x; idx < K * 32; idx += 256) {
        const int k = idx >> 5, q = idx & 31;
        const int c = c0 + 2 * q;
        wsm[k][q] = c < C ? make_float2(__ldg(wt + k * C + c), __ldg(wt + k * C + c + 1)) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (2 * cp >= cvalid) return;
    f32x2 win[NW];                          // the input window of the channel pair, fp32x2 packed
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        const uint32_t u = tile[tg * DW_OT + q][cp];
        win[q] = pack2(bf16_lo(u), bf16_hi(u));
    }
    const f32x2 b2 = pack2(__ldg(bias + c0 + 2 * cp), __ldg(bias + c0 + 2 * cp + 1));
    f32x2 acc[DW_OT];
#pragma unroll
    for (int o = 0; o < DW_OT; ++o) acc[o] = b2;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const f32x2 w = *reinterpret_cast<const f32x2*>(&wsm[k][cp]);
#pragma unroll
        for (int o = 0; o < DW_OT; ++o) acc[o] = fma2(w, win[o + k], acc[o]);     // one FFMA2 = both channels
    }
    float a0[DW_OT], a1[DW_OT];
#pragma unroll
    for (int o = 0; o < DW_OT; ++o) unpack2(acc[o], a0[o], a1[o]);
    __nv_bfloat16* on = out + static_cast<long long>(n) * L * C;
#pragma unroll
    for (int o = 0; o < DW_OT; ++o) {
        const int t = t0 + tg * DW_OT + o;
        if (t < L)
            *reinterpret_cast<uint32_t*>(on + static_cast<long long>(t) * C + c0 + 2 * cp) =
                pack_bf16(swoosh_r(a0[o]), swoosh_r(a1[o]));
    }
}

// ---------------------------------------------------------------------------------------
// Time / guidance embedding (reference: modules/zipformer.py:47-69): out[n] = [cos(t f) | sin(t f)]
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int N, int dim) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim >> 1;
    if (idx >= N * half) return;
    const int n = idx / half, i = idx - n * half;
    const float f = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
    const float a = t[n] * f;
    out[n * dim + i] = cosf(a);
    out[n * dim + half + i] = sinf(a);
}

// Small fp32 linear for the (N, <=512) time-embedding MLPs (reference: modules/zipformer.py:
// 224-228, 233-238, 676-680): out[n,o] = (addend[n,o]) + bias[o] + sum_k W[o,k] * act_in(in[n,k]);
// then act_out.  One warp per output element.
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                    const float* __restrict__ addend, float* __restrict__ out, int N, int K, int O,
                    int act_in, int act_out) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= N * O) return;
    const int n = gw / O, o = gw - n * O;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
        float v = in[n * K + k];
        if (act_in == ACT_SWOOSH_R_) v = swoosh_r(v);
        acc = fmaf(__ldg(W + static_cast<long long>(o) * K + k), v, acc);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
        if (bias != nullptr) acc += bias[o];
        if (addend != nullptr) acc += addend[n * O + o];
        if (act_out == ACT_SWOOSH_R_) acc = swoosh_r(acc);
        out[n * O + o] = acc;
    }
}

// ---------------------------------------------------------------------------------------
// Decoder input assembly (reference: models/zipvoice.py:163 and modules/solver.py:83-98):
// xin[n] = [x | text | speech] as bf16, zero-padded to `ldx` columns.  With cfg != 0 the batch
// is doubled as [uncond ; cond]: uncond rows get text = 0 and, when drop_speech != 0
// (t > 0.5), speech = 0.
__global__ void __launch_bounds__(256)
assemble_input_kernel(const float* __restrict__ x, const float* __restrict__ text,
                      const float* __restrict__ speech, __nv_bfloat16* __restrict__ xin, int B, int T,
                      int F, int Ft, int ldx, int cfg, int drop_speech) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int N = cfg ? 2 * B : B;
    if (idx >= static_cast<long long>(N) * T * ldx) return;
    const int c = static_cast<int>(idx % ldx);
    const long long rl = idx / ldx;
    const int n = static_cast<int>(rl / T);
    const int b = cfg ? (n >= B ? n - B : n) : n;
    const bool uncond = cfg && n < B;
    const long long r = static_cast<long long>(b) * T + (rl - static_cast<long long>(n) * T);
    float v = 0.f;
    if (c < F) v = x[r * F + c];
    else if (c < F + Ft) v = uncond ? 0.f : text[r * Ft + (c - F)];
    else if (c < 2 * F + Ft) v = (uncond && drop_speech) ? 0.f : speech[r * F + (c - F - Ft)];
    xin[idx] = __float2bfloat16(v);
}

// fp32 (rows, C) -> bf16 (rows, ldx) zero padded (seam-1 entry: caller passes the concatenated x)
__global__ void __launch_bounds__(256)
cast_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int C, int ldx) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= rows * ldx) return;
    const int c = static_cast<int>(idx % ldx);
    const long long r = idx / ldx;
    out[idx] = __float2bfloat16(c < C ? x[r * C + c] : 0.f);
}

// CFG blend + Euler update (reference: modules/solver.py:100-110, 239):
//   v = (1+g)*v_cond - g*v_uncond  (rows [0,B) of `v` are uncond, [B,2B) cond);  x += dt*v
// g = gscale * guidance[b] (guidance per utterance; gscale = 2 when t <= 0.5).  cfg == 0: x += dt*v.
// `vout` (nullable) receives the blended velocity for parity checks.
__global__ void __launch_bounds__(256)
cfg_euler_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ guidance,
                 float gscale, const float* __restrict__ ts, int step, float* __restrict__ vout, int B,
                 long long per_utt, int cfg) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(B) * per_utt) return;
    const float dt = ts[step + 1] - ts[step];
    float vel;
    if (cfg) {
        const int b = static_cast<int>(idx / per_utt);
        const float g = gscale * guidance[b];
        vel = (1.0f + g) * v[static_cast<long long>(B) * per_utt + idx] - g * v[idx];
    } else {
        vel = v[idx];
    }
    if (vout != nullptr) vout[idx] = vel;
    x[idx] = x[idx] + vel * dt;
}

// mask[n, ::ds] (reference: modules/zipformer.py:857-858)
__global__ void stride_mask_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ out, int N, int T,
                                   int Ld, int ds) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Ld) return;
    const int n = idx / Ld, l = idx - n * Ld;
    out[idx] = mask[static_cast<long long>(n) * T + l * ds];
}

}  // namespace zvb
