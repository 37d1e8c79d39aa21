// Memory-bound kernels of the TTSZipformer forward and of the sampler: fused, coalesced,
// 16-byte vectorised, warp-shuffle reductions.  Activations are (N, L, C) fp16, channel
// contiguous; the ODE state and velocities are fp32.
#pragma once
#include "ptx.cuh"

namespace zvb {

__device__ __forceinline__ void unpack8(const uint4& w, float* v) {
    v[0] = h2_lo(w.x); v[1] = h2_hi(w.x); v[2] = h2_lo(w.y); v[3] = h2_hi(w.y);
    v[4] = h2_lo(w.z); v[5] = h2_hi(w.z); v[6] = h2_lo(w.w); v[7] = h2_hi(w.w);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]),
                      pack_h2(v[6], v[7]));
}

// ---------------------------------------------------------------------------------------
// The residual stream is fp16 (N, L, C) and is itself the A operand of the next tensor-core kernel;
// kernels that start a module on `src + time_emb` (reference: modules/zipformer.py:532-534) read the
// separate time-embedded copy written here.  All arithmetic is fp32.

// streaming 16-byte load that does not allocate in L1 (every element is read exactly once)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void load8h(const __half* p, float* v) {
    unpack8(*reinterpret_cast<const uint4*>(p), v);
}
__device__ __forceinline__ void store8h(__half* p, const float* v) {
    *reinterpret_cast<uint4*>(p) = pack8(v);
}

// BiasNorm + bypass (reference: modules/scaling.py:358-363, modules/zipformer.py:634-637,
// 803-804):  y = x * rsqrt(mean((x-b)^2)) * exp(log_scale);  out = orig + (y-orig)*scale.
// One warp per row, C <= 256*KMAX, C % 8 == 0.  Lane l owns the 8 channels at 8*(32k + l): every warp
// instruction moves one contiguous 512-byte segment, and both operands of the row are requested
// before the reduction so each lane has 2*KMAX 16-byte loads in flight.
// Outputs (each nullable): stream `out`, `out_t` = out + temb[row / rows_per_group].
template <int KMAX>
__global__ void __launch_bounds__(256)
biasnorm_bypass_kernel(const __half* __restrict__ src, const __half* __restrict__ orig,
                       __half* __restrict__ out, __half* __restrict__ out_t,
                       const float* __restrict__ temb, int rows_per_group,
                       const float* __restrict__ nbias, const float* __restrict__ log_scale,
                       const float* __restrict__ bscale, long long rows, int C) {
    pdl_wait();
    pdl_launch();
    const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    uint4 xr[KMAX], orr[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (c < C) {
            xr[k] = ld_stream_u4(src + row * C + c);
            orr[k] = ld_stream_u4(orig + row * C + c);
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (c < C) {
            float x[8];
            unpack8(xr[k], x);
            const float4 n0 = __ldg(reinterpret_cast<const float4*>(nbias + c));
            const float4 n1 = __ldg(reinterpret_cast<const float4*>(nbias + c + 4));
            const float nb[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = x[i] - nb[i];
                ss = fmaf(d, d, ss);
            }
        }
    }
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, q);
    const float scale = rsqrtf(ss / static_cast<float>(C)) * __expf(__ldg(log_scale));
    const long long grp = out_t != nullptr ? row / rows_per_group : 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (c < C) {
            float x[8], o[8], y[8];
            unpack8(xr[k], x);
            unpack8(orr[k], o);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bscale + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bscale + c + 4));
            const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = o[i] + (x[i] * scale - o[i]) * bs[i];
            if (out != nullptr) store8h(out + row * C + c, y);
            if (out_t != nullptr) {
                const float4 t0 = __ldg(reinterpret_cast<const float4*>(temb + grp * C + c));
                const float4 t1 = __ldg(reinterpret_cast<const float4*>(temb + grp * C + c + 4));
                y[0] += t0.x; y[1] += t0.y; y[2] += t0.z; y[3] += t0.w;
                y[4] += t1.x; y[5] += t1.y; y[6] += t1.z; y[7] += t1.w;
                store8h(out_t + row * C + c, y);
            }
        }
    }
}

// Stack entry: the time-embedded copy of the stream, xt = x + temb[row / L]
// (reference: modules/zipformer.py:532-534).  One thread = 8 channels of 4 rows (the four 16-byte loads
// are issued before anything is consumed; a warp covers 256 contiguous channels of a row).
__global__ void __launch_bounds__(256)
stream_prep_kernel(const __half* __restrict__ x, __half* __restrict__ xt,
                   const float* __restrict__ temb, int rows_per_group, long long rows, int C) {
    pdl_wait();
    pdl_launch();
    constexpr int RPT = 4;
    const int cv = C >> 3;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long rgroups = (rows + RPT - 1) / RPT;
    if (idx >= rgroups * cv) return;
    const long long rg = idx / cv;
    const int c = static_cast<int>(idx - rg * cv) * 8;
    const long long row0 = rg * RPT;
    uint4 w[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
        if (row0 + i < rows) w[i] = ld_stream_u4(x + (row0 + i) * C + c);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const long long row = row0 + i;
        if (row >= rows) break;
        const float* tp = temb + (row / rows_per_group) * C + c;
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(tp));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(tp + 4));
        float v[8];
        unpack8(w[i], v);
        v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
        v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
        store8h(xt + row * C + c, v);
    }
}

// SimpleDownsample (reference: modules/zipformer.py:887-913): weighted sum over groups of ds
// frames, right-padded by repeating frame L-1.  w = softmax(bias) precomputed on the host.
// One thread = 8 channels of two output frames; all 2*ds 16-byte loads are issued before the first use.
__global__ void __launch_bounds__(256)
downsample_kernel(const __half* __restrict__ src, __half* __restrict__ out, int N, int L,
                  int Ld, int ds, float w0, float w1, float w2, float w3, int C) {
    pdl_wait();
    pdl_launch();
    constexpr int RPT = 2;
    const int cv = C >> 3;
    const int Lg = (Ld + RPT - 1) / RPT;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Lg * cv) return;
    const int c = static_cast<int>(idx % cv) * 8;
    const long long rl = idx / cv;
    const int lg = static_cast<int>(rl % Lg);
    const int n = static_cast<int>(rl / Lg);
    const float w[4] = {w0, w1, w2, w3};
    uint4 in[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int l = (lg * RPT + i) * ds + k;
            l = l < L ? l : L - 1;
            if (k < ds) in[i][k] = ld_stream_u4(src + (static_cast<long long>(n) * L + l) * C + c);
        }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int ld = lg * RPT + i;
        if (ld >= Ld) break;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < ds) {
                float v[8];
                unpack8(in[i][k], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[e], w[k], acc[e]);
            }
        }
        store8h(out + (static_cast<long long>(n) * Ld + ld) * C + c, acc);
    }
}

// SimpleUpsample + truncate + out_combiner bypass (reference: modules/zipformer.py:866-870,
// 925-935):  out[n,l] = orig[n,l] + (y[n, l/ds] - orig[n,l]) * scale
// One thread = 8 channels of the ds frames that share one low-rate frame (loads first).
__global__ void __launch_bounds__(256)
upsample_combine_kernel(const __half* __restrict__ orig, const __half* __restrict__ y,
                        __half* __restrict__ out, const float* __restrict__ scale, int N, int L, int Ld,
                        int ds, int C) {
    pdl_wait();
    pdl_launch();
    const int cv = C >> 3;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ld * cv) return;
    const int c = static_cast<int>(idx % cv) * 8;
    const long long rl = idx / cv;
    const int ld = static_cast<int>(rl % Ld);
    const int n = static_cast<int>(rl / Ld);
    const uint4 yv = ld_stream_u4(y + (static_cast<long long>(n) * Ld + ld) * C + c);
    uint4 ov[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int l = ld * ds + k;
        if (k < ds && l < L) ov[k] = ld_stream_u4(orig + (static_cast<long long>(n) * L + l) * C + c);
    }
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    float v[8];
    unpack8(yv, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int l = ld * ds + k;
        if (k < ds && l < L) {
            float o[8], r[8];
            unpack8(ov[k], o);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] = o[e] + (v[e] - o[e]) * sc[e];
            store8h(out + (static_cast<long long>(n) * L + l) * C + c, r);
        }
    }
}

// Depthwise Conv1d over time (cross-correlation, zero padding K/2) + bias + SwooshR
// (reference: modules/zipformer.py:1672-1678 with scaling.py:1200-1206).  The input is the
// already GLU-gated and key-masked tensor.  The kernel is bound by the fp32 FMA pipe (K taps per
// output), so everything else is kept off the critical path: a tile = 64 channels x 128 frames
// (+ K-1 halo rows) arrives by ONE TMA box load (rows before the utterance start / after its end are
// zero-filled by the tensor map = the convolution's zero padding), tiles are double buffered so the
// next load overlaps the FMAs, and a block walks its (utterance, time-tile) list persistently with the
// 64-channel group's taps staged once.  Thread = one channel pair x 16 consecutive frames with the
// 16+K-1 input window in registers as packed fp32x2: one FFMA2 advances both channels.
constexpr int DW_TT = 128;     // frames per tile
// Outputs per thread and threads per block by tap count.  The fp32x2 window (OT + K - 1 pairs) and the OT
// accumulators must live in registers: at K = 31 that is 62 + 32 register pairs, so the block is 128 threads
// (two blocks per SM, 255 registers each) with 32 outputs per thread -- 31 FFMA2 per output pair against
// (32 + 30) / 32 window loads and conversions.  (Round 1 ran 16 outputs per thread under an 80-register cap:
// ptxas re-loaded and re-converted the window inside the tap loop, FFMA2 was 52% of the issued instructions
// and the kernel sat at 40% of the FMA pipe, 0.24 of HBM peak.)
// MODE 0 (default): the register-resident shape above for K > 15, round 1's shape (16 outputs per thread, 256 threads, three
// blocks per SM under an 80-register cap) for the short kernels; MODE 1: round 1's shape everywhere; MODE 2: 20 outputs per
// thread, three warps per scheduler (ZVB_DW_MODE).  History (round 2, one B200, K = 31 launches of the forward): with the
// activation and the stores of an output inside one `if (t < L)` block, ALL shapes measured 205 us -- ptxas could not
// interleave the MUFU -> Horner chains of different outputs, they ran back to back whatever the block shape; with the
// activation straight-line for all outputs and the stores behind one tile-uniform branch: 184 us (round-1 shape) and
// 175 us (register-resident).  Ceiling: every FMA form issues at most 32 lanes x 1 FMA per clock and scheduler
// (tools/microbench/fma_rates.cu: FFMA 1.05 clk, FFMA2 2.18 clk, HFMA2 2.0 clk per warp instruction), i.e.
// 128 FMA/clk/SM, and 31 taps + SwooshR need 40 FMA per output: 121 us at 1.55 GHz, 0.40 of HBM peak.
template <int K, int MODE> struct DwShape {
    // MODE 2 (K > 15): 20 outputs per thread, four warps, three blocks per SM at <= 168 registers -- three warps per scheduler
    static constexpr int OT = (K > 15) ? (MODE == 0 ? 32 : MODE == 2 ? 20 : 16) : 16;
    static constexpr int WARPS = (K > 15 && MODE == 2) ? 4 : DW_TT / OT;
    static constexpr int TT = OT * WARPS;                      // frames per tile
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int MINB = (MODE == 0 && K > 15) ? 2 : 3;
};
template <int K, int MODE> constexpr int dw_smem_bytes() { return 2 * (DwShape<K, MODE>::TT + K - 1) * 128 + K * 32 * 8 + 128 /*align*/ + 16; }

// SwooshR of a channel pair (see swoosh_direct): everything but the two MUFU ops runs as FFMA2 --
// 8 packed FMAs per pair; |z| is two ALU-pipe LOP3s, the negation rides on the MUFU operand.
__device__ __forceinline__ void swoosh_r_pair(f32x2 acc, float& o0, float& o1) {
    constexpr float L2E = 1.4426950408889634f;
    const f32x2 z = fma2(acc, pack2(L2E, L2E), pack2(-SWOOSH_R_C * L2E, -SWOOSH_R_C * L2E));
    float z0, z1;
    unpack2(z, z0, z1);
    const float a0 = fabsf(z0), a1 = fabsf(z1);
    float t0, t1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(-a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(-a1));
    const f32x2 t = pack2(t0, t1);
    f32x2 q = fma2(t, pack2(0.031377589387161245f, 0.031377589387161245f), pack2(-0.1341354334221127f, -0.1341354334221127f));
    q = fma2(t, q, pack2(0.2878262894239249f, 0.2878262894239249f));
    q = fma2(t, q, pack2(-0.491347927069251f, -0.491347927069251f));
    q = fma2(t, q, pack2(0.9994349844843187f, 0.9994349844843187f));
    constexpr float K = SWOOSH_R_K0 - 0.42f * SWOOSH_R_C;
    q = fma2(t, q, pack2(K, K));
    q = fma2(acc, pack2(0.42f, 0.42f), q);
    q = fma2(pack2(a0, a1), pack2(0.5f / L2E, 0.5f / L2E), q);
    unpack2(q, o0, o1);
}

// tma_x: x viewed as (C, L, N) fp16, box = 64 channels x (DW_TT + K - 1) frames, no swizzle.
// grid = (blocks per channel group, channel groups); tiles = N * ceil(L / DW_TT) per channel group.
// ACT = 1: SwooshR (ConvolutionModule); ACT = 0: bias only (the vocoder's ConvNeXt blocks, vocoder.cuh).
template <int K, int ACT, int MODE>
__global__ void __launch_bounds__(DwShape<K, MODE>::THREADS, DwShape<K, MODE>::MINB)
dwconv_kernel(const __grid_constant__ CUtensorMap tma_x, __half* __restrict__ out,
              const float* __restrict__ wt /*[K][C]*/, const float* __restrict__ bias, int L,
              int C, int N) {
    constexpr int HALF = K / 2, TT = DwShape<K, MODE>::TT, WIN = TT + K - 1, OT = DwShape<K, MODE>::OT, NW = OT + K - 1,
                  THREADS = DwShape<K, MODE>::THREADS;
    extern __shared__ uint8_t dw_smem_raw[];
    const uint32_t raw = smem_u32(dw_smem_raw);
    uint8_t* smem = dw_smem_raw + (((raw + 127u) & ~127u) - raw);
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem);                       // [2][WIN][32] fp16 pairs
    float2* wsm = reinterpret_cast<float2*>(smem + 2 * WIN * 128);           // [K][32]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * WIN * 128 + K * 32 * 8);   // [2]
    const int c0 = blockIdx.y * 64;
    const int cp = threadIdx.x & 31;
    const int tg = threadIdx.x >> 5;
    const int n_tt = (L + TT - 1) / TT;
    const int total = N * n_tt;
    for (int idx = threadIdx.x; idx < K * 32; idx += THREADS) {
        const int k = idx >> 5, q = idx & 31;
        const int c = c0 + 2 * q;
        wsm[idx] = c < C ? make_float2(__ldg(wt + k * C + c), __ldg(wt + k * C + c + 1)) : make_float2(0.f, 0.f);
    }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_x);
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();                 // the taps above are constants; the input tile is the predecessor's output
    pdl_launch();
    int tl = blockIdx.x;
    if (threadIdx.x == 0 && tl < total) {
        mbar_arrive_expect_tx(&bar[0], WIN * 128);
        tma_load_3d(tile, &tma_x, &bar[0], c0, (tl % n_tt) * TT - HALF, tl / n_tt);
    }
    const bool ch_ok = c0 + 2 * cp < C;
    f32x2 b2 = pack2(0.f, 0.f);
    if (ch_ok) b2 = pack2(__ldg(bias + c0 + 2 * cp), __ldg(bias + c0 + 2 * cp + 1));
    for (uint32_t it = 0; tl < total; tl += gridDim.x, ++it) {
        const uint32_t buf = it & 1u;
        const int nxt = tl + gridDim.x;
        if (threadIdx.x == 0 && nxt < total) {        // buffer buf^1 was drained before the last __syncthreads
            mbar_arrive_expect_tx(&bar[buf ^ 1u], WIN * 128);
            tma_load_3d(tile + (buf ^ 1u) * (WIN * 32), &tma_x, &bar[buf ^ 1u], c0, (nxt % n_tt) * TT - HALF,
                        nxt / n_tt);
        }
        mbar_wait(&bar[buf], (it >> 1) & 1u);
        const uint32_t* tb = tile + buf * (WIN * 32) + (tg * OT) * 32 + cp;
        f32x2 win[NW];                          // the input window of the channel pair, fp32x2 packed
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            const uint32_t u = tb[q * 32];
            win[q] = pack2(h2_lo(u), h2_hi(u));
        }
        f32x2 acc[OT];
#pragma unroll
        for (int o = 0; o < OT; ++o) acc[o] = b2;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const f32x2 w = *reinterpret_cast<const f32x2*>(&wsm[k * 32 + cp]);
#pragma unroll
            for (int o = 0; o < OT; ++o) acc[o] = fma2(w, win[o + k], acc[o]);     // one FFMA2 = both channels
        }
        const int n = tl / n_tt;
        const int t0 = (tl - n * n_tt) * TT + tg * OT;
        __half* on = out + (static_cast<long long>(n) * L + t0) * C + c0 + 2 * cp;
        if (ch_ok) {
            // activation for all outputs first, straight-line (a store inside its own `if (t < L)` block per output kept
            // ptxas from interleaving the MUFU -> Horner chains of different outputs: they ran back to back), then the
            // stores: unpredicated for tiles inside the utterance, per-row bounds only in its last tile
            uint32_t res[OT];
#pragma unroll
            for (int o = 0; o < OT; ++o) {
                float a0, a1;
                if (ACT) swoosh_r_pair(acc[o], a0, a1);
                else unpack2(acc[o], a0, a1);
                res[o] = pack_h2(a0, a1);
            }
            const long long rowb = static_cast<long long>(C) * 2;          // bytes per frame
            uint8_t* ob = reinterpret_cast<uint8_t*>(on);
            if (t0 + OT <= L) {
#pragma unroll
                for (int o = 0; o < OT; ++o) *reinterpret_cast<uint32_t*>(ob + o * rowb) = res[o];
            } else {
#pragma unroll
                for (int o = 0; o < OT; ++o)
                    if (t0 + o < L) *reinterpret_cast<uint32_t*>(ob + o * rowb) = res[o];
            }
        }
        __syncthreads();                        // the window reads of this buffer are done: it may be refilled
    }
}

// ---------------------------------------------------------------------------------------
// Time / guidance embedding (reference: modules/zipformer.py:47-69): out[n] = [cos(t f) | sin(t f)]
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int N, int dim) {
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
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
// xin[n] = [x | text | speech] as fp16, zero-padded to `ldx` columns.  With cfg != 0 the batch
// is doubled as [uncond ; cond]: uncond rows get text = 0 and, when drop_speech != 0
// (t > 0.5), speech = 0.
__global__ void __launch_bounds__(256)
assemble_input_kernel(const float* __restrict__ x, const float* __restrict__ text,
                      const float* __restrict__ speech, __half* __restrict__ xin, int B, int T,
                      int F, int Ft, int ldx, int cfg, int drop_speech) {
    pdl_wait();
    pdl_launch();
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
    xin[idx] = f2h(v);
}

// fp32 (rows, C) -> fp16 (rows, ldx) zero padded (seam-1 entry: caller passes the concatenated x)
__global__ void __launch_bounds__(256)
cast_pad_kernel(const float* __restrict__ x, __half* __restrict__ out, long long rows, int C, int ldx) {
    pdl_wait();
    pdl_launch();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= rows * ldx) return;
    const int c = static_cast<int>(idx % ldx);
    const long long r = idx / ldx;
    out[idx] = f2h(c < C ? x[r * C + c] : 0.f);
}

// CFG blend + Euler update (reference: modules/solver.py:100-110, 239):
//   v = (1+g)*v_cond - g*v_uncond  (rows [0,B) of `v` are uncond, [B,2B) cond);  x += dt*v
// g = gscale * guidance[b] (guidance per utterance; gscale = 2 when t <= 0.5).  cfg == 0: x += dt*v.
// `vout` (nullable) receives the blended velocity for parity checks.
__global__ void __launch_bounds__(256)
cfg_euler_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ guidance,
                 float gscale, const float* __restrict__ ts, int step, float* __restrict__ vout, int B,
                 long long per_utt, int cfg) {
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Ld) return;
    const int n = idx / Ld, l = idx - n * Ld;
    out[idx] = mask[static_cast<long long>(n) * T + l * ds];
}

// Excluded-key bit words for the attention kernel: bit (j % 32) of word j / 32 is set when key j of the
// utterance is padded or j >= L; `words` per utterance covers whole 128-key tiles.
__global__ void mask_words_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ out, int N, int L,
                                  int words) {
    pdl_wait();
    pdl_launch();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * words) return;
    const int n = idx / words, w = idx - n * words;
    uint32_t bits = 0u;
    for (int b = 0; b < 32; ++b) {
        const int j = w * 32 + b;
        const bool ex = j < L ? mask[static_cast<long long>(n) * L + j] != 0 : true;
        bits |= (ex ? 1u : 0u) << b;
    }
    out[idx] = bits;
}

// ---------------------------------------------------------------------------------------
// Per-forward prologue in three launches instead of twenty-six (single utterances pay ~8.5 us per launch):
//   * prologue_masks_kernel: the strided key-padding masks mask[:, ::ds] (bytes) AND the attention kernel's excluded-key
//     bit words of every resolution in one launch;
//   * multi_small_linear_kernel: the per-stack time-embedding projections (reference: zipformer.py:676-680) in one launch;
//   * rowbias_all_kernel: for EVERY layer the row bias W1 * temb of the merged feed_forward1 / attention projection.
struct MaskJob {
    uint8_t* strided;      // [N][Ld] bytes (null for ds == 1: the input mask itself)
    uint32_t* words;       // [N][nwords]
    int ds, Ld, nwords;
};
struct MaskJobs { MaskJob j[3]; int n; };

__global__ void prologue_masks_kernel(const uint8_t* __restrict__ mask, const MaskJobs jobs, int N, int T) {
    pdl_wait();
    pdl_launch();
    if (static_cast<int>(blockIdx.y) >= jobs.n) return;
    const MaskJob jb = jobs.j[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * jb.nwords) return;
    const int n = idx / jb.nwords, w = idx - n * jb.nwords;
    const uint8_t* mrow = mask + static_cast<long long>(n) * T;
    uint32_t bits = 0u;
    for (int b = 0; b < 32; ++b) {
        const int j = w * 32 + b;
        bool ex = true;
        if (j < jb.Ld) {
            const uint8_t v = mrow[j * jb.ds];
            ex = v != 0;
            if (jb.strided != nullptr) jb.strided[static_cast<long long>(n) * jb.Ld + j] = v;
        }
        bits |= (ex ? 1u : 0u) << b;
    }
    jb.words[idx] = bits;
}

struct SmallJob { const float* W; const float* b; float* out; };
struct SmallJobs { SmallJob j[8]; int n; };

// out_s[n,o] = b_s[o] + sum_k W_s[o,k] * swoosh_r(in[n,k]) for every job s (one warp per output element)
__global__ void __launch_bounds__(256)
multi_small_linear_kernel(const float* __restrict__ in, const SmallJobs jobs, int N, int K, int O) {
    pdl_wait();
    pdl_launch();
    if (static_cast<int>(blockIdx.y) >= jobs.n) return;
    const SmallJob jb = jobs.j[blockIdx.y];
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= N * O) return;
    const int n = gw / O, o = gw - n * O;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32)
        acc = fmaf(__ldg(jb.W + static_cast<long long>(o) * K + k), swoosh_r(in[n * K + k]), acc);
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
    if (lane == 0) jb.out[n * O + o] = acc + (jb.b != nullptr ? jb.b[o] : 0.f);
}

struct RowBiasJob { const __half* W; const float* temb; float* out; };     // W [rows][kp] fp16, temb [N][D], out [N][ld]
constexpr int RB_MAX_JOBS = 64;
struct RowBiasJobs { RowBiasJob j[RB_MAX_JOBS]; int n; };
constexpr int RB_NCHUNK = 16;
constexpr int RB_COLS = 128;

// out_l[n][c] = sum_k W_l[c][k] * temb_l[n][k], c < cols.  One block per (128 columns, layer, chunk of 16 utterances): a
// thread owns ONE column (its 16 sums need no reduction; its weight row streams through L1 16 bytes at a time), the chunk's
// time embeddings sit in shared memory and are read as warp-wide broadcasts.  D <= 512, D % 8 == 0.
__global__ void __launch_bounds__(RB_COLS)
rowbias_all_kernel(const RowBiasJobs jobs, int N, int D, int kp, int cols, int ld) {
    __shared__ __align__(16) float te[RB_NCHUNK][512];
    pdl_wait();
    pdl_launch();
    const RowBiasJob jb = jobs.j[blockIdx.y];
    const int n0 = blockIdx.z * RB_NCHUNK;
    const int nn = N - n0 < RB_NCHUNK ? N - n0 : RB_NCHUNK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < nn; n += RB_COLS / 32)
        for (int k = lane * 4; k < D; k += 128)
            *reinterpret_cast<float4*>(&te[n][k]) = *reinterpret_cast<const float4*>(jb.temb + static_cast<long long>(n0 + n) * D + k);
    __syncthreads();
    const int c = blockIdx.x * RB_COLS + threadIdx.x;
    if (c >= cols) return;
    float acc[RB_NCHUNK];
#pragma unroll
    for (int n = 0; n < RB_NCHUNK; ++n) acc[n] = 0.f;
    const __half* wrow = jb.W + static_cast<long long>(c) * kp;
#pragma unroll 2
    for (int k0 = 0; k0 < D; k0 += 8) {
        float w[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(wrow + k0)), w);
#pragma unroll
        for (int n = 0; n < RB_NCHUNK; ++n) {
            if (n < nn) {
                const float4 t0 = *reinterpret_cast<const float4*>(&te[n][k0]);
                const float4 t1 = *reinterpret_cast<const float4*>(&te[n][k0 + 4]);
                acc[n] = fmaf(w[0], t0.x, acc[n]); acc[n] = fmaf(w[1], t0.y, acc[n]);
                acc[n] = fmaf(w[2], t0.z, acc[n]); acc[n] = fmaf(w[3], t0.w, acc[n]);
                acc[n] = fmaf(w[4], t1.x, acc[n]); acc[n] = fmaf(w[5], t1.y, acc[n]);
                acc[n] = fmaf(w[6], t1.z, acc[n]); acc[n] = fmaf(w[7], t1.w, acc[n]);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < RB_NCHUNK; ++n)
        if (n < nn) jb.out[static_cast<long long>(n0 + n) * ld + c] = acc[n];
}

// The same for N <= 16 utterances (single samples: latency, not throughput): one WARP per (layer, column), the lanes stride
// over k with coalesced 16-byte weight loads -- 18 k warps keep every SM's load pipes full where one thread per column would
// walk its row 16 bytes at a time (32 us against ~10 at N = 2).
__global__ void __launch_bounds__(256)
rowbias_small_kernel(const RowBiasJobs jobs, int N, int D, int kp, int cols, int ld) {
    __shared__ __align__(16) float te[RB_NCHUNK][512];
    pdl_wait();
    pdl_launch();
    const RowBiasJob jb = jobs.j[blockIdx.y];
    const int nn = N < RB_NCHUNK ? N : RB_NCHUNK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < nn; n += 8)
        for (int k = lane * 4; k < D; k += 128)
            *reinterpret_cast<float4*>(&te[n][k]) = *reinterpret_cast<const float4*>(jb.temb + static_cast<long long>(n) * D + k);
    __syncthreads();
    const int c = blockIdx.x * 8 + warp;
    if (c >= cols) return;
    float acc[RB_NCHUNK];
#pragma unroll
    for (int n = 0; n < RB_NCHUNK; ++n) acc[n] = 0.f;
    for (int k0 = lane * 8; k0 < D; k0 += 256) {
        float w[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(jb.W + static_cast<long long>(c) * kp + k0)), w);
#pragma unroll
        for (int n = 0; n < RB_NCHUNK; ++n) {
            if (n < nn) {
                const float4 t0 = *reinterpret_cast<const float4*>(&te[n][k0]);
                const float4 t1 = *reinterpret_cast<const float4*>(&te[n][k0 + 4]);
                acc[n] = fmaf(w[0], t0.x, acc[n]); acc[n] = fmaf(w[1], t0.y, acc[n]);
                acc[n] = fmaf(w[2], t0.z, acc[n]); acc[n] = fmaf(w[3], t0.w, acc[n]);
                acc[n] = fmaf(w[4], t1.x, acc[n]); acc[n] = fmaf(w[5], t1.y, acc[n]);
                acc[n] = fmaf(w[6], t1.z, acc[n]); acc[n] = fmaf(w[7], t1.w, acc[n]);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < RB_NCHUNK; ++n) {
        if (n < nn) {                                  // warp-uniform
            float a = acc[n];
#pragma unroll
            for (int q = 16; q > 0; q >>= 1) a += __shfl_xor_sync(0xffffffffu, a, q);
            if (lane == 0) jb.out[static_cast<long long>(n) * ld + c] = a;
        }
    }
}

// Debug / parity aid: counts fp16 elements whose magnitude is the largest finite value or above (what the
// saturating conversions `cvt.rn.satfinite.f16x2.f32` of the path produce on overflow, plus inf / NaN).
__global__ void __launch_bounds__(256)
count_saturated_kernel(const __half* __restrict__ p, long long n, unsigned long long* __restrict__ counter) {
    pdl_wait();
    pdl_launch();
    const long long vec = n >> 3;
    unsigned int c = 0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < vec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 w = ld_stream_u4(p + i * 8);
        const uint32_t q[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c += ((q[k] & 0x7FFFu) >= 0x7BFFu) ? 1u : 0u;
            c += (((q[k] >> 16) & 0x7FFFu) >= 0x7BFFu) ? 1u : 0u;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const unsigned short b = reinterpret_cast<const unsigned short*>(p)[vec * 8 + threadIdx.x];
        c += ((b & 0x7FFFu) >= 0x7BFFu) ? 1u : 0u;
    }
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) c += __shfl_xor_sync(0xffffffffu, c, q);
    if ((threadIdx.x & 31) == 0 && c != 0u) atomicAdd(counter, static_cast<unsigned long long>(c));
}

}  // namespace zvb
