// Kernels either side of the sampler:
//   * prompt log-mel (reference: zipvoice/utils/feature.py:47-116 VocosFbank = torchaudio MelSpectrogram(n_fft 1024,
//     hop 256, 100 mels, center=True, power=1) -> clamp(1e-7).log(); frame count rule of lhotse compute_num_frames);
//   * the Vocos vocoder's non-GEMM pieces (reference call sites: zipvoice/bin/infer_zipvoice.py:301-312, 409, 594;
//     the model is the external `vocos` 0.1.0 package: VocosBackbone (ConvNeXt blocks) + ISTFTHead): the 7-frame
//     window operand of the embedding convolution, LayerNorm, and the inverse STFT (exp / clip / cos / sin ->
//     inverse real FFT -> window -> overlap-add / window envelope, torch.istft(center=True) semantics).
// The transforms are exact fp32 radix-2 Stockham FFTs in shared memory (1024 points, one block per frame, looping):
// a log-mel needs relative accuracy per BIN (quiet high-frequency bins sit 60 dB below the frame energy), which an
// fp16 tensor-core DFT does not give; the GEMMs of the vocoder backbone do run on the tensor cores (gemm.cuh).
#pragma once
#include "ptx.cuh"

namespace zvb {

constexpr int AUD_NFFT = 1024;
constexpr int AUD_NBINS = AUD_NFFT / 2 + 1;
constexpr int AUD_THREADS = 256;
// shared: two complex ping-pong buffers + twiddles exp(-2 pi i m / N), m < N/2
constexpr int AUD_FFT_SMEM = 2 * AUD_NFFT * 8 + (AUD_NFFT / 2) * 8;

__device__ __forceinline__ void fft_twiddles(float2* tw) {
    for (int m = threadIdx.x; m < AUD_NFFT / 2; m += AUD_THREADS) {
        float s, c;
        sincospif(-2.0f * static_cast<float>(m) / static_cast<float>(AUD_NFFT), &s, &c);
        tw[m] = make_float2(c, s);
    }
}
// In: a (N complex), scratch b.  Forward DFT X[k] = sum_n x[n] exp(-2 pi i k n / N); INVERSE uses the conjugate
// twiddles (no 1/N).  Returns the buffer holding the result (natural order).  Ends with a block barrier.
template <bool INVERSE>
__device__ __forceinline__ float2* fft1024(float2* a, float2* b, const float2* tw) {
    float2* in = a;
    float2* out = b;
    __syncthreads();
#pragma unroll 1
    for (int ns = 1; ns < AUD_NFFT; ns <<= 1) {
        const int tstride = AUD_NFFT / (2 * ns);
        for (int j = threadIdx.x; j < AUD_NFFT / 2; j += AUD_THREADS) {
            const int k = j & (ns - 1);
            float2 w = tw[k * tstride];
            if (INVERSE) w.y = -w.y;
            const float2 v0 = in[j];
            const float2 u = in[j + AUD_NFFT / 2];
            const float2 v1 = make_float2(u.x * w.x - u.y * w.y, u.x * w.y + u.y * w.x);
            const int j0 = ((j - k) << 1) + k;
            out[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
            out[j0 + ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        float2* t = in; in = out; out = t;
    }
    return in;
}

// ------------------------------------------------------------------------------------------------ log-mel
// wav [B][s_pitch] fp32, lens [B] samples; out [B][T][n_mels]: frames t < (len + hop/2) / hop hold
// scale * log(max(mel, 1e-7)), later frames are zero.  window [1024]; fb [n_mels][513] with the non-zero bin range
// of every filter in fb_range [n_mels] = (lo, hi).
__global__ void __launch_bounds__(AUD_THREADS)
fbank_kernel(const float* __restrict__ wav, const int* __restrict__ lens, int B, int s_pitch,
             const float* __restrict__ window, const float* __restrict__ fb, const int2* __restrict__ fb_range,
             int n_mels, int hop, float scale, float* __restrict__ out, int T) {
    extern __shared__ uint8_t aud_smem[];
    float2* a = reinterpret_cast<float2*>(aud_smem);
    float2* b = a + AUD_NFFT;
    float2* tw = b + AUD_NFFT;
    fft_twiddles(tw);
    pdl_wait();
    pdl_launch();
    const long long total = static_cast<long long>(B) * T;
    for (long long f = blockIdx.x; f < total; f += gridDim.x) {
        const int n = static_cast<int>(f / T), t = static_cast<int>(f - static_cast<long long>(n) * T);
        const int S = lens[n];
        const int frames = (S + hop / 2) / hop;
        float* o = out + f * n_mels;
        if (t >= frames) {                                   // block-uniform
            for (int m = threadIdx.x; m < n_mels; m += AUD_THREADS) o[m] = 0.0f;
            continue;
        }
        const float* w = wav + static_cast<long long>(n) * s_pitch;
        __syncthreads();                                     // previous frame's mel reads of `a`/`b` are done
        for (int i = threadIdx.x; i < AUD_NFFT; i += AUD_THREADS) {
            int idx = t * hop - AUD_NFFT / 2 + i;            // center=True, reflect padding by n_fft / 2
            if (idx < 0) idx = -idx;
            if (idx >= S) idx = 2 * (S - 1) - idx;
            idx = idx < 0 ? 0 : idx;                         // signals shorter than the pad (not produced by the reference)
            a[i] = make_float2(w[idx] * window[i], 0.0f);
        }
        float2* X = fft1024<false>(a, b, tw);
        float* mag = reinterpret_cast<float*>(X == a ? b : a);       // the idle buffer
        for (int k = threadIdx.x; k < AUD_NBINS; k += AUD_THREADS) mag[k] = sqrtf(X[k].x * X[k].x + X[k].y * X[k].y);
        __syncthreads();
        for (int m = threadIdx.x; m < n_mels; m += AUD_THREADS) {
            const int2 r = fb_range[m];
            const float* fr = fb + static_cast<long long>(m) * AUD_NBINS;
            float acc = 0.0f;
            for (int k = r.x; k < r.y; ++k) acc = fmaf(fr[k], mag[k], acc);
            o[m] = scale * logf(fmaxf(acc, 1e-7f));
        }
    }
}

// ------------------------------------------------------------------------------------------------ vocoder
// rows [N*T]: mask[row] = 1 where t >= lens[n] (frames past the utterance)
__global__ void voc_mask_kernel(const int* __restrict__ lens, uint8_t* __restrict__ mask, int N, int T) {
    pdl_wait();
    pdl_launch();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * T) return;
    mask[idx] = (idx % T) >= lens[idx / T] ? 1 : 0;
}

// A[(n,t)][k*C + c] = scale * mel[n][t + k - KW/2][c] inside the utterance, 0 outside / for t >= len (the
// zero padding of Conv1d(C, dim, KW, padding=KW/2) applied to each utterance alone); pitch ldA (zero padded).
__global__ void voc_window_kernel(const float* __restrict__ mel, const int* __restrict__ lens, __half* __restrict__ A,
                                  int N, int T, int C, int KW, int ldA, float scale) {
    pdl_wait();
    pdl_launch();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * T * ldA) return;
    const int col = static_cast<int>(idx % ldA);
    const long long row = idx / ldA;
    const int n = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<long long>(n) * T);
    const int len = lens[n];
    float v = 0.0f;
    if (col < KW * C && t < len) {
        const int k = col / C, c = col - k * C;
        const int ts = t + k - KW / 2;
        if (ts >= 0 && ts < len) v = scale * mel[(static_cast<long long>(n) * T + ts) * C + c];
    }
    A[idx] = f2h(v);
}

// LayerNorm over C channels (eps inside the square root, biased variance: torch.nn.LayerNorm), fp16 in / out,
// fp32 statistics; one warp per row, C % 256 == 0, C <= 1024.  mask (nullable): rows with mask != 0 are written as zeros.
template <int KMAX>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ out, const float* __restrict__ w,
                 const float* __restrict__ b, const uint8_t* __restrict__ mask, long long rows, int C, float eps) {
    pdl_wait();
    pdl_launch();
    const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nk = C >> 8;
    float v[KMAX][8];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < nk) {
            unpack8(*reinterpret_cast<const uint4*>(x + row * C + (k * 32 + lane) * 8), v[k]);
#pragma unroll
            for (int e = 0; e < 8; ++e) s += v[k][e];
        }
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) s += __shfl_xor_sync(0xffffffffu, s, q);
    const float mean = s / static_cast<float>(C);
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < nk) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = v[k][e] - mean; ss = fmaf(d, d, ss); }
        }
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, q);
    const float rstd = rsqrtf(ss / static_cast<float>(C) + eps);
    const bool zero = mask != nullptr && mask[row] != 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < nk) {
            const int c0 = (k * 32 + lane) * 8;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = zero ? 0.0f : fmaf((v[k][e] - mean) * rstd, __ldg(w + c0 + e), __ldg(b + c0 + e));
            *reinterpret_cast<uint4*>(out + row * C + c0) = pack8(o);
        }
}

// Head output S [rows][ld] fp32 = [log-magnitude (513) | phase (513)] per frame -> windowed time frame [rows][1024]:
// spectrum = min(exp(mag), 100) * (cos p + i sin p), irfft (1/N, imaginary parts of DC / Nyquist ignored), * window.
// Frames with mask != 0 are skipped (never read by the overlap-add).
__global__ void __launch_bounds__(AUD_THREADS)
voc_istft_frames_kernel(const float* __restrict__ S, int ld, const uint8_t* __restrict__ mask,
                        const float* __restrict__ window, float* __restrict__ frames, long long rows) {
    extern __shared__ uint8_t aud_smem[];
    float2* a = reinterpret_cast<float2*>(aud_smem);
    float2* b = a + AUD_NFFT;
    float2* tw = b + AUD_NFFT;
    fft_twiddles(tw);
    pdl_wait();
    pdl_launch();
    for (long long f = blockIdx.x; f < rows; f += gridDim.x) {
        if (mask[f] != 0) continue;                          // block-uniform
        const float* sp = S + f * ld;
        __syncthreads();
        for (int k = threadIdx.x; k < AUD_NBINS; k += AUD_THREADS) {
            const float mag = fminf(expf(sp[k]), 100.0f);
            float sn, cs;
            sincosf(sp[AUD_NBINS + k], &sn, &cs);
            float re = mag * cs, im = mag * sn;
            if (k == 0 || k == AUD_NFFT / 2) im = 0.0f;
            a[k] = make_float2(re, im);
            if (k != 0 && k != AUD_NFFT / 2) a[AUD_NFFT - k] = make_float2(re, -im);     // Hermitian half
        }
        float2* x = fft1024<true>(a, b, tw);
        float* o = frames + f * AUD_NFFT;
        for (int i = threadIdx.x; i < AUD_NFFT; i += AUD_THREADS) o[i] = x[i].x * (1.0f / AUD_NFFT) * window[i];
    }
}

// wav[n][p], p < hop * (len_n - 1): overlap-add of the windowed frames divided by the window envelope, with the
// n_fft / 2 samples torch.istft(center=True) trims from both ends; zeros beyond.  pitch = hop * (T - 1).
__global__ void voc_overlap_add_kernel(const float* __restrict__ frames, const int* __restrict__ lens,
                                       const float* __restrict__ window, float* __restrict__ wav, int N, int T, int hop,
                                       int clamp) {
    pdl_wait();
    pdl_launch();
    const int pitch = hop * (T - 1);
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * pitch) return;
    const int n = static_cast<int>(idx / pitch), p = static_cast<int>(idx - static_cast<long long>(n) * pitch);
    const int len = lens[n];
    float v = 0.0f;
    if (p < hop * (len - 1)) {
        const int q = p + AUD_NFFT / 2;
        int f_hi = q / hop;
        if (f_hi > len - 1) f_hi = len - 1;
        int f_lo = (q - AUD_NFFT) / hop + 1;                 // smallest f with q - f*hop < n_fft
        if (q - AUD_NFFT < 0) f_lo = 0;
        float acc = 0.0f, env = 0.0f;
        for (int f = f_lo; f <= f_hi; ++f) {
            const int i = q - f * hop;
            const float w = window[i];
            acc += frames[(static_cast<long long>(n) * T + f) * AUD_NFFT + i];
            env = fmaf(w, w, env);
        }
        v = acc / env;
        if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
    }
    wav[idx] = v;
}

}  // namespace zvb
