/* zipvoice_b200 — C ABI of the B200-native ZipVoice sampler hot path.
 *
 * The reference (ayutaz/ZipVoice) has no C/FFI plugin API; its operator seam is attribute
 * replacement on the model object as done by `load_trt`
 * (reference: zipvoice/utils/tensorrt.py:128-143).  These entry points are what a binding for
 * that seam calls:
 *   seam 1  model.fm_decoder(x, t, padding_mask, guidance_scale)   -> zvb_decoder_forward_f32
 *           (reference: zipvoice/models/zipvoice.py:180-184, modules/zipformer.py:242-293,
 *            zipvoice/utils/tensorrt.py:69-126)
 *   seam 2  model.solver.sample(x, text_condition, speech_condition, padding_mask, num_step,
 *           guidance_scale, t_start, t_end, t_shift)                -> zvb_sample
 *           (reference: modules/solver.py:182-240, 40-110, 113-165)
 *   text    model.text_encoder(x, t=None, padding_mask)             -> zvb_decoder_forward_f32
 *           on a plan built from the text-encoder description (reference: zipvoice.py:209-211)
 *
 * Conventions: plain pointers and sizes only; every device buffer (weights, workspace,
 * inputs, outputs) is allocated and owned by the caller (PyTorch); the library never
 * allocates device memory and never synchronises, so every call is CUDA-graph capturable on
 * the given stream.  All functions return 0 on success, a negative zvb_status otherwise
 * (no exceptions cross the ABI); zvb_last_error() describes the last failure of the calling
 * thread.  One host thread drives one plan at a time; distinct plans may run concurrently
 * on distinct streams (cf. the TensorRT context pool, reference: tensorrt.py:44-52).
 */
#ifndef ZIPVOICE_B200_H_
#define ZIPVOICE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZVB_ABI_VERSION 6
#define ZVB_MAX_STACKS 8
#define ZVB_MAX_LAYERS 64
#define ZVB_VOC_MAX_LAYERS 16

typedef enum {
    ZVB_OK = 0,
    ZVB_ERR_INVALID = -1,     /* bad argument / unsupported shape */
    ZVB_ERR_WORKSPACE = -2,   /* workspace too small */
    ZVB_ERR_CUDA = -3,        /* CUDA runtime / driver error */
    ZVB_ERR_NO_DEVICE = -4    /* no sm_100 device: there is no CPU fallback */
} zvb_status;

/* One nn.Linear: W is fp16 (IEEE half) row-major (out_features, k_pitch) with k_pitch a multiple of 8,
 * zero padded; bias is fp32 (nullable).  "Gated" projections are re-ordered per 256-row tile
 * as [128 rows of the first operand | 128 rows of the second] and zero padded to a multiple
 * of 256 rows (bias likewise), see zipvoice_b200/weights.py. */
typedef struct {
    const void* w;
    const float* b;
    int32_t out_features;
    int32_t in_features;
    int32_t k_pitch;
    int32_t rows;             /* rows physically present in w (>= out_features for gated) */
} zvb_linear;

/* Zipformer2EncoderLayer (reference: modules/zipformer.py:370-404) */
typedef struct {
    zvb_linear attn_in;       /* self_attn_weights.in_proj : D -> H*(2*32+4)               */
    const void* pos_table;    /* E = linear_pos(pos_emb) for this plan's L, folded on the host into
                               * [H][2L-1+256] 16-byte entries of fp16 column pairs {log2e*E[r][d],
                               * log2e*E[r+1][d]} (d = 0..3, entry = r + 128, zeros outside), followed by
                               * [H] fp32 max_r |E[h][r]|_2 (zipvoice_b200/weights.py: pack_pos_table) */
    const void* pos_table_tc; /* the same E for the tensor-core bias (csrc/attn3.cuh): [H][2][LZ] entries of 4 fp16 =
                               * log2e*E[r][0..3] at index 128 + r of copy 0 (zeros elsewhere), copy 1 = copy 0 shifted by
                               * one entry, LZ = even(2L + 264); followed by [H] fp32 max_r |log2e*E[h][r]|_2
                               * (zipvoice_b200/weights.py: pack_pos_table_tc).  null: the CUDA-core kernel is used */
    zvb_linear ff_in[3], ff_out[3];
    zvb_linear ff1_attn;      /* rows [feed_forward1.in_proj ; self_attn_weights.in_proj] in ONE weight (both read the layer input;
                               * the time embedding of feed_forward1's input becomes a per-utterance row bias W1*temb).
                               * w == null: the two projections run as separate GEMMs (attn_in, ff_in[0]) */
    zvb_linear na_sx;         /* nonlin_attention.in_proj rows (s,x), gated-packed          */
    zvb_linear na_y;          /* nonlin_attention.in_proj rows y                            */
    zvb_linear na_out;
    zvb_linear sa_in[2], sa_out[2];
    zvb_linear conv_in[2];    /* conv_module{1,2}.in_proj rows (x,s), gated-packed          */
    const float* dw_w[2];     /* depthwise weight transposed to fp32 [K][D]                 */
    const float* dw_b[2];
    zvb_linear conv_out[2];
    const float* norm_bias;   /* BiasNorm.bias [D]                                          */
    const float* norm_log_scale; /* BiasNorm.log_scale, device scalar                       */
    const float* bypass_scale;
    const float* bypass_mid_scale;
} zvb_layer;

typedef struct {
    int32_t downsample;       /* 1, 2 or 4                                                   */
    int32_t num_layers;
    int32_t conv_kernel;      /* 7, 9, 15 or 31                                              */
    int32_t first_layer;      /* index into zvb_model.layers                                 */
    float ds_weights[4];      /* softmax(downsample.bias) (reference: zipformer.py:906)      */
    const float* out_combiner_scale;  /* [D], null when downsample == 1                      */
    const float* time_w;      /* encoders.i.time_emb.1.weight fp32 [D][time_dim] (nullable)  */
    const float* time_b;
} zvb_stack;

/* TTSZipformer (reference: modules/zipformer.py:109-240; two-stream: the caller passes the
 * in/out projection pair selected by input width, zipformer_two_stream.py:236-262) */
typedef struct {
    int32_t abi_version;
    int32_t dim;              /* D                                                           */
    int32_t num_heads;        /* H (query_head_dim 32, pos_head_dim 4 are fixed)             */
    int32_t value_head_dim;   /* 12                                                          */
    int32_t in_dim;           /* 300 / 500 / 192                                             */
    int32_t out_dim;          /* 100 / 200                                                   */
    int32_t ff_dims[3];
    int32_t na_hidden;
    int32_t time_dim;         /* 192, or 0 when the network takes no time embedding          */
    int32_t use_guidance_embed;
    int32_t guidance_dim;     /* width of timestep_embedding(guidance_scale) = guidance_w's in_features
                               * (reference: zipformer.py:128, 233-238: 192, independent of time_dim)   */
    int32_t num_stacks;
    int32_t num_layers;
    zvb_linear in_proj, out_proj;
    const float* time0_w; const float* time0_b;   /* time_embed.0 fp32 [2*time_dim][time_dim] */
    const float* time2_w; const float* time2_b;   /* time_embed.2 fp32 [time_dim][2*time_dim] */
    const float* guidance_w;                      /* guidance_scale_embed.weight fp32 [time_dim][guidance_dim] (nullable) */
    zvb_stack stacks[ZVB_MAX_STACKS];
    const zvb_layer* layers;  /* host array of num_layers entries                            */
} zvb_model;

/* Device buffers inside the workspace a plan reads its inputs from / writes its output to. */
typedef struct {
    void* xin;                /* fp16 [N][T][xin_pitch] = [x | text | speech | 0-pad]        */
    float* t;                 /* [N]                                                         */
    float* g;                 /* [N] guidance scale (distill only)                           */
    uint8_t* mask;            /* [N][T] non-zero = padded frame                              */
    float* out;               /* fp32 [N][T][out_dim]                                        */
    int32_t xin_pitch;
} zvb_io;

typedef struct zvb_plan zvb_plan;

const char* zvb_last_error(void);
int zvb_abi_version(void);
/* sha256 of the sources the library was built from (zipvoice_b200/build.py rebuilds on mismatch) */
const char* zvb_source_hash(void);
/* number of kernels of this library launched by the calling thread since process start */
long long zvb_launch_count(void);

/* Workspace size (bytes) for a plan over N rows of T frames.  The workspace must be zero
 * filled once by the caller before zvb_plan_create. */
int zvb_plan_workspace_bytes(const zvb_model* model, int N, int T, size_t* bytes);
int zvb_plan_create(const zvb_model* model, int N, int T, void* workspace, size_t workspace_bytes,
                    zvb_plan** plan);
void zvb_plan_destroy(zvb_plan* plan);
int zvb_plan_io(const zvb_plan* plan, zvb_io* io);

/* Forward over the resident io buffers: out = fm_decoder(xin, t, mask[, g]). */
int zvb_decoder_forward(zvb_plan* plan, void* stream);

/* Parity aid: with a non-null device counter every forward of the plan also counts, kernel by kernel, the
 * fp16 outputs that reached the largest finite fp16 magnitude (the conversions of the path saturate) or are
 * inf / NaN, and adds them to *counter.  null switches the check off again. */
int zvb_plan_set_saturation_counter(zvb_plan* plan, unsigned long long* counter);

/* Profiling aid (NOT graph capturable: records one CUDA event per kernel and synchronises the
 * stream): runs one forward over the resident io buffers and returns, per launched kernel, its
 * duration in ms, its zvb_op_category, its GEMM shape and its algorithmic work (FLOPs for the tensor-core
 * kernels, bytes for the memory-bound ones). */
typedef enum {
    ZVB_CAT_GEMM_LINEAR = 0, ZVB_CAT_GEMM_GATED = 1, ZVB_CAT_GEMM_PV = 2, ZVB_CAT_ATTN_WEIGHTS = 3,
    ZVB_CAT_BIASNORM = 4, ZVB_CAT_ELEMENTWISE = 5, ZVB_CAT_RESAMPLE = 6, ZVB_CAT_DWCONV = 7,
    ZVB_CAT_OTHER = 8
} zvb_op_category;
int zvb_decoder_profile(zvb_plan* plan, void* stream, int max_ops, float* ms, int* category, double* work,
                        double* bytes /* nullable: algorithmic HBM bytes of every kernel */,
                        int* shapes /* nullable, 4 ints per kernel: rows, cols, K, tile N */, int* num_ops);

/* Seam 1: x fp32 [N][T][in_dim], t fp32 [N] (null for the text encoder), mask u8 [N][T],
 * g fp32 [N] or null, out fp32 [N][T][out_dim]. */
int zvb_decoder_forward_f32(zvb_plan* plan, const float* x, const float* t, const uint8_t* mask,
                            const float* g, float* out, void* stream);

/* Seam 2: Euler ODE with classifier-free guidance.
 *   x        fp32 [B][T][F]   state, updated in place (x0 in, x(t_end) out)
 *   text     fp32 [B][T][Ft]  ; speech fp32 [B][T][F] ; mask u8 [B][T]
 *   guidance fp32 [B] per-utterance scale (device)
 *   ts       fp32 [num_step+1] time grid (device), ts_host the same values on the host
 *   mode     0: single pass, no guidance input (all scales zero)
 *            1: CFG, batch doubled [uncond ; cond] (plan N == 2B), scale doubled and speech
 *               kept for the uncond half while t <= 0.5 (reference: solver.py:83-110)
 *            2: distilled model: single pass, guidance fed to the network (plan N == B)
 *   vrec     optional fp32 [num_step][B][T][F]: every step's (blended) velocity */
int zvb_sample(zvb_plan* plan, float* x, const float* text, const float* speech, const uint8_t* mask,
               const float* guidance, const float* ts, const float* ts_host, int num_step, int mode,
               int B, int F, int Ft, float* vrec, void* stream);

/* ---- the stages either side of the sampler ------------------------- */

/* Prompt log-mel, the reference's VocosFbank (reference: zipvoice/utils/feature.py:47-116: torchaudio
 * MelSpectrogram(sample_rate 24000, n_fft 1024, hop 256, n_mels 100, center=True (reflect), power=1) followed by
 * clamp(min=1e-7).log(), trimmed to lhotse's compute_num_frames = (samples + hop/2) / hop frames).
 *   wav      fp32 [B][s_pitch] (device), lens int32 [B] valid samples per row (device)
 *   window   fp32 [1024] (torch.hann_window, periodic); fb fp32 [n_mels][513] mel filterbank; fb_range int32
 *            [n_mels][2] = first / one-past-last non-zero bin of every filter
 *   out      fp32 [B][T][n_mels]: scale * log-mel for t < (len + hop/2) / hop, zeros after
 * n_fft is fixed at 1024 (the reference's VocosFbankConfig). */
int zvb_fbank(const float* wav, const int32_t* lens, int B, int s_pitch, const float* window, const float* fb,
              const int32_t* fb_range, int n_mels, int hop, float scale, float* out, int T, void* stream);

/* Vocoder: `vocoder.decode(mel)` of the reference's call sites (zipvoice/bin/infer_zipvoice.py:301-312, 409, 594) =
 * the external vocos 0.1.0 model (uv.lock:1239-1241) VocosBackbone + ISTFTHead(padding="center"):
 *   Conv1d(n_mels, dim, 7, padding 3) -> LayerNorm(eps 1e-6) -> n_layers x ConvNeXtBlock [depthwise Conv1d(dim, 7) ->
 *   LayerNorm -> Linear(dim, intermediate) -> GELU -> Linear(intermediate, dim) -> * gamma -> + residual] ->
 *   LayerNorm -> Linear(dim, n_fft + 2) -> (mag, phase) -> exp, clip 1e2 -> mag*(cos p + i sin p) -> torch.istft.
 * Every utterance of a batch is decoded as if alone (frames past its length are zero padding of the convolutions). */
typedef struct {
    const float* dw_w;        /* depthwise taps fp32 [7][dim] */
    const float* dw_b;        /* fp32 [dim] */
    const float* ln_w; const float* ln_b;
    zvb_linear pw1;           /* Linear(dim, intermediate), GELU in the epilogue */
    zvb_linear pw2;           /* Linear(intermediate, dim) with the block's layer scale gamma folded in */
} zvb_voc_layer;

typedef struct {
    int32_t abi_version;
    int32_t dim, intermediate, n_layers, n_mels, n_fft, hop, kernel;   /* 512, 1536, 8, 100, 1024, 256, 7 */
    zvb_linear embed;         /* Conv1d as a linear over the 7-frame window: W[dim][k * n_mels + c] (k_pitch % 8 == 0) */
    const float* norm_w; const float* norm_b;
    zvb_voc_layer layers[ZVB_VOC_MAX_LAYERS];
    const float* final_w; const float* final_b;
    zvb_linear head;          /* Linear(dim, n_fft + 2): rows [0, n_fft/2] log-magnitude, the rest phase */
    const float* window;      /* fp32 [n_fft] (head.istft.window) */
} zvb_vocoder;

typedef struct zvb_vocoder_plan zvb_vocoder_plan;
int zvb_vocoder_workspace_bytes(const zvb_vocoder* voc, int N, int T, size_t* bytes);
int zvb_vocoder_create(const zvb_vocoder* voc, int N, int T, void* workspace, size_t workspace_bytes,
                       zvb_vocoder_plan** plan);
void zvb_vocoder_destroy(zvb_vocoder_plan* plan);
/* mel fp32 [N][T][n_mels] (device; multiplied by `scale`, the caller's 1 / feat_scale), lens int32 [N] frames per
 * utterance (device, 1 <= len <= T); wav fp32 [N][hop * (T - 1)]: hop * (len - 1) samples per row, zeros after;
 * clamp != 0 applies the caller's .clamp(-1, 1) (infer_zipvoice.py:409). */
int zvb_vocoder_decode(zvb_vocoder_plan* plan, const float* mel, const int32_t* lens, float scale, int clamp, float* wav,
                       void* stream);
/* per-kernel timing of one decode over the resident buffers, as zvb_decoder_profile */
int zvb_vocoder_profile(zvb_vocoder_plan* plan, void* stream, int max_ops, float* ms, int* category, double* work,
                        double* bytes, int* num_ops);

/* ---- single-kernel entry points used by the parity tests -------------------------------- */
/* C = epilogue(A·Wᵀ): A fp16 [M][lda], W fp16 [n_out][k_pitch]; act 0 none/1 SwooshL/2 SwooshR/3 GELU (erf);
 * resid fp16 [M][ldc] nullable (added); orig fp16 [M][ldc] + bypass_scale fp32 [n_out] nullable:
 * C = orig + (C - orig)*scale (needs resid); out_mode 0: out fp16 / 1: out fp32. */
int zvb_test_linear(const void* A, int M, int K, int lda, const void* W, const float* bias, int n_out,
                    int k_pitch, int block_n, int act, const void* resid, const void* orig,
                    const float* bypass_scale, void* out, int ldc, int out_mode, void* stream);
/* P receives the unnormalised weights 2^12 exp(s - m), inv_l [N][H][L] the reciprocal row sums;
 * pos_table as in zvb_layer; scratch: N * 4 * ceil(L/128) 32-bit words (excluded-key bits) */
int zvb_test_attn_weights(const void* qkp, int ld, const void* pos_table, const uint8_t* mask, void* scratch,
                          void* P, float* inv_l, int N, int H, int L, int Lk, void* stream);
int zvb_test_attn_weights_tc(const void* qkp, int ld, const void* pos_table_tc, const uint8_t* mask, void* scratch,
                             void* P, float* inv_l, int N, int H, int L, int Lk, void* stream);
/* mul: fp16 [N*L][hd] gate of NonlinAttention (per_head == 0 only, nullable) */
int zvb_test_pv(const void* P, const float* inv_l, const void* Vt, void* out, int N, int H, int L, int Lk, int hd,
                int hp, int per_head, const void* mul, void* stream);
/* gated projection on tile-packed weights (rows = tiles*256); gate_mode 1: x*tanh(s), 2: GLU */
int zvb_test_gated(const void* A, int M, int K, int lda, const void* W, const float* bias, int rows, int n_out,
                   int k_pitch, int gate_mode, const uint8_t* row_mask, void* out, int ldc, void* stream);
/* src, orig, out, out_t (= out + temb[row / rows_per_group], nullable): fp16 [rows][C] */
int zvb_test_biasnorm_bypass(const void* src, const void* orig, void* out, void* out_t,
                             const float* temb, int rows_per_group, const float* nbias, const float* log_scale,
                             const float* bscale, long long rows, int C, void* stream);
int zvb_test_dwconv(const void* x, void* out, const float* wt, const float* bias, int N, int L, int C, int K,
                    void* stream);
/* the same depthwise convolution without the activation (vocoder ConvNeXt blocks) */
int zvb_test_dwconv_linear(const void* x, void* out, const float* wt, const float* bias, int N, int L, int C, int K,
                           void* stream);
/* torch.nn.LayerNorm over C (fp16 in / out, C % 256 == 0, <= 1024); rows with mask != 0 (nullable) are zeroed */
int zvb_test_layernorm(const void* x, void* out, const float* w, const float* b, const uint8_t* mask, long long rows, int C,
                       float eps, void* stream);
/* linear with a row mask: rows with mask != 0 are written as zeros (resid fp16 nullable, as zvb_test_linear) */
int zvb_test_linear_masked(const void* A, int M, int K, int lda, const void* W, const float* bias, int n_out, int k_pitch,
                           int act, const void* resid, const uint8_t* row_mask, void* out, int ldc, void* stream);
/* inverse STFT of head outputs S fp32 [N*T][ld] = [log-mag 513 | phase 513]: frames scratch fp32 [N*T][1024],
 * mask scratch u8 [N*T]; wav fp32 [N][hop * (T - 1)] */
int zvb_test_istft(const float* S, int ld, const int32_t* lens, const float* window, float* frames, uint8_t* mask, float* wav,
                   int N, int T, int hop, int clamp, void* stream);
int zvb_test_cfg_euler(float* x, const float* v, const float* guidance, float gscale, const float* ts,
                       int step, int B, long long per_utt, int cfg, void* stream);
/* linear with the transposed store of the value projections: out[(row / t_L)*t_batch_rows + drow(col)][row % t_L],
 * pitch t_pitch, drow(col) = col + (col / t_hd)*(t_hp - t_hd) (SelfAttention heads padded t_hd -> t_hp rows) */
int zvb_test_linear_t(const void* A, int M, int K, int lda, const void* W, const float* bias, int n_out, int k_pitch,
                      void* out, int t_L, int t_pitch, int t_batch_rows, int t_hd, int t_hp, void* stream);
/* SimpleDownsample / SimpleUpsample + out_combiner on fp16 [N][L][C] (reference: zipformer.py:873-935) */
int zvb_test_downsample(const void* src, void* out, int N, int L, int ds, const float* w4, int C, void* stream);
int zvb_test_upsample_combine(const void* orig, const void* y, void* out, const float* scale, int N, int L, int ds,
                              int C, void* stream);
/* xt = x + temb[row / rows_per_group] */
int zvb_test_stream_prep(const void* x, void* xt, const float* temb, int rows_per_group, long long rows, int C,
                         void* stream);
/* decoder input [x | text | speech | 0] as fp16 with the CFG doubling (reference: solver.py:83-98) */
int zvb_test_assemble_input(const float* x, const float* text, const float* speech, void* xin, int B, int T, int F,
                            int Ft, int ldx, int cfg, int drop_speech, void* stream);
/* out = act_out(addend + bias + W * act_in(in)), fp32 (time-embedding MLPs); act 0 none / 2 SwooshR */
int zvb_test_small_linear(const float* in, const float* W, const float* bias, const float* addend, float* out,
                          int N, int K, int O, int act_in, int act_out, void* stream);
int zvb_test_timestep_embedding(const float* t, float* out, int N, int dim, void* stream);
/* strided key mask mask[:, ::ds] and the attention kernel's excluded-key bit words (either output nullable) */
int zvb_test_masks(const uint8_t* mask, int N, int T, int ds, uint8_t* strided, uint32_t* words, void* stream);

/* Host only (no device needed): the launch shape the engine picks for a linear layer of `rows` x `n_out` outputs over `k`
 * inputs on a device with `num_sms` SMs -- tile width, CTA pair (cta_group::2) or single CTA, and the attention kernel's
 * key-split cluster size for `attn_ctas` = query tiles x heads x utterances with `q_tiles` key tiles.  lean_kind: 0 generic
 * epilogue only, 1 plain lean epilogue possible, 2 residual lean epilogue possible.  Any output pointer may be null.
 * Diagnostic: it evaluates the rules for `num_sms` under the library's init lock; do not call it while another thread
 * creates plans. */
int zvb_debug_launch_shape(long long rows, int n_out, int k, int lean_kind, int num_sms, long long attn_ctas, int q_tiles,
                           int* block_n, int* pair, int* attn_split);

#ifdef __cplusplus
}
#endif
#endif /* ZIPVOICE_B200_H_ */
