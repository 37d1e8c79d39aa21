// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C = epilogue(A · Bᵀ)
//   A: (rows, K) bf16 K-major, B: (cols, K) bf16 K-major (an nn.Linear weight, or Vᵀ),
//   both streamed by TMA (128B swizzle) through a 4-stage mbarrier ring; fp32 accumulators
//   live in TMEM (2 stages x <=256 columns) so the epilogue of tile i overlaps the MMAs of
//   tile i+1.  Warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
// Serves every dense contraction of the TTSZipformer forward (reference:
// modules/zipformer.py:1172,1377,1393,1434-1437,1511,1534,1542,1655,1678, 265, 291).
#pragma once
#include "ptx.cuh"

namespace zvb {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;   // 16 KB
constexpr int GEMM_B_BYTES = 256 * GEMM_BLOCK_K * 2;            // 32 KB (block_n <= 256)
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_TMEM_COLS = 512;

enum { EPI_LINEAR = 0, EPI_GATED = 1, EPI_MUL = 2 };
enum { ACT_NONE = 0, ACT_SWOOSH_L = 1, ACT_SWOOSH_R = 2 };
enum { GATE_TANH_SX = 1, GATE_GLU_XS = 2 };

struct GemmParams {
    // problem / tiling
    int M;                 // valid rows per batch (A rows beyond M are TMA zero-fill, never stored)
    int n_out;             // valid output columns per batch
    int num_k_blocks;
    int block_n;           // UMMA N: multiple of 16 (32 for EPI_GATED), <= 256
    int num_m_tiles, num_n_tiles, batches;
    int a_zb, a_zn;        // A tensor-map z = b*a_zb + n_tile*a_zn
    int b_zb;              // B tensor-map z = b*b_zb
    // output
    void* out;             // bf16 (or fp32 when out_f32), row = b*M + m
    int ldc;
    int out_f32;
    int out_col_stride;    // first output column of a tile = n_tile*out_col_stride
    int n_valid;           // valid accumulator columns inside one tile
    // epilogue operands (nullable)
    const float* bias;     // [num_n_tiles*block_n], accumulator-column indexed
    const float* rowbias;  // [(row / rows_per_group)*ld_rowbias + outcol]
    int rows_per_group;
    int ld_rowbias;
    const __nv_bfloat16* resid;
    int ldr;
    const __nv_bfloat16* orig;      // bypass: orig + (v - orig)*scale[col]
    const float* bypass_scale;
    int act;
    int gate_mode;
    const uint8_t* row_mask;        // [rows] non-zero -> output row is zero
    const __nv_bfloat16* mul;
    int ldm;
    // transposed store: dst[(row / t_L)*t_batch_rows + drow(col)][row % t_L], pitch t_pitch,
    // drow(col) = col + (col / t_hd)*(t_hp - t_hd)
    int transposed;
    int t_L, t_pitch, t_batch_rows, t_hd, t_hp;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_SWOOSH_L) return swoosh_l(v);
    if (act == ACT_SWOOSH_R) return swoosh_r(v);
    return v;
}

// Stores 16 consecutive output columns [col0, col0+16) of one row.
__device__ __forceinline__ void store_row16(const GemmParams& p, long long row, int col0, int ncols,
                                            const float* v) {
    if (p.transposed) {
        const int n = static_cast<int>(row / p.t_L);
        const int l = static_cast<int>(row - static_cast<long long>(n) * p.t_L);
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < ncols) {
                const int c = col0 + i;
                const int drow = c + (c / p.t_hd) * (p.t_hp - p.t_hd);
                dst[(static_cast<long long>(n) * p.t_batch_rows + drow) * p.t_pitch + l] =
                    __float2bfloat16(v[i]);
            }
        }
        return;
    }
    if (p.out_f32) {
        float* dst = reinterpret_cast<float*>(p.out) + row * p.ldc + col0;
        if (ncols == 16 && (p.ldc & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < ncols) dst[i] = v[i];
        }
        return;
    }
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + col0;
    if (ncols == 16 && (p.ldc & 7) == 0 && (col0 & 7) == 0) {
        uint4 a, b;
        a.x = pack_bf16(v[0], v[1]);   a.y = pack_bf16(v[2], v[3]);
        a.z = pack_bf16(v[4], v[5]);   a.w = pack_bf16(v[6], v[7]);
        b.x = pack_bf16(v[8], v[9]);   b.y = pack_bf16(v[10], v[11]);
        b.z = pack_bf16(v[12], v[13]); b.w = pack_bf16(v[14], v[15]);
        *reinterpret_cast<uint4*>(dst) = a;
        *reinterpret_cast<uint4*>(dst + 8) = b;
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < ncols) dst[i] = __float2bfloat16(v[i]);
    }
}

// Loads 16 bf16 of one row as floats (vectorised when aligned, zero beyond ncols).
__device__ __forceinline__ void load_row16(const __nv_bfloat16* base, long long row, int ld, int col0,
                                           int ncols, float* v) {
    const __nv_bfloat16* src = base + row * ld + col0;
    if (ncols == 16 && (ld & 7) == 0 && (col0 & 7) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(src + 8));
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[2 * i] = bf16_lo(w[i]);
            v[2 * i + 1] = bf16_hi(w[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = i < ncols ? __bfloat162float(src[i]) : 0.0f;
    }
}

template <int KIND>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
    uint64_t* full_bar = bars;                        // [STAGES] TMA -> MMA
    uint64_t* empty_bar = bars + GEMM_STAGES;         // [STAGES] MMA -> TMA
    uint64_t* tmem_full = bars + 2 * GEMM_STAGES;     // [2] MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * GEMM_STAGES + 2;  // [2] epilogue -> MMA
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.batches * p.num_m_tiles * p.num_n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, GEMM_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint32_t stage_bytes = GEMM_A_BYTES + static_cast<uint32_t>(p.block_n) * GEMM_BLOCK_K * 2;
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.num_n_tiles;
                const int rest = tile / p.num_n_tiles;
                const int m_tile = rest % p.num_m_tiles;
                const int b = rest / p.num_m_tiles;
                const int az = b * p.a_zb + n_tile * p.a_zn;
                const int bz = b * p.b_zb;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                    uint8_t* sa = smem + stage * GEMM_STAGE_BYTES;
                    tma_load_3d(sa, &tma_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_tile * GEMM_BLOCK_M, az);
                    tma_load_3d(sa + GEMM_A_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BLOCK_K,
                                n_tile * p.block_n, bz);
                    if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(static_cast<uint32_t>(p.block_n));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc) * 256u;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * GEMM_STAGE_BYTES);
                    const uint64_t da = umma_desc_k_sw128(sa);
                    const uint64_t db = umma_desc_k_sw128(sa + GEMM_A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes inside the 128B swizzle atom: +2 in >>4 units
                        umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                                  idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);           // frees the smem slot when MMAs retire
                    if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1u; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (4 warps)
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.num_n_tiles;
            const int rest = tile / p.num_n_tiles;
            const int m_tile = rest % p.num_m_tiles;
            const int b = rest / p.num_m_tiles;
            const int m = m_tile * GEMM_BLOCK_M + quarter * 32 + lane;
            const bool row_ok = m < p.M;
            const long long row = static_cast<long long>(b) * p.M + m;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc) * 256u +
                                   (static_cast<uint32_t>(quarter * 32) << 16);
            const int out_base = n_tile * p.out_col_stride;
            const int acc_base = n_tile * p.block_n;
            bool masked = false;
            if (p.row_mask != nullptr && row_ok) masked = p.row_mask[row] != 0;
            long long grp = 0;
            if (p.rowbias != nullptr && row_ok) grp = row / p.rows_per_group;

            if (KIND == EPI_GATED) {
                const int half = p.block_n >> 1;
                for (int c0 = 0; c0 < half; c0 += 16) {
                    uint32_t ra[16], rb[16];
                    tmem_ld16(taddr + c0, ra);
                    tmem_ld16(taddr + half + c0, rb);
                    tmem_ld_wait();
                    const int oc = out_base + c0;
                    int ncols = p.n_out - oc;
                    ncols = ncols > 16 ? 16 : ncols;
                    if (row_ok && ncols > 0) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float a = __uint_as_float(ra[i]) + __ldg(p.bias + acc_base + c0 + i);
                            const float g = __uint_as_float(rb[i]) + __ldg(p.bias + acc_base + half + c0 + i);
                            float r = p.gate_mode == GATE_TANH_SX ? g * fast_tanh(a) : a * fast_sigmoid(g);
                            v[i] = masked ? 0.0f : r;
                        }
                        store_row16(p, row, oc, ncols, v);
                    }
                }
            } else {
                for (int c0 = 0; c0 < p.block_n; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                    const int oc = out_base + c0;
                    int ncols = p.n_valid - c0;
                    if (p.n_out - oc < ncols) ncols = p.n_out - oc;
                    ncols = ncols > 16 ? 16 : ncols;
                    if (row_ok && ncols > 0) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                        if (KIND == EPI_MUL) {
                            float y[16];
                            load_row16(p.mul, row, p.ldm, oc, ncols, y);
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] *= y[i];
                        } else {
                            if (p.bias != nullptr) {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (i < ncols) v[i] += __ldg(p.bias + acc_base + c0 + i);
                            }
                            if (p.rowbias != nullptr) {
                                const float* rbp = p.rowbias + grp * p.ld_rowbias + oc;
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (i < ncols) v[i] += __ldg(rbp + i);
                            }
                            if (p.act != ACT_NONE) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], p.act);
                            }
                            if (p.resid != nullptr) {
                                float y[16];
                                load_row16(p.resid, row, p.ldr, oc, ncols, y);
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] += y[i];
                            }
                            if (p.orig != nullptr) {
                                float o[16];
                                load_row16(p.orig, row, p.ldr, oc, ncols, o);
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (i < ncols) v[i] = o[i] + (v[i] - o[i]) * __ldg(p.bypass_scale + oc + i);
                            }
                        }
                        store_row16(p, row, oc, ncols, v);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
    }
}

}  // namespace zvb
