// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C = epilogue(A · Bᵀ)
//   A: (rows, K) fp16 K-major, B: (cols, K) fp16 K-major (an nn.Linear weight, or Vᵀ), both
//   streamed by TMA (128B swizzle) through a 3..8-stage mbarrier ring; fp32 accumulators live in
//   TMEM (2 stages x <=256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   The epilogue's tile-shaped operands (the fp16 residual stream, the bypass `orig`, or the `y` gate
//   of NonlinAttention) are streamed by a second TMA ring of 16 KB sub-tiles (128 rows x 64 columns),
//   so the epilogue threads never wait on a global load; the result is staged in place over the
//   consumed residual sub-tile and leaves as one TMA box store.
//   Warps 0..15 = epilogue (four warps per TMEM lane quarter), warp 16 = TMA producer (A/B), warp 17 = MMA
//   issuer (+TMEM alloc), warp 18 = TMA producer (aux), warp 19 = TMA store thread.
// Serves every dense contraction of the TTSZipformer forward (reference:
// modules/zipformer.py:1172,1377,1393,1434-1437,1511,1534,1542,1655,1678, 265, 291).
#pragma once
#include <type_traits>
#include "ptx.cuh"

namespace zvb {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;   // 16 KB
constexpr int GEMM_B_BYTES = 256 * GEMM_BLOCK_K * 2;            // 32 KB (block_n <= 256)
// The operand ring is 144 KB; a stage is A (16 KB) + the B rows this CTA stages (block_n, or block_n / 2
// in a CTA pair, x 128 B), so narrow tiles get a deeper ring: 3 stages at block_n = 256, 4 for a CTA
// pair, up to 8 for the 16..48-column tiles that only stream A (more bytes in flight per SM -- those
// kernels are HBM-latency bound).  Host: gemm_ring() fills GemmParams::stages / stage_bytes.
constexpr int GEMM_OPERAND_BYTES = 3 * (GEMM_A_BYTES + GEMM_B_BYTES);        // default operand ring (144 KB)
constexpr int GEMM_MAX_STAGES = 10;     // 10 only for the SelfAttention P.V ring over the unused staging area (engine.cu: build_pv)
constexpr int GEMM_AUX_SLOTS = 4;                               // default aux / staging slots
constexpr int GEMM_AUX_SLOTS_MAX = 8;
constexpr int GEMM_AUX_BYTES = 128 * 128;                       // 128 rows x 128 B
// The operand ring and the aux / staging slots share one 208 KB budget; both are latency bound (bytes in
// flight per SM / L2 latency), so the host gives each GEMM the split its traffic wants (gemm_layout):
//   default          ring 144 KB, 4 aux slots (a tile's residual sub-tiles, staged in place; or 2 staging
//                    buffers per epilogue half);
//   wide ring        no aux operand, TMA stores: 1 staging buffer per half (32 KB), ring 176 KB -- one more
//                    stage in flight for the K = 512 feed-forward input GEMMs;
//   deep aux         K <= 128 (SelfAttention out_proj): the ring shrinks to the k-blocks of two tiles and
//                    the residual / output sub-tiles get up to 8 slots.
constexpr int GEMM_SHARED_BUDGET = GEMM_OPERAND_BYTES + GEMM_AUX_SLOTS * GEMM_AUX_BYTES;
inline void gemm_ring(int block_n, int cluster, int ring_bytes, int* stages, int* stage_bytes) {
    const int sb = GEMM_A_BYTES + (block_n / cluster) * GEMM_BLOCK_K * 2;      // multiple of 1024 (block_n % 16 == 0)
    int st = ring_bytes / sb;
    *stages = st > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : st;
    *stage_bytes = sb;
}
constexpr int GEMM_BIAS_BYTES = 16 * 64 * 4;               // per epilogue warp: bias of the unit in flight (32 columns, 2 x 32 gated)
// Lean epilogues: bias (+ per-utterance row bias) of the tile in flight, [2 tile parities][2 utterances][256 columns] fp32.
// The kernel leaves the L1 no capacity (227 KB of shared memory), so a global bias load costs an L2 round trip per unit.
constexpr int GEMM_CBIAS_BYTES = 2 * 2 * 256 * 4;
constexpr int GEMM_LAYOUT_BYTES = GEMM_SHARED_BUDGET + GEMM_BIAS_BYTES + GEMM_CBIAS_BYTES + 640 /*barriers: 64 x 8 B + TMEM holder*/;
constexpr int GEMM_SMEM_BYTES = 232448;                    // all 227 KB; layout + alignment pad must fit (checked)
static_assert(GEMM_LAYOUT_BYTES <= GEMM_SMEM_BYTES, "shared-memory layout too large");
constexpr int GEMM_THREADS = 640;
constexpr int GEMM_EPI_WARPS = 16;
constexpr int GEMM_TMEM_COLS = 512;

enum { EPI_LINEAR = 0, EPI_GATED = 1 };
enum { ACT_NONE = 0, ACT_SWOOSH_L = 1, ACT_SWOOSH_R = 2, ACT_GELU = 3 };
enum { GATE_TANH_SX = 1, GATE_GLU_XS = 2 };
enum { AUX_NONE = 0, AUX_ADD_H16 = 1, AUX_MUL_H16 = 2 };      // tile operand: residual (added) / gate (multiplied)
enum { OUT_H16 = 0, OUT_F32 = 1, OUT_T_H16 = 3 };

struct GemmParams {
    // problem / tiling
    int M;                 // valid rows per batch (A rows beyond M are TMA zero-fill, never stored)
    int n_out;             // valid output columns per batch
    int num_k_blocks;
    int block_n;           // UMMA N: multiple of 16, <= 256 (256 for EPI_GATED)
    int stages, stage_bytes;   // operand ring (gemm_ring)
    int ring_bytes;            // shared memory given to the operands (resident A + ring); the aux slots follow it
    int pre_b;                 // 1: B is a model weight (written long before the launch): the producer requests the B halves of
                               // its first tile's stages BEFORE the dependency wait (griddepcontrol.wait); 2: the same for A
                               // (P.V: the attention weights are at least two launches old).  Host: plan builders only
    int a_resident;            // A-stationary mode (CTA pairs, K <= 512): the pair walks a CONTIGUOUS range of tiles,
                               // n-tile fastest, keeps the 128 x K A tile of its m-group in shared memory and streams
                               // only B; `a_bytes` = num_k_blocks x 16 KB in front of the (B-only) ring
    int aux_slots;             // 16 KB slots of the aux / staging area (2 .. GEMM_AUX_SLOTS_MAX)
    int stage_depth;           // staging buffers per epilogue half on the aux-less TMA-store path (1 or 2)
    int a_bytes;
    int fast_resid;            // lean epilogue of the residual-stream GEMMs (EPI_LINEAR, no activation, fp16 residual tile
                               // through the aux ring, fp16 TMA-store output, optional per-utterance row bias; no bypass /
                               // row scale / row mask; n_out % block_n == 0, block_n % 64 == 0)
    int fast_epi;              // lean epilogue (EPI_LINEAR, fp16 TMA-store output, no tile operand / row scale / row bias /
                               // row mask, block_n % 64 == 0, n_out % 32 == 0): the feed-forward input GEMMs
    int num_m_tiles, num_n_tiles, batches;
    int a_zb, a_zn;        // A tensor-map z = b*a_zb + n_tile*a_zn
    int b_zb;              // B tensor-map z = b*b_zb
    // output: row = b*M + m
    int out_mode;
    void* out;             // fp16 (OUT_H16, OUT_T_H16) or fp32 (OUT_F32)
    int ldc;               // pitch of `out`, in elements
    int out_col_stride;    // first output column of a tile = n_tile*out_col_stride
    int n_valid;           // valid accumulator columns inside one tile
    // epilogue operands (nullable)
    const float* bias;     // accumulator-column indexed, at least n_out (LINEAR) / tiles*256 (GATED)
    const float* rowscale; // fp32 [(b*rs_zb + n_tile*rs_zn)*M + m]: accumulator row scale (softmax 1/l)
    int rs_zb, rs_zn;
    const float* rowbias;  // fp32 [(row / rows_per_group)*ld_rowbias + outcol]
    int rows_per_group;
    int ld_rowbias;
    int tma_store;         // outputs leave through TMA stores (tensor map tma_out)
    int aux_mode;          // tile operand streamed by TMA: fp16 residual (added) or fp16 multiplier
    int aux_zb;            // aux tensor-map z = b*aux_zb
    int orig_tma;          // bypass: orig + (v - orig)*scale[col]; `orig` (fp16) rides through the aux ring
    const float* bypass_scale;
    int act;
    int act_cols;          // > 0: the activation applies to output columns < act_cols only (merged projections)
    int gate_mode;
    const uint8_t* row_mask;        // [rows] non-zero -> output row is zero
    // OUT_T_H16: dst[(row / t_L)*t_batch_rows + drow(col)][row % t_L], pitch t_pitch,
    // drow(col) = col + (col / t_hd)*(t_hp - t_hd)
    int t_L, t_pitch, t_batch_rows, t_hd, t_hp;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_SWOOSH_L) return swoosh_l(v);
    if (act == ACT_SWOOSH_R) return swoosh_r(v);
    if (act == ACT_GELU) return gelu_erf(v);
    return v;
}

// Direct per-thread store of 32 consecutive output columns of one row: used for the transposed
// layout (lanes = consecutive rows -> coalesced) and for ragged / unaligned column ranges.
__device__ __forceinline__ void store_row32_direct(const GemmParams& p, long long row, int col0, int ncols,
                                                   const float* v) {
    if (p.out_mode == OUT_T_H16) {
        const int n = static_cast<int>(row / p.t_L);
        const int l = static_cast<int>(row - static_cast<long long>(n) * p.t_L);
        __half* dst = reinterpret_cast<__half*>(p.out) +
                             static_cast<long long>(n) * p.t_batch_rows * p.t_pitch + l;
        if (p.t_hp == p.t_hd) {
            dst += static_cast<long long>(col0) * p.t_pitch;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < ncols) dst[static_cast<long long>(i) * p.t_pitch] = f2h(v[i]);
        } else {
            int hh = col0 / p.t_hd;                   // head of the first column, then incremental
            int rem = col0 - hh * p.t_hd;
            const int pad = p.t_hp - p.t_hd;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (i < ncols) dst[static_cast<long long>(col0 + i + hh * pad) * p.t_pitch] = f2h(v[i]);
                if (++rem == p.t_hd) { rem = 0; ++hh; }
            }
        }
        return;
    }
    if (p.out_mode == OUT_F32) {
        float* dst = reinterpret_cast<float*>(p.out) + row * p.ldc + col0;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols) dst[i] = v[i];
        return;
    }
    __half* dst = reinterpret_cast<__half*>(p.out) + row * p.ldc + col0;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < ncols) dst[i] = f2h(v[i]);
}

// fp16 staging of one 32-column unit into 16-byte chunks cofs..cofs+3 of the thread's 128-byte row
__device__ __forceinline__ void stage_h16_unit(uint8_t* stage, int lane, const float* v, int cofs) {
    uint8_t* my = stage + lane * 128;
    const int sw = lane & 7;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(my + (((cofs + j) ^ sw) << 4)) =
            make_uint4(pack_h2(v[8 * j], v[8 * j + 1]), pack_h2(v[8 * j + 2], v[8 * j + 3]),
                       pack_h2(v[8 * j + 4], v[8 * j + 5]), pack_h2(v[8 * j + 6], v[8 * j + 7]));
}
// Writes the staged fp16 columns [c_lo, c_hi) (relative to `colbase`, multiples of 8, within 0..64) of
// the warp's 32 rows: 8 lanes x 16 B = one 128-byte row segment per row, 4 rows per instruction
// (the store path is bound by the number of <=128-byte write transactions, not by bytes).
__device__ __forceinline__ void flush_h16_units(__half* out, int ldc, const uint8_t* stage, int lane,
                                                 long long row0, int rows_ok, int colbase, int c_lo, int c_hi) {
    __half* dst = out + row0 * ldc + colbase;
    const int ch = lane & 7;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int rr = 4 * k + (lane >> 3);
        const uint4 q = *reinterpret_cast<const uint4*>(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
        if (rr < rows_ok && ch * 8 >= c_lo && ch * 8 < c_hi)
            *reinterpret_cast<uint4*>(dst + static_cast<long long>(rr) * ldc + ch * 8) = q;
    }
}

// Coalesced store of a warp's 32-row x 32-column unit: TMEM hands every thread one ROW, which as a
// direct global store touches 32 different 128-byte lines per instruction (measured ~6 GB/s per SM).
// The unit is therefore transposed through 4 KB of (swizzled, conflict-free) shared memory --
// `stage`, 32 rows x 128 B -- so that each store instruction writes whole rows: 8 lanes x 16 B per
// fp32 row, 4 lanes x 16 B per fp16 row.  `row0` = global row of the warp's first row, `rows_ok` =
// number of valid rows among the 32, ncols a multiple of 4 (fp32) / 8 (fp16).  `cofs` = first 16-byte
// chunk of the 128-byte row used for the fp16 staging.
__device__ __forceinline__ void store_unit_staged(const GemmParams& p, uint8_t* stage, int lane, long long row0,
                                                  int rows_ok, int col0, int ncols, const float* v, int cofs) {
    const int sw = lane & 7;
    if (p.out_mode == OUT_F32) {
        uint8_t* my = stage + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(my + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        float* dst = reinterpret_cast<float*>(p.out) + row0 * p.ldc + col0;
        const int ch = lane & 7;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int rr = 4 * k + (lane >> 3);
            const float4 q = *reinterpret_cast<const float4*>(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
            if (rr < rows_ok && ch * 4 < ncols) *reinterpret_cast<float4*>(dst + static_cast<long long>(rr) * p.ldc + ch * 4) = q;
        }
        return;
    }
    stage_h16_unit(stage, lane, v, cofs);
    __syncwarp();
    flush_h16_units(reinterpret_cast<__half*>(p.out), p.ldc, stage, lane, row0, rows_ok, col0 - 8 * cofs, 8 * cofs,
                    8 * cofs + ncols);
}

// One warp stores its unit: staged+coalesced when the column range is 16-byte aligned, direct otherwise.
__device__ __forceinline__ void store_unit(const GemmParams& p, uint8_t* stage, int lane, bool row_ok, long long row,
                                           long long row0, int rows_ok, int col0, int ncols, const float* v, int cofs) {
    const bool f32 = p.out_mode == OUT_F32;
    const bool aligned = p.out_mode != OUT_T_H16 && (p.ldc & 7) == 0 && (col0 & 7) == 0 &&
                         (ncols & (p.out_mode == OUT_F32 ? 3 : 7)) == 0 && (!f32 || (ncols & 3) == 0);
    if (aligned) {
        store_unit_staged(p, stage, lane, row0, rows_ok, col0, ncols, v, cofs);
        __syncwarp();
    } else if (row_ok) {
        store_row32_direct(p, row, col0, ncols, v);
    }
}


// One 32-column unit of a gated projection: v = (a + ba) gate (g + bg), bias from the warp's staging (bs[0..31] first
// operand, bs[32..63] second).  TANH: x * tanh(s) with a = s, g = x (NonlinAttention); else GLU x * sigmoid(s) with a = x,
// g = s (ConvolutionModule).  Straight-line over the 16 column pairs.
template <bool TANH>
__device__ __forceinline__ void gated_unit(const uint32_t* ra, const uint32_t* rb, const float* bs, bool masked, float* v) {
    constexpr float NL2E = -1.4426950408889634f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {           // two columns per packed fp32x2 instruction
        const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * j);
        const float4 g4 = *reinterpret_cast<const float4*>(bs + 32 + 4 * j);
        const float abv[4] = {b4.x, b4.y, b4.z, b4.w};
        const float gbv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            const f32x2 a2 = add2(pack2(__uint_as_float(ra[4 * j + e]), __uint_as_float(ra[4 * j + e + 1])),
                                  pack2(abv[e], abv[e + 1]));
            const f32x2 g2 = add2(pack2(__uint_as_float(rb[4 * j + e]), __uint_as_float(rb[4 * j + e + 1])),
                                  pack2(gbv[e], gbv[e + 1]));
            f32x2 o2;
            if (TANH) {
                float a0, a1;
                unpack2(a2, a0, a1);
                o2 = mul2(g2, pack2(fast_tanh(a0), fast_tanh(a1)));
            } else {
                float z0, z1;
                unpack2(mul2(g2, pack2(NL2E, NL2E)), z0, z1);
                const f32x2 d2 = add2(pack2(fast_exp2(z0), fast_exp2(z1)), pack2(1.0f, 1.0f));
                float d0, d1;
                unpack2(d2, d0, d1);
                o2 = mul2(a2, pack2(fast_rcp(d0), fast_rcp(d1)));
            }
            float o0, o1;
            unpack2(o2, o0, o1);
            v[4 * j + e] = masked ? 0.0f : o0;
            v[4 * j + e + 1] = masked ? 0.0f : o1;
        }
    }
}

// CLUSTER == 2: a CTA pair computes a 256 x BN tile with `tcgen05.mma.cta_group::2` (UMMA_M = 256):
// each CTA stages its own 128 rows of A and HALF of the B tile, the leader's MMA thread consumes both
// CTAs' shared memory, and each CTA's TMEM receives its 128 accumulator rows.  Per SM and k-block
// this moves 16 KB + BN/2 x 128 B from L2 instead of 16 KB + BN x 128 B -- the K <= 1920 GEMMs of this
// network are bound by that L2->SM operand traffic (~41 B/clk/SM measured), not by the tensor pipe.
// (A 2-CTA cluster with a multicast B tile was measured first and saves nothing: L2 already
// de-duplicates the two requests.)  Protocol: both producers wait on their LOCAL `empty` barrier and
// complete bytes on the LEADER's `full` barrier; the leader commits with a multicast arrive to both
// CTAs' `empty` / `tmem_full`; the peer's epilogue warps arrive remotely on the leader's `tmem_empty`.
// LEAN selects the epilogue at COMPILE time: 0 = generic (every operand / store mode), 1 = lean plain epilogue
// (GemmParams::fast_epi), 2 = lean residual-stream epilogue (GemmParams::fast_resid).  Separate instantiations, because
// 3 = lean residual epilogue with the bypass (fast_resid == 2); one kernel carrying all paths costs the generic path registers (round 2: +20% on the N = 272 / N = 48 projections
// when the lean paths were runtime branches of the same kernel).
#ifdef ZVB_TIMELINE
// Debug build only (tools/timeline_c1.py): CTA 0 of every GEMM launch stamps clock64 at the points of its critical path.
constexpr int TL_MAX = 1 << 15;
__device__ unsigned long long g_tl[TL_MAX][20];
__device__ unsigned int g_tl_n;
__device__ volatile unsigned long long g_tl_setup[4];       // set-up stamps of the CTA 0 in flight (one launch at a time writes them)
#define TL_STAMP(k) do { if (blockIdx.x == 0 && tl_slot < TL_MAX) g_tl[tl_slot][k] = clock64(); } while (0)
#else
#define TL_STAMP(k) do { } while (0)
#endif
template <int KIND, int ACT, int CLUSTER, int LEAN = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_aux, const __grid_constant__ CUtensorMap tma_out,
            const __grid_constant__ CUtensorMap tma_orig,
            const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
#ifdef ZVB_TIMELINE
    const unsigned long long tl_c0 = clock64(), tl_g0 = globaltimer_ns();
#endif
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int STAGES = p.stages;
    const int STAGE_BYTES = p.stage_bytes;
    uint8_t* aux_smem = smem + p.ring_bytes;
    float* bias_smem = reinterpret_cast<float*>(aux_smem + p.aux_slots * GEMM_AUX_BYTES);      // [16 warps][64]
    const int AUX_SLOTS = p.aux_slots;
    const uint32_t dsh = static_cast<uint32_t>(p.stage_depth - 1);     // staging buffer of sub-tile k: k & dsh
    float* cbias_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bias_smem) + GEMM_BIAS_BYTES);   // [2][2][256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(cbias_smem) + GEMM_CBIAS_BYTES);
    if (threadIdx.x == 0 && ((smem - smem_raw) + GEMM_LAYOUT_BYTES > GEMM_SMEM_BYTES ||
                             p.ring_bytes + p.aux_slots * GEMM_AUX_BYTES > GEMM_SHARED_BUDGET)) {
        printf("zvb: gemm shared-memory layout does not fit (base misaligned by %d)\n", (int)(smem - smem_raw));
        __trap();
    }
    uint64_t* full_bar = bars;                                  // [STAGES] TMA -> MMA
    uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;           // [STAGES] MMA -> TMA
    uint64_t* tmem_full = empty_bar + GEMM_MAX_STAGES;          // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;                       // [2] epilogue -> MMA
    uint64_t* aux_full = tmem_empty + 2;                        // [AUX_SLOTS] TMA -> epilogue
    uint64_t* aux_empty = aux_full + GEMM_AUX_SLOTS_MAX;        // [AUX_SLOTS] epilogue -> TMA
    uint64_t* staged = aux_empty + GEMM_AUX_SLOTS_MAX;          // [2 halves][2] epilogue -> store thread
    uint64_t* sfree = staged + 4;                               // [2 halves][2] store thread -> epilogue
    uint64_t* a_full = sfree + 4;                               // [8] resident A k-block landed (TMA -> MMA)
    uint64_t* a_empty = a_full + 8;                             // [8] resident A k-block released (MMA -> TMA)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(a_empty + 8);
#ifdef ZVB_TIMELINE
    volatile unsigned int* tl_slot_p = tmem_holder + 1;         // inside the 512 bytes reserved for the barriers
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned int sl = atomicAdd(&g_tl_n, 1u);
        *tl_slot_p = sl;
        if (sl < TL_MAX) {
            g_tl[sl][0] = tl_c0;
            g_tl[sl][8] = tl_g0;
            g_tl[sl][10] = (static_cast<unsigned long long>(p.M) << 32) | static_cast<unsigned>(p.n_out);
            g_tl[sl][11] = (static_cast<unsigned long long>(p.num_k_blocks) << 32) | (static_cast<unsigned>(p.block_n) << 8) |
                           (static_cast<unsigned>(CLUSTER) << 4) | static_cast<unsigned>(LEAN);
        }
    }
#endif

    // Roles: warps 0..15 = epilogue, 16 = TMA producer (A/B), 17 = MMA issuer (+TMEM alloc), 18 = TMA producer
    // (aux), 19 = TMA store thread.  The single-thread roles sit in the HIGHEST warp ids because the
    // SMSP arbiter favours higher warp ids: an MMA issue or TMA request then never queues behind the
    // FFMA streams of the epilogue warps sharing its scheduler.
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int W_TMA = GEMM_EPI_WARPS, W_MMA = W_TMA + 1, W_AUX = W_TMA + 2, W_STORE = W_TMA + 3;
    // tile schedule: a "slot" is one tile per CTA of the cluster; slot s -> (b, m_group, n_tile) and the
    // CTA of rank `crank` takes m_tile = m_group*CLUSTER + crank (a tile past num_m_tiles is all padding)
    const int crank = CLUSTER > 1 ? static_cast<int>(cluster_ctarank()) : 0;
    const int m_groups = (p.num_m_tiles + CLUSTER - 1) / CLUSTER;
    const int total_tiles = p.batches * m_groups * p.num_n_tiles;
    const int first_tile = blockIdx.x / CLUSTER;
    const int tile_step = gridDim.x / CLUSTER;
    // tile range of this CTA (pair): strided over the grid, or -- A-stationary -- one contiguous range so that the
    // n-tiles of an m-group follow each other and its A tile is loaded once
    const bool resident = CLUSTER == 2 && p.a_resident != 0;
    const int t_first = resident ? static_cast<int>(static_cast<long long>(total_tiles) * first_tile / tile_step) : first_tile;
    const int t_end = resident ? static_cast<int>(static_cast<long long>(total_tiles) * (first_tile + 1) / tile_step) : total_tiles;
    const int t_step = resident ? 1 : tile_step;
    uint8_t* ring = smem + (resident ? p.a_bytes : 0);          // operand ring ([A | B] stages, or B-only stages)
    // accumulator columns are consumed in units of 32; a 128-byte sub-tile row holds 64 fp16 (2 units) or
    // 32 fp32 (1 unit) columns
    const int n_units = (p.block_n + 31) >> 5;
    const bool out_f32 = p.out_mode == OUT_F32;
    const int units_per_sub = ((p.tma_store && out_f32) || p.out_mode == OUT_T_H16) ? 1 : 2;
    const int aux_parts = p.orig_tma ? 2 : 1;          // ring entries per sub-tile: operand (+ bypass `orig`)
    const int n_sub = (n_units + units_per_sub - 1) / units_per_sub;

    if (warp == W_TMA) {
        // the 60 barriers are initialised by the 32 lanes of the producer warp (set-up is ~0.8 us from kernel entry to the block
        // barrier and +0.8 us for a pair's cluster barrier, tools/timeline_c1.py; one thread initialising them all measured the
        // same sample time -- the set-up is dominated by the launch's cold start and the TMEM allocation, not by this loop)
        if (lane == 0) {
            tma_prefetch_desc(&tma_a);
            tma_prefetch_desc(&tma_b);
            if (p.aux_mode != AUX_NONE) tma_prefetch_desc(&tma_aux);
            if (p.tma_store) tma_prefetch_desc(&tma_out);
            if (p.orig_tma) tma_prefetch_desc(&tma_orig);
        }
        constexpr int I_TMEM_FULL = 2 * GEMM_MAX_STAGES, I_TMEM_EMPTY = I_TMEM_FULL + 2, I_AUX_FULL = I_TMEM_EMPTY + 2,
                      I_AUX_EMPTY = I_AUX_FULL + GEMM_AUX_SLOTS_MAX, I_STAGED = I_AUX_EMPTY + GEMM_AUX_SLOTS_MAX,
                      I_SFREE = I_STAGED + 4, I_END = I_SFREE + 4 + 16;
        for (int i = lane; i < I_END; i += 32) {
            uint32_t cnt = 1;
            if (i >= I_TMEM_EMPTY && i < I_AUX_FULL) cnt = GEMM_EPI_WARPS * CLUSTER;
            else if (i >= I_AUX_EMPTY && i < I_STAGED) cnt = p.tma_store ? 1 : (units_per_sub == 2 ? 8 : 4);
            else if (i >= I_STAGED && i < I_SFREE) cnt = units_per_sub == 2 ? 8 : 4;
            mbar_init(&bars[i], cnt);
        }
        fence_barrier_init();
#ifdef ZVB_TIMELINE
        if (blockIdx.x == 0) g_tl_setup[0] = clock64();
#endif
    }
    if (warp == W_MMA) {
        if (CLUSTER == 2) { tmem_alloc_2sm(tmem_holder, GEMM_TMEM_COLS); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_holder, GEMM_TMEM_COLS); tmem_relinquish(); }
#ifdef ZVB_TIMELINE
        if (blockIdx.x == 0 && lane == 0) g_tl_setup[1] = clock64();
#endif
    }
    tc_fence_before();
    __syncthreads();
#ifdef ZVB_TIMELINE
    if (blockIdx.x == 0 && threadIdx.x == 0) g_tl_setup[2] = clock64();
#endif
    if (CLUSTER > 1) cluster_sync_all();        // peers' barriers are initialised before any remote arrive
#ifdef ZVB_TIMELINE
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned int sl = *tl_slot_p;
        if (sl < TL_MAX) { g_tl[sl][16] = g_tl_setup[0]; g_tl[sl][17] = g_tl_setup[1]; g_tl[sl][18] = g_tl_setup[2]; g_tl[sl][19] = clock64(); }
    }
#endif
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
#ifdef ZVB_TIMELINE
    const unsigned int tl_slot = blockIdx.x == 0 ? *tl_slot_p : TL_MAX;
    if (threadIdx.x == 0) TL_STAMP(1);
#endif
    // Weights do not depend on the predecessor: the B halves of the first tile's stages are requested before the dependency
    // wait, so that only the activations' L2 round trip remains after it (single-utterance critical path, tools/timeline_c1.py).
    int npre = 0;
    if (warp == W_TMA && lane == 0 && p.pre_b != 0 && !resident && t_first < t_end) {
        const uint32_t stage_tx = CLUSTER * GEMM_A_BYTES + static_cast<uint32_t>(p.block_n) * GEMM_BLOCK_K * 2;
        const int b_rows = p.block_n / CLUSTER;
        const int n_tile = t_first % p.num_n_tiles;
        const int rest = t_first / p.num_n_tiles;
        const int b = rest / m_groups;
        const int bz = b * p.b_zb;
        const int az = b * p.a_zb + n_tile * p.a_zn;
        const int m_tile = (rest % m_groups) * CLUSTER + crank;
        npre = STAGES < p.num_k_blocks ? STAGES : p.num_k_blocks;
        for (int kb = 0; kb < npre; ++kb) {
            uint8_t* sa = ring + kb * STAGE_BYTES;
            if (CLUSTER == 2) {
                if (crank == 0) mbar_arrive_expect_tx(&full_bar[kb], stage_tx);
                if (p.pre_b == 2) tma_load_3d_2sm(sa, &tma_a, &full_bar[kb], kb * GEMM_BLOCK_K, m_tile * GEMM_BLOCK_M, az);
                else tma_load_3d_2sm(sa + GEMM_A_BYTES, &tma_b, &full_bar[kb], kb * GEMM_BLOCK_K, n_tile * p.block_n + crank * b_rows, bz);
            } else {
                mbar_arrive_expect_tx(&full_bar[kb], stage_tx);
                if (p.pre_b == 2) tma_load_3d(sa, &tma_a, &full_bar[kb], kb * GEMM_BLOCK_K, m_tile * GEMM_BLOCK_M, az);
                else tma_load_3d(sa + GEMM_A_BYTES, &tma_b, &full_bar[kb], kb * GEMM_BLOCK_K, n_tile * p.block_n, bz);
            }
        }
    }
    pdl_wait();                 // set-up above overlaps the previous kernel's tail
    pdl_launch();
    if (threadIdx.x == 0) TL_STAMP(2);

    if (warp == W_TMA) {
        // ------------------------------------------------------------------ TMA producer (A, B)
        if (lane == 0) {
            // bytes landing on the (leader's) full barrier per stage: both CTAs' A tiles + the whole B tile
            const uint32_t stage_tx = CLUSTER * GEMM_A_BYTES + static_cast<uint32_t>(p.block_n) * GEMM_BLOCK_K * 2;
            const int b_rows = p.block_n / CLUSTER;                    // B rows staged by this CTA
            int stage = 0;
            uint32_t phase = 0;
            uint32_t mg = 0;                        // m-groups started so far (A-stationary mode)
            for (int tile = t_first; tile < t_end; tile += t_step) {
                const int n_tile = tile % p.num_n_tiles;
                const int rest = tile / p.num_n_tiles;
                const int m_tile = (rest % m_groups) * CLUSTER + crank;
                const int b = rest / m_groups;
                const int az = b * p.a_zb + n_tile * p.a_zn;
                const int bz = b * p.b_zb;
                if (CLUSTER == 2 && resident) {
                    const bool new_mg = tile == t_first || n_tile == 0;
                    const uint32_t b_tx = static_cast<uint32_t>(p.block_n) * GEMM_BLOCK_K * 2;
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        if (new_mg) {       // this k-block of the previous m-group's A has been consumed by its last n-tile
                            mbar_wait(&a_empty[kb], (mg & 1u) ^ 1u);
                            if (crank == 0) mbar_arrive_expect_tx(&a_full[kb], 2 * GEMM_A_BYTES);
                            tma_load_3d_2sm(smem + kb * GEMM_A_BYTES, &tma_a, &a_full[kb], kb * GEMM_BLOCK_K,
                                            m_tile * GEMM_BLOCK_M, az);
                        }
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], b_tx);
                        tma_load_3d_2sm(ring + stage * STAGE_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BLOCK_K,
                                        n_tile * p.block_n + crank * b_rows, bz);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    if (tile == t_end - 1 || n_tile == p.num_n_tiles - 1) ++mg;
                    continue;
                }
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    uint8_t* sa = ring + stage * STAGE_BYTES;
                    if (tile == t_first && kb < npre) {        // barrier armed and one operand requested before the dependency wait
                        if (p.pre_b == 2) {
                            if (CLUSTER == 2)
                                tma_load_3d_2sm(sa + GEMM_A_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BLOCK_K,
                                                n_tile * p.block_n + crank * b_rows, bz);
                            else
                                tma_load_3d(sa + GEMM_A_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BLOCK_K, n_tile * p.block_n, bz);
                        } else if (CLUSTER == 2) {
                            tma_load_3d_2sm(sa, &tma_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_tile * GEMM_BLOCK_M, az);
                        } else {
                            tma_load_3d(sa, &tma_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_tile * GEMM_BLOCK_M, az);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (CLUSTER == 2) {
                        if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                        tma_load_3d_2sm(sa, &tma_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_tile * GEMM_BLOCK_M, az);
                        tma_load_3d_2sm(sa + GEMM_A_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BLOCK_K,
                                        n_tile * p.block_n + crank * b_rows, bz);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                        tma_load_3d(sa, &tma_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_tile * GEMM_BLOCK_M, az);
                        tma_load_3d(sa + GEMM_A_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BLOCK_K,
                                    n_tile * p.block_n, bz);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0 && crank == 0) {               // the leader issues for the pair
            const uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(p.block_n), 128u * CLUSTER);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            uint32_t mg = 0;
            for (int tile = t_first; tile < t_end; tile += t_step) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc) * 256u;
                const int n_tile = tile % p.num_n_tiles;
                const bool new_mg = resident && (tile == t_first || n_tile == 0);
                const bool last_of_mg = resident && (tile == t_end - 1 || n_tile == p.num_n_tiles - 1);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    if (new_mg) mbar_wait(&a_full[kb], mg & 1u);
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (tile == t_first && kb == 0) TL_STAMP(3);
                    const uint32_t sa = smem_u32(ring + stage * STAGE_BYTES);
                    const uint64_t da = umma_desc_k_sw128(resident ? smem_u32(smem + kb * GEMM_A_BYTES) : sa);
                    const uint64_t db = umma_desc_k_sw128(resident ? sa : sa + GEMM_A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
                        // advance 16 fp16 = 32 bytes inside the 128B swizzle atom: +2 in >>4 units
                        if (CLUSTER == 2)
                            umma_f16_2sm(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                                          idesc, (kb | k) != 0 ? 1u : 0u);
                        else
                            umma_f16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                                      idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // free the smem slot (in both CTAs) when the MMAs retire; publish the accumulator
                    if (CLUSTER == 2) {
                        umma_commit_2sm(&empty_bar[stage], 0x3);
                        if (last_of_mg) umma_commit_2sm(&a_empty[kb], 0x3);
                        if (kb == p.num_k_blocks - 1) umma_commit_2sm(&tmem_full[acc], 0x3);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[acc]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (last_of_mg) ++mg;
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                if (tile == t_first) TL_STAMP(4);
            }
        }
    } else if (warp == W_AUX) {
        // ------------------------------------------------------------------ TMA producer (aux tiles)
        if (lane == 0 && p.aux_mode != AUX_NONE) {
            constexpr int sub_cols = 64;
            uint32_t q = 0;
            for (int tile = t_first; tile < t_end; tile += t_step) {
                const int n_tile = tile % p.num_n_tiles;
                const int rest = tile / p.num_n_tiles;
                const int m_tile = (rest % m_groups) * CLUSTER + crank;
                const int b = rest / m_groups;
                for (int s = 0; s < n_sub; ++s) {
                    for (int part = 0; part < aux_parts; ++part, ++q) {
                        const int slot = q % AUX_SLOTS;
                        const uint32_t par = (q / AUX_SLOTS) & 1u;
                        mbar_wait(&aux_empty[slot], par ^ 1u);
                        mbar_arrive_expect_tx(&aux_full[slot], GEMM_AUX_BYTES);
                        tma_load_3d(aux_smem + slot * GEMM_AUX_BYTES, part == 0 ? &tma_aux : &tma_orig, &aux_full[slot],
                                    n_tile * p.out_col_stride + s * sub_cols, m_tile * GEMM_BLOCK_M, b * p.aux_zb);
                    }
                }
            }
        }
    } else if (warp == W_STORE) {
        // ------------------------------------------------------------------ TMA store thread
        // Epilogue halves (8 warps each) stage 128-row x 128-byte sub-tiles and signal `staged`; this
        // thread issues the box stores (16 KB each: few, large TMA operations -- per-warp 4 KB stores
        // were measured to delay the operand loads) and hands a buffer back (`sfree`, and the aux slot
        // that was staged in place) once its store has read it.  Up to ST_PEND stores stay in flight: a
        // buffer is only reclaimed when a later store has been issued, or when no staged sub-tile is
        // waiting (then everything outstanding is drained first -- the epilogue may need those buffers).
        if (lane == 0 && p.tma_store) {
            const int subs = KIND == EPI_GATED ? 2 : n_sub;
            uint32_t kc[2] = {0u, 0u};
            uint32_t tile_iter = 0;
            constexpr uint32_t ST_PEND = 2;
            int pend_bi[4] = {0, 0, 0, 0}, pend_slot[4] = {0, 0, 0, 0};
            uint32_t issued = 0, freed = 0;
            auto release_oldest = [&]() {
                const int bi0 = pend_bi[freed & 3u], sl0 = pend_slot[freed & 3u];
                mbar_arrive(&sfree[bi0]);
                if (p.aux_mode != AUX_NONE) mbar_arrive(&aux_empty[sl0]);
                ++freed;
            };
            for (int tile = t_first; tile < t_end; tile += t_step, ++tile_iter) {
                const int n_tile = tile % p.num_n_tiles;
                const int rest = tile / p.num_n_tiles;
                const int m_tile = (rest % m_groups) * CLUSTER + crank;
                const int b = rest / m_groups;
                const int out_base = n_tile * p.out_col_stride;
                for (int s = 0; s < subs; ++s) {
                    const int h = s & 1;
                    const uint32_t k = kc[h]++;
                    const int bi = h * 2 + static_cast<int>(k & dsh);
                    int slot = 0, slot2 = 0;
                    const uint8_t* src = aux_smem + ((k & dsh) * 2 + h) * GEMM_AUX_BYTES;
                    int col = out_base + 64 * h;
                    if (KIND != EPI_GATED) {
                        col = out_base + s * units_per_sub * 32;
                        if (p.aux_mode != AUX_NONE) {
                            const uint32_t q = (tile_iter * static_cast<uint32_t>(n_sub) + static_cast<uint32_t>(s)) *
                                               static_cast<uint32_t>(aux_parts);
                            slot = q % AUX_SLOTS;
                            slot2 = (q + 1) % AUX_SLOTS;
                            src = aux_smem + slot * GEMM_AUX_BYTES;
                        }
                    }
                    if (!mbar_test(&staged[bi], (k >> dsh) & 1u)) {
                        bulk_wait_read<0>();                                 // nothing to issue: drain
                        while (freed < issued) release_oldest();
                        mbar_wait(&staged[bi], (k >> dsh) & 1u);
                    }
                    if (aux_parts == 2) mbar_arrive(&aux_empty[slot2]);      // `orig` rows are consumed
                    tma_store_3d(&tma_out, src, col, m_tile * GEMM_BLOCK_M, b);
                    bulk_commit();
                    pend_bi[issued & 3u] = bi; pend_slot[issued & 3u] = slot;
                    ++issued;
                    if (issued - freed > ST_PEND) {
                        bulk_wait_read<ST_PEND>();
                        release_oldest();
                    }
                }
            }
            bulk_wait_read<0>();
            while (freed < issued) release_oldest();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (16 warps)
        // Four warps share a TMEM lane quarter (32 accumulator rows): `half` selects every other 128-byte
        // wide sub-tile, `part` the 32-column unit inside a 64-column fp16 sub-tile, so the two parts fill
        // the two halves of the same staging rows.  The tile-shaped math (bias, activation, residual,
        // bypass, conversion) is latency bound per warp; four warps per scheduler keep the FMA/MUFU pipes
        // busy (8 warps: 5.8 activations/clk/SM, 32 warps: 9.6, tools/microbench/epi_math.cu).
        // Results are staged into the half's 128-row box -- in place over the consumed aux sub-tile, or in
        // the half's two-buffer ring.  No block-level barrier: one lane per warp arrives on `staged`, the
        // store thread issues the box store and returns the buffer through `sfree`.
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
        const int ew = warp;                              // 0..15
        const int half = (ew >> 2) & 1;
        const int part = ew >> 3;
        const int r = quarter * 32 + lane;                // accumulator row inside the tile
        uint8_t* private_stage = aux_smem + ew * 4096;    // staging of the non-TMA store paths
        uint32_t kcount = 0;                              // sub-tiles this half handed to the store thread
        // a staging buffer may be rewritten once the store of `depth` sub-tiles ago has read it (one or
        // two buffers per half, or the consumed aux slot)
        const uint32_t depth = dsh + 1u;
        auto wait_sfree = [&]() {
            if (kcount >= depth) {
                const uint32_t kd = kcount - depth;
                mbar_wait(&sfree[half * 2 + (kd & dsh)], (kd >> dsh) & 1u);
            }
        };
        auto signal_staged = [&]() {
            fence_proxy_async_smem();                     // generic-proxy writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) mbar_arrive(&staged[half * 2 + (kcount & dsh)]);
            ++kcount;
        };
        float* bs = bias_smem + ew * 64;                  // this warp's bias staging: 32 floats (2 x 32 gated)
        // LINEAR: which sub-tiles / unit this warp takes.  64-column sub-tiles: both parts work on the same
        // sub-tile; 32-column sub-tiles: with TMA stores the staging ring is per half (part 1 idles, only
        // the small fp32 outputs take this path), otherwise the four groups alternate.
        const int s_step = units_per_sub == 2 ? 2 : (p.tma_store ? 2 : 4);
        const int s_first = units_per_sub == 2 ? half : (p.tma_store ? half : half + 2 * part);
        const bool lin_active = units_per_sub == 2 || !p.tma_store || part == 0;
        const int uu = units_per_sub == 2 ? part : 0;     // unit inside the sub-tile
        // lean path (fast_epi): 32-bit shared-window addresses, computed once
        const uint32_t fast_stage = smem_u32(aux_smem) + static_cast<uint32_t>(half * GEMM_AUX_BYTES + quarter * 4096 + lane * 128);
        const uint32_t fast_staged = smem_u32(&staged[half * 2]);
        const uint32_t fast_sfree = smem_u32(&sfree[half * 2]);
        uint32_t fast_chunk[4];                           // swizzled 16-byte chunks of this warp's unit inside a staging row
#pragma unroll
        for (int j = 0; j < 4; ++j) fast_chunk[j] = static_cast<uint32_t>(((4 * part + j) ^ (lane & 7)) << 4);
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t tile_iter = 0;
        for (int tile = t_first; tile < t_end; tile += t_step, ++tile_iter) {
            const int n_tile = tile % p.num_n_tiles;
            const int rest = tile / p.num_n_tiles;
            const int m_tile = (rest % m_groups) * CLUSTER + crank;
            const int b = rest / m_groups;
            const int m = m_tile * GEMM_BLOCK_M + r;
            const bool row_ok = m < p.M;
            const long long row = static_cast<long long>(b) * p.M + m;
            const int m0w = m_tile * GEMM_BLOCK_M + quarter * 32;          // first row of this warp
            const long long row0 = static_cast<long long>(b) * p.M + m0w;
            int rows_ok = p.M - m0w;
            rows_ok = rows_ok < 0 ? 0 : (rows_ok > 32 ? 32 : rows_ok);
            const int out_base = n_tile * p.out_col_stride;
            const int acc_base = n_tile * p.block_n;
            const int bias_lim = KIND == EPI_GATED ? p.num_n_tiles * 256 : p.n_out;
            // bias of accumulator column cc of this tile (zero outside the valid range)
            auto bias_at = [&](int cc) -> float {
                const int c = acc_base + cc;
                return (p.bias != nullptr && cc < p.block_n && c < bias_lim) ? __ldg(p.bias + c) : 0.0f;
            };
            // requested before the accumulator wait: the bias values of this warp's first unit
            float nb0, nb1 = 0.0f;
            if (KIND == EPI_GATED) {
                const int u = 2 * half + part;
                nb0 = bias_at(32 * u + lane);
                nb1 = bias_at((p.block_n >> 1) + 32 * u + lane);
            } else {
                nb0 = bias_at((s_first * units_per_sub + uu) * 32 + lane);
            }
            bool masked = false;
            if (p.row_mask != nullptr && row_ok) masked = p.row_mask[row] != 0;
            long long grp = 0;
            if (p.rowbias != nullptr && row_ok) grp = row / p.rows_per_group;
            float rscale = 1.0f;
            if (p.rowscale != nullptr && row_ok)
                rscale = __ldg(p.rowscale + static_cast<long long>(b * p.rs_zb + n_tile * p.rs_zn) * p.M + m);

            // lean paths: the 512 epilogue threads fill this tile's bias cache (one L2 round trip per tile, requested
            // before the accumulator wait), double buffered by tile parity; one named barrier per tile
            uint32_t cb_addr = 0;
            if (KIND == EPI_LINEAR && LEAN != 0) {
                const int e = ew * 32 + lane;                     // 0..511
                const int ccol = acc_base + (e & 255);
                const long long row_first = static_cast<long long>(b) * p.M + m_tile * GEMM_BLOCK_M;
                long long g0 = 0, g_last = 0;
                if (p.rowbias != nullptr) {
                    g0 = row_first / p.rows_per_group;
                    g_last = (static_cast<long long>(b) * p.M + p.M - 1) / p.rows_per_group;
                }
                float cv = (p.bias != nullptr && ccol < p.n_out) ? __ldg(p.bias + ccol) : 0.0f;
                if (p.rowbias != nullptr && ccol < p.n_out) {
                    long long g = g0 + (e >> 8);
                    g = g > g_last ? g_last : g;
                    cv += __ldg(p.rowbias + g * p.ld_rowbias + ccol);
                }
                if (LEAN == 3 && (e >> 8) != 0) cv = ccol < p.n_out ? __ldg(p.bypass_scale + ccol) : 0.0f;   // row 1: bypass scale
                float* cb = cbias_smem + (tile_iter & 1u) * 512;
                cb[e] = cv;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                // rows of the tile's second utterance (a tile spans at most two when rows_per_group >= 128) read row 1
                const uint32_t rsel = (p.rowbias != nullptr && row_ok && grp != g0) ? 1024u : 0u;
                cb_addr = smem_u32(cb) + rsel;
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            if (threadIdx.x == 0 && tile == t_first) TL_STAMP(5);
            const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc) * 256u +
                                   (static_cast<uint32_t>(quarter * 32) << 16);

            if (KIND == EPI_LINEAR && LEAN == 1) {
                // Lean path of the feed-forward input GEMMs.  These kernels are bound by the epilogue's issue slots
                // (ncu, round 2: 415 executed instructions per 32 x 32 unit of which 176 are FFMA2 / MUFU; tensor pipe
                // 51%), so everything that is not arithmetic is hoisted: shared-memory and barrier addresses are
                // 32-bit window addresses computed once per kernel, the bias is read as warp-uniform 16-byte loads
                // (no staging through shared memory, no warp barriers), no per-unit bounds beyond `live`.
                // Units whose columns lie past `act_cols` (the attention projections of a merged GEMM) take no activation.  The
                // two kinds run in two loops over the same body, not behind a branch inside one: with the branch ptxas
                // reconciled the two paths' register assignments with ~40 moves per unit ahead of it (ncu r2f: IMAD.MOV the most
                // executed opcode of the merged GEMM).
                auto unit = [&](const int s, auto with_act) {
                    constexpr bool kAct = decltype(with_act)::value;
                    const int c0 = (2 * s + part) * 32;
                    const int gc = acc_base + c0;
                    const bool live = gc < p.n_out;               // whole unit inside the output (n_out % 32 == 0)
                    const uint32_t buf = kcount & dsh;
                    if (live) {
                        uint32_t acc_r[32];
                        tmem_ld32(taddr + c0, acc_r);
                        f32x2 v2[16];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {             // bias from the tile's shared-memory cache (broadcast reads)
                            const uint4 bq = lds128_u32(cb_addr + static_cast<uint32_t>(c0 * 4 + j * 16));
                            v2[2 * j] = pack2(__uint_as_float(bq.x), __uint_as_float(bq.y));
                            v2[2 * j + 1] = pack2(__uint_as_float(bq.z), __uint_as_float(bq.w));
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            v2[j] = add2(pack2(__uint_as_float(acc_r[2 * j]), __uint_as_float(acc_r[2 * j + 1])), v2[j]);
                        if (!kAct) {
                        } else if (ACT == ACT_SWOOSH_L) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) swoosh_x2_group<4>(v2 + j, SWOOSH_L_C, SWOOSH_L_K0);
                        } else if (ACT == ACT_SWOOSH_R) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) swoosh_x2_group<4>(v2 + j, SWOOSH_R_C, SWOOSH_R_K0);
                        } else if (ACT == ACT_GELU) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float g0, g1;
                                unpack2(v2[j], g0, g1);
                                v2[j] = pack2(gelu_erf(g0), gelu_erf(g1));
                            }
                        }
                        if (kcount >= depth) {
                            const uint32_t kd = kcount - depth;
                            mbar_wait_u32(fast_sfree + 8u * (kd & dsh), (kd >> dsh) & 1u);
                        }
                        const uint32_t row_addr = fast_stage + buf * (2u * GEMM_AUX_BYTES);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float a0, a1, a2, a3, a4, a5, a6, a7;
                            unpack2(v2[4 * j], a0, a1); unpack2(v2[4 * j + 1], a2, a3);
                            unpack2(v2[4 * j + 2], a4, a5); unpack2(v2[4 * j + 3], a6, a7);
                            sts128_u32(row_addr + fast_chunk[j], pack_h2(a0, a1), pack_h2(a2, a3), pack_h2(a4, a5),
                                       pack_h2(a6, a7));
                        }
                    } else if (kcount >= depth) {
                        const uint32_t kd = kcount - depth;
                        mbar_wait_u32(fast_sfree + 8u * (kd & dsh), (kd >> dsh) & 1u);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_u32(fast_staged + 8u * buf);
                    ++kcount;
                };
                int s_act = n_sub;                                // units s < s_act of this warp carry the activation
                if (ACT != ACT_NONE && p.act_cols > 0) {
                    const int units = (p.act_cols - acc_base) / 32 - part;          // act_cols % 32 == 0
                    s_act = units <= 0 ? 0 : (units + 1) >> 1;
                    s_act = s_act < n_sub ? s_act : n_sub;
                }
                int s = half;
                for (; s < s_act; s += 2) unit(s, std::true_type());
                if (ACT != ACT_NONE)
                    for (; s < n_sub; s += 2) unit(s, std::false_type());
            } else if (KIND == EPI_LINEAR && ACT == ACT_NONE && LEAN == 2) {
                // Lean path of the residual-stream GEMMs (out = resid + A.W^T + b [+ time embedding of the row's utterance]):
                // the generic path below executed ~620 warp instructions per 32 x 32 unit for ~6 per column pair of
                // arithmetic (ncu round 2: issue slots 52-56% busy, tensor pipe 5-47%, 90 us against a 51-73 us HBM bound).
                // Same hand-shakes (aux ring in, staging in place over the consumed residual rows, store thread out), but
                // bias + row bias come from the tile's shared-memory cache, the residual as four 16-byte shared-memory reads on
                // precomputed window addresses, and there are no per-unit bounds (every sub-tile is inside the output).
                const uint32_t aux_row0 = smem_u32(aux_smem) + static_cast<uint32_t>(r * 128);
                for (int s = half; s < n_sub; s += 2) {
                    const uint32_t q = tile_iter * static_cast<uint32_t>(n_sub) + static_cast<uint32_t>(s);
                    const uint32_t slot = q % static_cast<uint32_t>(AUX_SLOTS);
                    const int c0 = (2 * s + part) * 32;
                    uint32_t acc_r[32];
                    tmem_ld32(taddr + c0, acc_r);
                    f32x2 v2[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                 // bias + row bias from the tile's shared-memory cache
                        const uint4 bq = lds128_u32(cb_addr + static_cast<uint32_t>(c0 * 4 + j * 16));
                        v2[2 * j] = pack2(__uint_as_float(bq.x), __uint_as_float(bq.y));
                        v2[2 * j + 1] = pack2(__uint_as_float(bq.z), __uint_as_float(bq.w));
                    }
                    mbar_wait(&aux_full[slot], (q / static_cast<uint32_t>(AUX_SLOTS)) & 1u);
                    const uint32_t arow = aux_row0 + slot * static_cast<uint32_t>(GEMM_AUX_BYTES);
                    uint4 a4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) a4[j] = lds128_u32(arow + fast_chunk[j]);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v2[j] = add2(pack2(__uint_as_float(acc_r[2 * j]), __uint_as_float(acc_r[2 * j + 1])), v2[j]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v2[4 * j] = add2(v2[4 * j], pack2(h2_lo(a4[j].x), h2_hi(a4[j].x)));
                        v2[4 * j + 1] = add2(v2[4 * j + 1], pack2(h2_lo(a4[j].y), h2_hi(a4[j].y)));
                        v2[4 * j + 2] = add2(v2[4 * j + 2], pack2(h2_lo(a4[j].z), h2_hi(a4[j].z)));
                        v2[4 * j + 3] = add2(v2[4 * j + 3], pack2(h2_lo(a4[j].w), h2_hi(a4[j].w)));
                    }
                    wait_sfree();                      // earlier stores no longer read what is overwritten
#pragma unroll
                    for (int j = 0; j < 4; ++j) {      // in place over this thread's own residual bytes
                        float a0, a1, a2, a3, a4f, a5, a6, a7;
                        unpack2(v2[4 * j], a0, a1); unpack2(v2[4 * j + 1], a2, a3);
                        unpack2(v2[4 * j + 2], a4f, a5); unpack2(v2[4 * j + 3], a6, a7);
                        sts128_u32(arow + fast_chunk[j], pack_h2(a0, a1), pack_h2(a2, a3), pack_h2(a4f, a5), pack_h2(a6, a7));
                    }
                    signal_staged();
                }
            } else if (KIND == EPI_LINEAR && ACT == ACT_NONE && LEAN == 3) {
                // Lean residual epilogue with the bypass (feed_forward2 + bypass_mid, reference: zipformer.py:584-590):
                //   v = resid + A.W^T + b;  out = orig + (v - orig) * scale[col]
                // `orig` rides the aux ring next to the residual sub-tile (two slots per sub-tile); the scale comes from row 1
                // of the tile's bias cache.  Chunk by chunk (8 columns) to stay inside the register budget.
                const uint32_t aux_row0 = smem_u32(aux_smem) + static_cast<uint32_t>(r * 128);
                for (int s = half; s < n_sub; s += 2) {
                    const uint32_t q = (tile_iter * static_cast<uint32_t>(n_sub) + static_cast<uint32_t>(s)) * 2u;
                    const uint32_t slot = q % static_cast<uint32_t>(AUX_SLOTS);
                    const uint32_t slot2 = (q + 1u) % static_cast<uint32_t>(AUX_SLOTS);
                    const int c0 = (2 * s + part) * 32;
                    uint32_t acc_r[32];
                    tmem_ld32(taddr + c0, acc_r);
                    mbar_wait(&aux_full[slot], (q / static_cast<uint32_t>(AUX_SLOTS)) & 1u);
                    mbar_wait(&aux_full[slot2], ((q + 1u) / static_cast<uint32_t>(AUX_SLOTS)) & 1u);
                    const uint32_t arow = aux_row0 + slot * static_cast<uint32_t>(GEMM_AUX_BYTES);
                    const uint32_t orow = aux_row0 + slot2 * static_cast<uint32_t>(GEMM_AUX_BYTES);
                    tmem_ld_wait();
                    wait_sfree();                      // earlier stores no longer read what is overwritten
                    const uint32_t cbb = cb_addr + static_cast<uint32_t>(c0 * 4);     // row 0: bias, row 1 (+1024 B): scale (no row bias here)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 a4 = lds128_u32(arow + fast_chunk[j]);
                        const uint4 o4 = lds128_u32(orow + fast_chunk[j]);
                        const uint32_t av[4] = {a4.x, a4.y, a4.z, a4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
                        uint32_t hv[4];
#pragma unroll
                        for (int h2i = 0; h2i < 2; ++h2i) {
                            const uint4 bq = lds128_u32(cbb + static_cast<uint32_t>(j * 32 + h2i * 16));
                            const uint4 sq = lds128_u32(cbb + 1024u + static_cast<uint32_t>(j * 32 + h2i * 16));
                            const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w}, sw4[4] = {sq.x, sq.y, sq.z, sq.w};
#pragma unroll
                            for (int e2 = 0; e2 < 2; ++e2) {
                                const int pi = 2 * h2i + e2;                    // pair inside the chunk
                                const int ci = 8 * j + 2 * pi;                  // first column of the pair inside the unit
                                f32x2 v = add2(pack2(__uint_as_float(acc_r[ci]), __uint_as_float(acc_r[ci + 1])),
                                               pack2(__uint_as_float(bw[2 * e2]), __uint_as_float(bw[2 * e2 + 1])));
                                v = add2(v, pack2(h2_lo(av[pi]), h2_hi(av[pi])));
                                const f32x2 o2 = pack2(h2_lo(ov[pi]), h2_hi(ov[pi]));
                                const f32x2 d = fma2(o2, pack2(-1.0f, -1.0f), v);
                                const f32x2 y = fma2(d, pack2(__uint_as_float(sw4[2 * e2]), __uint_as_float(sw4[2 * e2 + 1])), o2);
                                float y0, y1;
                                unpack2(y, y0, y1);
                                hv[pi] = pack_h2(y0, y1);
                            }
                        }
                        sts128_u32(arow + fast_chunk[j], hv[0], hv[1], hv[2], hv[3]);   // in place over the residual bytes
                    }
                    signal_staged();
                }
            } else if (KIND == EPI_GATED) {
                const int hcols = p.block_n >> 1;                       // 128
                // the two parts of a (quarter, half) take adjacent 32-column units so that their fp16 rows
                // leave as one 128-byte segment
                const int u = 2 * half + part;
                const bool g_tma = p.tma_store && 2 * half * 32 < hcols;
                // this warp's rows of the half's staging ring (two 16 KB buffers)
                uint8_t* tbuf = aux_smem + ((kcount & dsh) * 2 + half) * GEMM_AUX_BYTES + quarter * 4096;
                if (g_tma) wait_sfree();
                __syncwarp();
                bs[lane] = nb0; bs[32 + lane] = nb1;
                __syncwarp();
                const int c0 = u * 32;
                const int oc = out_base + c0;
                int ncols = p.n_out - oc;
                ncols = ncols > 32 ? 32 : ncols;
                if (c0 < hcols && (ncols > 0 || g_tma)) {
                    uint32_t ra[32], rb[32];
                    tmem_ld32(taddr + c0, ra);
                    tmem_ld32(taddr + hcols + c0, rb);
                    tmem_ld_wait();
                    float v[32];
                    // the gate is chosen OUTSIDE the column loop: with the branch inside it every pair was its own basic
                    // block and ptxas ran the MUFU chains (ex2 -> add -> rcp -> mul) of the 16 pairs back to back
                    if (p.gate_mode == GATE_TANH_SX) gated_unit<true>(ra, rb, bs, masked, v);
                    else gated_unit<false>(ra, rb, bs, masked, v);
                    if (g_tma) stage_h16_unit(tbuf, lane, v, 4 * part);
                    else store_unit(p, private_stage, lane, row_ok, row, row0, rows_ok, oc, ncols, v, 0);
                }
                if (g_tma) signal_staged();
            } else if (lin_active) {
                for (int s = s_first; s < n_sub; s += s_step) {
                    const uint8_t* aux_row = nullptr;
                    const uint8_t* orig_row = nullptr;
                    int slot = 0, slot2 = 0;
                    if (p.aux_mode != AUX_NONE) {
                        const uint32_t q = (tile_iter * static_cast<uint32_t>(n_sub) + static_cast<uint32_t>(s)) *
                                           static_cast<uint32_t>(aux_parts);
                        slot = q % AUX_SLOTS;
                        mbar_wait(&aux_full[slot], (q / AUX_SLOTS) & 1u);
                        aux_row = aux_smem + slot * GEMM_AUX_BYTES + r * 128;
                        if (aux_parts == 2) {
                            slot2 = (q + 1) % AUX_SLOTS;
                            mbar_wait(&aux_full[slot2], ((q + 1) / AUX_SLOTS) & 1u);
                            orig_row = aux_smem + slot2 * GEMM_AUX_BYTES + r * 128;
                        }
                    }
                    // TMA-store staging of this warp's 32 rows x 128 B: in place over its rows of the consumed
                    // aux sub-tile, else the half's two-buffer ring
                    uint8_t* tbuf = (p.aux_mode != AUX_NONE ? aux_smem + slot * GEMM_AUX_BYTES
                                                            : aux_smem + ((kcount & dsh) * 2 + half) * GEMM_AUX_BYTES) +
                                    quarter * 4096;
                    __syncwarp();                            // the previous sub-tile's bias reads are done
                    bs[lane] = nb0;
                    __syncwarp();
                    if (s + s_step < n_sub)                  // request the next unit's bias now
                        nb0 = bias_at(((s + s_step) * units_per_sub + uu) * 32 + lane);
                    const int c0 = (s * units_per_sub + uu) * 32;
                    const int oc = out_base + c0;
                    int ncols = p.n_valid - c0;
                    if (p.n_out - oc < ncols) ncols = p.n_out - oc;
                    ncols = ncols > 32 ? 32 : ncols;
                    if (c0 < p.block_n && (ncols > 0 || p.tma_store)) {
                        if (threadIdx.x == 0 && tile == t_first && s == s_first) TL_STAMP(12);
                        uint32_t acc_r[32];
                        tmem_ld32(taddr + c0, acc_r);
                        tmem_ld_wait();
                        if (threadIdx.x == 0 && tile == t_first && s == s_first) TL_STAMP(13);
                        float v[32];
                        const f32x2 rs2 = pack2(rscale, rscale);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {       // acc * rowscale + bias, two columns per FFMA2
                            const float4 bq = *reinterpret_cast<const float4*>(bs + 4 * j);
                            const f32x2 lo = fma2(pack2(__uint_as_float(acc_r[4 * j]), __uint_as_float(acc_r[4 * j + 1])),
                                                  rs2, pack2(bq.x, bq.y));
                            const f32x2 hi = fma2(pack2(__uint_as_float(acc_r[4 * j + 2]), __uint_as_float(acc_r[4 * j + 3])),
                                                  rs2, pack2(bq.z, bq.w));
                            unpack2(lo, v[4 * j], v[4 * j + 1]);
                            unpack2(hi, v[4 * j + 2], v[4 * j + 3]);
                        }
                        const bool vec = ncols == 32 && (p.ldc & 3) == 0;
                        if (p.rowbias != nullptr && row_ok) {
                            const float* rbp = p.rowbias + grp * p.ld_rowbias + oc;
                            if (vec && (p.ld_rowbias & 3) == 0) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 q = __ldg(reinterpret_cast<const float4*>(rbp) + j);
                                    v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (i < ncols) v[i] += __ldg(rbp + i);
                            }
                        }
                        if (ACT != ACT_NONE && p.act_cols > 0 && oc + 32 > p.act_cols) {
                            // unit at or across the activation boundary of a merged GEMM: per column
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (oc + i < p.act_cols) v[i] = apply_act(v[i], ACT);
                        } else if (ACT == ACT_SWOOSH_L) {
#pragma unroll
                            for (int i = 0; i < 32; i += 2) swoosh_direct2(v[i], v[i + 1], SWOOSH_L_C, SWOOSH_L_K0);
                        } else if (ACT == ACT_SWOOSH_R) {
#pragma unroll
                            for (int i = 0; i < 32; i += 2) swoosh_direct2(v[i], v[i + 1], SWOOSH_R_C, SWOOSH_R_K0);
                        } else if (ACT == ACT_GELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                        }
                        if (p.aux_mode == AUX_ADD_H16) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 a = *reinterpret_cast<const uint4*>(aux_row + (((4 * uu + j) ^ (r & 7)) << 4));
                                v[8 * j] += h2_lo(a.x);     v[8 * j + 1] += h2_hi(a.x);
                                v[8 * j + 2] += h2_lo(a.y); v[8 * j + 3] += h2_hi(a.y);
                                v[8 * j + 4] += h2_lo(a.z); v[8 * j + 5] += h2_hi(a.z);
                                v[8 * j + 6] += h2_lo(a.w); v[8 * j + 7] += h2_hi(a.w);
                            }
                        } else if (p.aux_mode == AUX_MUL_H16) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 a = *reinterpret_cast<const uint4*>(aux_row + (((4 * uu + j) ^ (r & 7)) << 4));
                                v[8 * j] *= h2_lo(a.x);     v[8 * j + 1] *= h2_hi(a.x);
                                v[8 * j + 2] *= h2_lo(a.y); v[8 * j + 3] *= h2_hi(a.y);
                                v[8 * j + 4] *= h2_lo(a.z); v[8 * j + 5] *= h2_hi(a.z);
                                v[8 * j + 6] *= h2_lo(a.w); v[8 * j + 7] *= h2_hi(a.w);
                            }
                        }
                        if (orig_row != nullptr) {        // bypass, `orig` sub-tile staged by TMA
                            const float* sp = p.bypass_scale + oc;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 o4 = *reinterpret_cast<const uint4*>(orig_row + (((4 * uu + j) ^ (r & 7)) << 4));
                                const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp) + 2 * j);
                                const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp) + 2 * j + 1);
                                float o;
                                o = h2_lo(o4.x); v[8 * j] = fmaf(v[8 * j] - o, s0.x, o);
                                o = h2_hi(o4.x); v[8 * j + 1] = fmaf(v[8 * j + 1] - o, s0.y, o);
                                o = h2_lo(o4.y); v[8 * j + 2] = fmaf(v[8 * j + 2] - o, s0.z, o);
                                o = h2_hi(o4.y); v[8 * j + 3] = fmaf(v[8 * j + 3] - o, s0.w, o);
                                o = h2_lo(o4.z); v[8 * j + 4] = fmaf(v[8 * j + 4] - o, s1.x, o);
                                o = h2_hi(o4.z); v[8 * j + 5] = fmaf(v[8 * j + 5] - o, s1.y, o);
                                o = h2_lo(o4.w); v[8 * j + 6] = fmaf(v[8 * j + 6] - o, s1.z, o);
                                o = h2_hi(o4.w); v[8 * j + 7] = fmaf(v[8 * j + 7] - o, s1.w, o);
                            }
                        }
                        if (masked) {                  // frames past the utterance: the row is zero (vocoder, audio.cuh)
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                        }
                        if (threadIdx.x == 0 && tile == t_first && s == s_first) TL_STAMP(14);
                        if (p.tma_store) {             // stage into the (swizzled) TMA box, own row only
                            wait_sfree();              // earlier stores no longer read what is overwritten
                            if (out_f32) {
                                uint8_t* my = tbuf + lane * 128;
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    *reinterpret_cast<float4*>(my + ((j ^ (lane & 7)) << 4)) =
                                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            } else {
                                stage_h16_unit(tbuf, lane, v, 4 * uu);
                            }
                        } else {
                            // in place over the consumed aux rows of this warp, or in the warp's private area
                            uint8_t* stage = p.aux_mode != AUX_NONE
                                                 ? aux_smem + slot * GEMM_AUX_BYTES + quarter * 32 * 128
                                                 : private_stage;
                            __syncwarp();              // every lane has consumed its aux row
                            store_unit(p, stage, lane, row_ok, row, row0, rows_ok, oc, ncols, v,
                                       p.aux_mode != AUX_NONE ? 4 * uu : 0);
                        }
                    } else if (p.tma_store) {
                        wait_sfree();                  // keeps the buffer hand-shake in step (no columns to stage)
                    }
                    if (threadIdx.x == 0 && tile == t_first && s == s_first) TL_STAMP(15);
                    if (p.tma_store) {
                        signal_staged();
                        continue;
                    }
                    if (p.aux_mode != AUX_NONE) {
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(&aux_empty[slot]);
                            if (aux_parts == 2) mbar_arrive(&aux_empty[slot2]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (threadIdx.x == 0 && tile == t_first) TL_STAMP(6);
            if (lane == 0) {                          // the leader's MMA thread owns the accumulator hand-shake
                if (CLUSTER == 2 && crank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }

    }

    tc_fence_before();
    __syncthreads();
#ifdef ZVB_TIMELINE
    if (threadIdx.x == 0 && blockIdx.x == 0 && tl_slot < TL_MAX) { g_tl[tl_slot][7] = clock64(); g_tl[tl_slot][9] = globaltimer_ns(); }
#endif
    if (CLUSTER > 1) cluster_sync_all();        // no CTA exits while a peer may still write to it
    if (warp == W_MMA) {
        __syncwarp();
        tc_fence_after();
        if (CLUSTER == 2) tmem_dealloc_2sm(tmem_base, GEMM_TMEM_COLS);
        else tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
    }
}

}  // namespace zvb
