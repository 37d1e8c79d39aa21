// Attention-weights kernel, tensor-core rel-pos bias (reference: modules/zipformer.py:1149-1306
// RelPositionMultiheadAttentionWeights.forward, eval path):
//   P[n,h,i,:] = softmax_j( q_i·k_j + p_i·E_h[j-i] , key-padding mask -> -1000 )   (fp32 softmax)
// Same contract as attn.cuh (unnormalised fp16 weights 2^12·2^(s2 - m2) + fp32 1/rowsum, two passes, cheap /
// exact row maximum), but the rel-pos bias no longer costs CUDA-core work per score (4 half2 FMAs + a 16-byte
// shared-memory read per score pair in attn.cuh: the kernel was issue bound, 14.6 warp instructions per 32
// scores, ncu round 2).  The skewed bias  bias[r][c] = p_r · E[(j0+c) - (i0+r)]  of a 128 x 128 score tile IS a
// matrix product once the row's lane inside its TMEM quarter is folded into the contraction:
//     r = 32 q + l :  bias[r][c] = D[r][x],  x = c - 32 q + 96,
//     D[r][x] = sum_{l',d} A'[r][(l',d)] · E[b0 + 31 + x - l'][d],   A'[r][(l',d)] = p_r[d] · [l' == r mod 32]
// (b0 = j0 - i0 + L - 128).  A' is a 128 x 128 fp16 operand (four non-zeros per row) built once per CTA; the B
// operand is a Toeplitz matrix, and in the un-swizzled K-major core-matrix layout a Toeplitz operand needs no
// materialisation: with the contraction ordered as chunks kc = 0..15 of {(l'=31-2kc, d0..3), (l'=30-2kc, d0..3)}
// the 16-byte chunk (row y, chunk kc) of the operand for the EVEN columns x = 2y is {E[b0+2(y+kc)], E[b0+2(y+kc)+1]}
// -- 16 bytes at offset 16·(y+kc) of the plain E array starting at entry b0 -- so the shared-memory descriptor is
// just that array with LBO = 16 B (next chunk) and SBO = 128 B (next 8 rows): overlapping core matrices.  The ODD
// columns x = 2y+1 use the same array starting one entry later.  Two MMAs (N = 112 each, K = 128) per tile write
// D_even / D_odd into TMEM; quarter q of the softmax warps reads its window of D at a warp-uniform column offset
// (tcgen05.ld takes one column base per warp, which is why the lane had to go into the contraction).
// Per score the CUDA cores now do: 1 FADD + 1 FFMA (packed x2), a half conversion, the MUFU exponential, the row-sum
// add and the staging store -- about 3 instructions instead of 10.
// One CTA per (128 queries, head, utterance), one CTA per SM (TMEM: 2 x 128 score columns + 2 x 112 bias columns);
// 16 softmax warps = 4 TMEM lane quarters x 4 column units of 32 keys; every thread pulls its 32 scores + 32 bias
// values into registers at the start of a tile and hands the TMEM buffers back at once, so the next tile's MMAs
// run under this tile's exponentials.
#pragma once
#include "ptx.cuh"

namespace zvb {

constexpr int A3_BM = 128;
constexpr int A3_BN = 128;
constexpr int A3_KSTAGES = 4;
constexpr int A3_TILE_BYTES = 128 * 64 * 2;        // one 128-row x 64-col fp16 box (32 cols used), 128B swizzle
constexpr int A3_SM_WARPS = 16;
constexpr int A3_THREADS = 32 * (A3_SM_WARPS + 2);
constexpr int A3_TMEM_COLS = 512;
constexpr int A3_ND = 112;                          // columns of D_even / D_odd
constexpr int A3_COL_DE = 256, A3_COL_DO = 256 + A3_ND;
constexpr int A3_APRIME_BYTES = 128 * 128 * 2;      // A': 16 row groups x 16 chunks x 128 B
constexpr int A3_EWIN_BYTES = 2048;                 // 128 chunks of 16 B per operand (127 used)
constexpr int A3_ESTAGES = 2;
constexpr int A3_PADZ = 128;                        // zero entries in front of E in the table (weights.py)
constexpr int A3_STAGE_BYTES = 16 * 2 * 2048;       // [16 warps][2 buffers][32 rows x 64 B] (64-byte swizzle)
constexpr float A3_BOUND_SLACK_L2 = 10.0f;
constexpr int A3_SMEM_BYTES = (1 + A3_KSTAGES) * A3_TILE_BYTES + A3_APRIME_BYTES + A3_ESTAGES * 2 * A3_EWIN_BYTES +
                              A3_STAGE_BYTES + 4 * 128 * 4 + 64 + 256 + 1024 /*alignment*/;

struct Attn3Params {
    int L, Lk, H, N;
    int qd;                          // H * 32: column of head-0 keys inside a qkp row
    const __half* qkp;               // [N*L, ld] = [q | k | p]
    int ld;
    const uint2* Z;                  // [H][2][LZ] entries of 4 fp16 = log2e·E_h[r][0..3] at index PADZ + r of copy 0,
                                     // zeros elsewhere; copy 1 is copy 0 shifted by one entry (Z1[m] = Z0[m+1])
    int LZ;                          // entries per copy (even)
    const float* emax;               // [H] max_r |log2e·E[h][r]|_2
    const uint32_t* maskw;           // [N][mask_words] excluded-key bits (mask_words_kernel)
    int mask_words;
    __half* P;                       // [N][H][L][Lk] (written through tma_p)
    float* inv_l;                    // [N][H][L]
    int dbg;                         // timing ablations (ZVB_ATTN_DBG): 1 bias in both passes, 2 no bias MMAs, 4 no stores
};

// un-swizzled K-major shared-memory descriptor: LBO = byte offset between the two 16-byte K chunks of one MMA,
// SBO = byte offset between groups of 8 rows
__device__ __forceinline__ uint64_t umma_desc_k_plain(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;                        // layout type 0: no swizzle
}

// tma_qk: [q | k | p] rows, box 64 x 128 (128B swizzle); tma_p: P viewed as (Lk, L, N*H), box 32 columns x 32 rows
// (64B swizzle): every softmax warp stores its own 32 x 32 block, no barrier between warps.
__global__ void __launch_bounds__(A3_THREADS, 1)
attn_weights_tc_kernel(const __grid_constant__ CUtensorMap tma_qk, const __grid_constant__ CUtensorMap tma_p,
                       const Attn3Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* q_tile = smem;
    uint8_t* k_tiles = smem + A3_TILE_BYTES;
    uint8_t* aprime = smem + (1 + A3_KSTAGES) * A3_TILE_BYTES;
    uint8_t* ewin = aprime + A3_APRIME_BYTES;                                   // [ESTAGES][even | odd][2048]
    uint8_t* stage_all = ewin + A3_ESTAGES * 2 * A3_EWIN_BYTES;                  // [2][8][32][128 B]
    float* xch = reinterpret_cast<float*>(stage_all + A3_STAGE_BYTES);           // [4 units][128 rows]
    float* flag = xch + 4 * 128;                                                 // [16] per-warp max |p_i|
    uint64_t* bars = reinterpret_cast<uint64_t*>(flag + 16);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                      // [KSTAGES]
    uint64_t* k_empty = k_full + A3_KSTAGES;          // [KSTAGES]
    uint64_t* s_full = k_empty + A3_KSTAGES;          // [2]
    uint64_t* s_empty = s_full + 2;                   // [2]
    uint64_t* d_full = s_empty + 2;                   // [1]
    uint64_t* d_empty = d_full + 1;                   // [1]
    uint64_t* e_full = d_empty + 1;                   // [ESTAGES]
    uint64_t* e_empty = e_full + A3_ESTAGES;          // [ESTAGES]
    uint64_t* a_ready = e_empty + A3_ESTAGES;         // [1] A' built
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(a_ready + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * A3_BM;
    const int h = blockIdx.y;
    const int n = blockIdx.z;
    const int num_jt = (p.L + A3_BN - 1) / A3_BN;
    const int total_it = 2 * num_jt;
    constexpr int W_TMA = A3_SM_WARPS, W_MMA = A3_SM_WARPS + 1;

    if (warp == W_TMA && lane == 0) {
        tma_prefetch_desc(&tma_qk);
        tma_prefetch_desc(&tma_p);
        mbar_init(q_full, 1);
        for (int s = 0; s < A3_KSTAGES; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], A3_SM_WARPS);
        }
        mbar_init(d_full, 1);
        mbar_init(d_empty, A3_SM_WARPS);
        for (int s = 0; s < A3_ESTAGES; ++s) {
            mbar_init(&e_full[s], 1);
            mbar_init(&e_empty[s], 1);
        }
        mbar_init(a_ready, A3_SM_WARPS);
        fence_barrier_init();
    }
    if (warp == W_MMA) {
        tmem_alloc(tmem_holder, A3_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                 // set-up above overlaps the previous kernel's tail
    pdl_launch();

    // CTA-uniform choice between the cheap softmax shift (row max of q.k + |p_i| max|E|, first pass without bias)
    // and the exact one (first pass evaluates the bias too); see attn.cuh.  Decided by the softmax warps from the
    // p rows they load anyway, shared through `flag` and named barrier 2.
    const float emax_h = __ldg(p.emax + h);

    if (warp == W_TMA) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        auto load_k = [&](int it) {
            const int jt = it >= num_jt ? it - num_jt : it;
            mbar_wait(&k_empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&k_full[stage], A3_TILE_BYTES);
            tma_load_3d(k_tiles + stage * A3_TILE_BYTES, &tma_qk, &k_full[stage], p.qd + h * 32, jt * A3_BN, n);
            if (++stage == A3_KSTAGES) { stage = 0; phase ^= 1u; }
        };
        const int prefill = total_it < A3_KSTAGES ? total_it : A3_KSTAGES;
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, A3_TILE_BYTES);
            tma_load_3d(q_tile, &tma_qk, q_full, h * 32, i0, n);
            for (int it = 0; it < prefill; ++it) load_k(it);
        }
        __syncwarp();
        asm volatile("bar.sync 2, 576;" ::: "memory");
        if (lane == 0) {
            float pmax = flag[0];
            for (int w = 1; w < A3_SM_WARPS; ++w) pmax = fmaxf(pmax, flag[w]);
            const bool exact_max = 2.0f * pmax * emax_h > A3_BOUND_SLACK_L2 || (p.dbg & 1);
            const int e_first = (p.dbg & 2) ? total_it : exact_max ? 0 : num_jt;            // first iteration that needs the rel-pos window
            const uint2* Zh = p.Z + static_cast<long long>(h) * 2 * p.LZ;
            for (int it = 0; it < total_it; ++it) {
                const int jt = it >= num_jt ? it - num_jt : it;
                if (it >= prefill) load_k(it);
                if (it >= e_first) {
                    const uint32_t eu = static_cast<uint32_t>(it - e_first);
                    const int eb = static_cast<int>(eu % A3_ESTAGES);
                    mbar_wait(&e_empty[eb], ((eu / A3_ESTAGES) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&e_full[eb], 2 * A3_EWIN_BYTES);
                    // entry index of E[b0] in copy 0; each operand starts at an EVEN entry of one of the two copies
                    const int s0 = A3_PADZ + (jt * A3_BN - i0) + p.L - 128;
                    const uint2* src_even = (s0 & 1) ? Zh + p.LZ + (s0 - 1) : Zh + s0;
                    const uint2* src_odd = (s0 & 1) ? Zh + (s0 + 1) : Zh + p.LZ + s0;
                    uint8_t* dst = ewin + eb * 2 * A3_EWIN_BYTES;
                    bulk_load_1d(dst, src_even, A3_EWIN_BYTES, &e_full[eb]);
                    bulk_load_1d(dst + A3_EWIN_BYTES, src_odd, A3_EWIN_BYTES, &e_full[eb]);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        asm volatile("bar.sync 2, 576;" ::: "memory");
        if (lane == 0) {
            float pmax = flag[0];
            for (int w = 1; w < A3_SM_WARPS; ++w) pmax = fmaxf(pmax, flag[w]);
            const bool exact_max = 2.0f * pmax * emax_h > A3_BOUND_SLACK_L2 || (p.dbg & 1);
            const int e_first = (p.dbg & 2) ? total_it : exact_max ? 0 : num_jt;
            const uint32_t idesc_s = umma_idesc_f16(A3_BN);
            const uint32_t idesc_d = umma_idesc_f16(A3_ND);
            mbar_wait(q_full, 0);
            tc_fence_after();
            const uint64_t dq = umma_desc_k_sw128(smem_u32(q_tile));
            // A': chunk kc at +128 B, 8-row groups at +2048 B; one MMA covers two chunks
            const uint64_t da0 = umma_desc_k_plain(smem_u32(aprime), 128u, 2048u);
            int stage = 0;
            uint32_t phase = 0;
            bool a_waited = false;
            for (int it = 0; it < total_it; ++it) {
                const int acc = it & 1;
                mbar_wait(&s_empty[acc], (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
                mbar_wait(&k_full[stage], phase);
                tc_fence_after();
                const uint64_t dk = umma_desc_k_sw128(smem_u32(k_tiles + stage * A3_TILE_BYTES));
                const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(acc) * A3_BN;
                umma_f16(tmem_s, dq, dk, idesc_s, 0u);            // head-dim columns  0..15
                umma_f16(tmem_s, dq + 2, dk + 2, idesc_s, 1u);    // head-dim columns 16..31
                umma_commit(&k_empty[stage]);
                umma_commit(&s_full[acc]);
                if (++stage == A3_KSTAGES) { stage = 0; phase ^= 1u; }
                if (it >= e_first) {
                    const uint32_t eu = static_cast<uint32_t>(it - e_first);
                    const int eb = static_cast<int>(eu % A3_ESTAGES);
                    if (!a_waited) { mbar_wait(a_ready, 0); a_waited = true; }
                    mbar_wait(d_empty, (eu & 1u) ^ 1u);
                    mbar_wait(&e_full[eb], (eu / A3_ESTAGES) & 1u);
                    tc_fence_after();
                    const uint32_t eaddr = smem_u32(ewin + eb * 2 * A3_EWIN_BYTES);
                    const uint64_t dbe = umma_desc_k_plain(eaddr, 16u, 128u);
                    const uint64_t dbo = umma_desc_k_plain(eaddr + A3_EWIN_BYTES, 16u, 128u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {     // K = 128: 8 steps of two 16-byte chunks; A' +256 B, E +32 B per step
                        umma_f16(tmem_base + A3_COL_DE, da0 + static_cast<uint64_t>(16 * k), dbe + static_cast<uint64_t>(2 * k),
                                 idesc_d, k != 0 ? 1u : 0u);
                        umma_f16(tmem_base + A3_COL_DO, da0 + static_cast<uint64_t>(16 * k), dbo + static_cast<uint64_t>(2 * k),
                                 idesc_d, k != 0 ? 1u : 0u);
                    }
                    umma_commit(&e_empty[eb]);
                    umma_commit(d_full);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps (0..15)
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
        const int unit = warp >> 2;                       // key columns [32*unit, 32*unit + 32) of every tile
        const int r = quarter * 32 + lane;                // row inside the query tile
        const int i = i0 + r;
        const bool row_ok = i < p.L;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        uint2 pw = make_uint2(0u, 0u);
        if (row_ok) {
            const __half* pp = p.qkp + (static_cast<long long>(n) * p.L + i) * p.ld + 2 * p.qd + h * 4;
            pw = *reinterpret_cast<const uint2*>(pp);
            p0 = h2_lo(pw.x); p1 = h2_hi(pw.x); p2 = h2_lo(pw.y); p3 = h2_hi(pw.y);
        }
        const float pn = sqrtf(p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3);
        {
            float bi = pn;
#pragma unroll
            for (int q = 16; q > 0; q >>= 1) bi = fmaxf(bi, __shfl_xor_sync(0xffffffffu, bi, q));
            if (lane == 0) flag[warp] = bi;
        }
        // A' (un-swizzled K-major, chunk kc at +128 B, 8-row group at +2048 B, row at +16 B): zero it, then every row
        // places its four p values in chunk kc = (31 - l) >> 1, first half for odd l, second half for even l
        {
            uint4* az = reinterpret_cast<uint4*>(aprime);
            for (int k = threadIdx.x; k < A3_APRIME_BYTES / 16; k += 32 * A3_SM_WARPS) az[k] = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (unit == 0) {
            const int l = lane;
            const int kc = (31 - l) >> 1;
            uint8_t* dst = aprime + (r >> 3) * 2048 + kc * 128 + (r & 7) * 16 + ((l & 1) ? 0 : 8);
            *reinterpret_cast<uint2*>(dst) = pw;
        }
        fence_proxy_async_smem();                         // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
        asm volatile("bar.sync 2, 576;" ::: "memory");
        float pmax_cta = flag[0];
#pragma unroll
        for (int w = 1; w < A3_SM_WARPS; ++w) pmax_cta = fmaxf(pmax_cta, flag[w]);
        const bool exact_max = 2.0f * pmax_cta * emax_h > A3_BOUND_SLACK_L2 || (p.dbg & 1);
        const int e_first = (p.dbg & 2) ? total_it : exact_max ? 0 : num_jt;

        const uint32_t* mwrow = p.maskw + static_cast<long long>(n) * p.mask_words + unit;
        // staging: each warp owns two 2 KB buffers of 32 rows x 64 B; chunk j of row `lane` sits at j ^ ((lane >> 1) & 3)
        // (the 64-byte swizzle of the store's tensor map: conflict-free 16-byte shared-memory writes)
        const int qbar = 3 + quarter;                     // named barriers 3..6: the four unit warps of a quarter
        const uint32_t stage_mine = smem_u32(stage_all) + static_cast<uint32_t>(warp * 4096);
        uint8_t* stage_ptr = stage_all + warp * 4096;
        uint32_t chunk_off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) chunk_off[j] = static_cast<uint32_t>(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
        const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(32 * unit);
        const uint32_t dcol = static_cast<uint32_t>((32 * unit - 32 * quarter + 96) >> 1);     // multiple of 16
        const uint32_t t_de = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + A3_COL_DE + dcol;
        const uint32_t t_do = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + A3_COL_DO + dcol;
        constexpr float LOG2E = 1.4426950408889634f;
        float m_run = -INFINITY, m_l2 = 0.f, l_run = 0.f;
        uint32_t tile2 = 0;                               // tiles staged so far (second pass)
        for (int it = 0; it < total_it; ++it) {
            const int pass = it >= num_jt ? 1 : 0;
            const int jt = pass ? it - num_jt : it;
            const int acc = it & 1;
            const bool with_bias = it >= e_first;
            if (it == num_jt) {                           // between the passes: m >= every score of the row (log2 units)
                xch[unit * 128 + r] = m_run;
                asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
                const float mr = fmaxf(fmaxf(xch[r], xch[128 + r]), fmaxf(xch[256 + r], xch[384 + r]));
                const float mb = mr == -INFINITY ? 0.f : mr;
                m_l2 = (exact_max ? mb : mb + pn * emax_h) - 12.0f;       // weights are stored scaled by 2^12
            }
            const uint32_t excl = __ldg(mwrow + 4 * jt);
            uint32_t s_r[32], e_r[16], o_r[16];
            mbar_wait(&s_full[acc], static_cast<uint32_t>(it >> 1) & 1u);
            if (with_bias) mbar_wait(d_full, static_cast<uint32_t>(it - e_first) & 1u);
            tc_fence_after();
            tmem_ld32(t_s + static_cast<uint32_t>(acc * A3_BN), s_r);
            if (with_bias) {
                tmem_ld16(t_de, e_r);
                tmem_ld16(t_do, o_r);
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                              // the values live in registers: hand the TMEM buffers back
                mbar_arrive(&s_empty[acc]);
                if (with_bias) mbar_arrive(d_empty);
            }
            if (pass == 0) {
                float cm = -INFINITY;
                if (with_bias) {                          // exact maximum of the biased scores
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float a = fmaf(__uint_as_float(s_r[2 * c]), LOG2E, __uint_as_float(e_r[c]));
                        const float b = fmaf(__uint_as_float(s_r[2 * c + 1]), LOG2E, __uint_as_float(o_r[c]));
                        cm = fmaxf(cm, ((excl >> (2 * c)) & 1u) ? -INFINITY : a);
                        cm = fmaxf(cm, ((excl >> (2 * c + 1)) & 1u) ? -INFINITY : b);
                    }
                } else if (excl == 0u) {
#pragma unroll
                    for (int c = 0; c < 32; c += 2)
                        cm = fmaxf(fmaxf(cm, __uint_as_float(s_r[c])), __uint_as_float(s_r[c + 1]));
                    cm *= LOG2E;
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        cm = fmaxf(cm, ((excl >> c) & 1u) ? -INFINITY : __uint_as_float(s_r[c]));
                    cm *= LOG2E;
                }
                m_run = fmaxf(m_run, cm);
                continue;
            }
            // ---- second pass: weights of this warp's 32 columns
            const int jc = jt * A3_BN + 32 * unit;        // first key column of this warp
            if (jc >= p.Lk || (p.dbg & 4)) continue;      // warp-uniform: past the padded width
            const uint32_t buf = tile2 & 1u;
            if (lane == 0) bulk_wait_read<1>();           // the store that read this buffer two tiles ago is done
            __syncwarp();
            uint32_t w[16];
            {
                const f32x2 nm = pack2(-m_l2, -m_l2);
                const f32x2 l2 = pack2(LOG2E, LOG2E);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const f32x2 b = add2(pack2(__uint_as_float(e_r[c]), __uint_as_float(o_r[c])), nm);
                    const f32x2 e2 = fma2(pack2(__uint_as_float(s_r[2 * c]), __uint_as_float(s_r[2 * c + 1])), l2, b);
                    float ea, eb2;
                    unpack2(e2, ea, eb2);
                    w[c] = ex2_h2(pack_h2(ea, eb2));
                }
                if (excl != 0u) {           // rare: a real branch, so unmasked tiles issue no selects
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const uint32_t keep = (((excl >> (2 * c)) & 1u) ? 0u : 0x0000FFFFu) |
                                              (((excl >> (2 * c + 1)) & 1u) ? 0u : 0xFFFF0000u);
                        w[c] &= keep;
                    }
                }
                // half2 partial row sums of 8 pairs (<= 8 x 2^12 per half), then fp32
                const uint32_t sa = hadd2(hadd2(hadd2(w[0], w[1]), hadd2(w[2], w[3])), hadd2(hadd2(w[4], w[5]), hadd2(w[6], w[7])));
                const uint32_t sb = hadd2(hadd2(hadd2(w[8], w[9]), hadd2(w[10], w[11])),
                                          hadd2(hadd2(w[12], w[13]), hadd2(w[14], w[15])));
                l_run += (h2_lo(sa) + h2_hi(sa)) + (h2_lo(sb) + h2_hi(sb));
                const uint32_t row_addr = stage_mine + buf * 2048u;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128_u32(row_addr + chunk_off[j], w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tma_p, stage_ptr + buf * 2048, jc, i0 + quarter * 32, n * p.H + h);
                bulk_commit();
            }
            ++tile2;
        }
        xch[unit * 128 + r] = l_run;
        asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
        if (unit == 0 && row_ok) {
            const float l = (xch[r] + xch[128 + r]) + (xch[256 + r] + xch[384 + r]);
            p.inv_l[(static_cast<long long>(n) * p.H + h) * p.L + i] = l > 0.f ? 1.0f / l : 0.f;
        }
        if (lane == 0) bulk_wait_read<0>();               // staging must outlive the stores reading it
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, A3_TMEM_COLS);
    }
}

}  // namespace zvb
