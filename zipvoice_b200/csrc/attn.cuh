// Attention-weights kernel (reference: modules/zipformer.py:1149-1306
// RelPositionMultiheadAttentionWeights.forward, eval path):
//   P[n,h,i,:] = softmax_j( q_i·k_j + p_i·E_h[j-i] , key-padding mask -> -1000 )   (fp32 softmax)
// q·kᵀ runs on tcgen05 (K = 32: two UMMA_K steps per 128x128 score tile, accumulators in
// TMEM, double buffered); the rel-pos bias is a 4-term dot product per score with the
// per-layer table E = linear_pos(pos_emb) staged in shared memory as the (i-tile, j-tile)
// window of 255 relative offsets.  Softmax is split so that every score costs one exp and one
// bias evaluation: pass 1 only takes the row max of q·kᵀ from TMEM (no bias, no exp); with
// m = max_j q·k_j + |p_i|·max_r|E_h[r]| >= every score of the row, pass 2 recomputes the cheap q·kᵀ,
// adds the bias and writes the UNNORMALISED weights exp(s - m) in (0, 1] as fp16 together with the
// fp32 row sum's reciprocal; the three consumers of the layer (NonlinAttention, SelfAttention x2)
// apply 1/l in their GEMM epilogue.  Masked keys get exactly 0 (exp(-1000 - m) == 0 in fp32).
// The softmax is latency bound (TMEM load -> LDS of the rel-pos entries -> FFMA chain -> MUFU), so a
// CTA runs EIGHT softmax warps -- two per TMEM lane quarter, each owning 64 of the tile's 128 key
// columns and exchanging row max / row sum through shared memory -- and two CTAs share an SM
// (<= 102 registers per thread); P leaves through TMA box stores (32 rows x 64 columns per warp).
#pragma once
#include "ptx.cuh"

namespace zvb {

constexpr int ATT_BM = 128;          // queries per CTA
constexpr int ATT_BN = 128;          // keys per score tile
constexpr int ATT_KSTAGES = 3;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;      // one 128-row x 64-col fp16 box (only 32 cols used)
constexpr int ATT_SM_WARPS = 8;      // softmax warps: two per TMEM lane quarter, 64 key columns each
constexpr int ATT_THREADS = 64 + 32 * ATT_SM_WARPS;
constexpr int ATT_TMEM_COLS = 256;
constexpr int ATT_EWIN = 256;        // 255 offsets used
constexpr int ATT_STAGE_BYTES = ATT_SM_WARPS * 4096; // per softmax warp: 32 rows x 128 B TMA-store staging
constexpr int ATT_SMEM_BYTES = (1 + ATT_KSTAGES) * ATT_TILE_BYTES + ATT_STAGE_BYTES + 2 * ATT_EWIN * 16 +
                               2 * 16 + 2 * 128 * 4 + 1024 + 256;

struct AttnParams {
    int L, Lk, H, N;
    int qd;                          // H * 32: column of head-0 keys inside a qkp row
    const __half* qkp;        // [N*L, ld] = [q | k | p]
    int ld;
    const float* E;                  // [H][2L-1][4] followed by [H] floats: max_r |E[h][r]|_2
    const uint8_t* mask;             // [N][L], non-zero = padded key
    __half* P;                // [N][H][L][Lk] unnormalised weights exp(s - m) (written through tma_p)
    float* inv_l;                    // [N][H][L]     1 / row sum
};

// tma_qk: [q | k | p] rows, box 64 x 128; tma_p: P viewed as (Lk, L, N*H), box 64 columns x 32 rows.
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_weights_kernel(const __grid_constant__ CUtensorMap tma_qk, const __grid_constant__ CUtensorMap tma_p,
                    const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* q_tile = smem;
    uint8_t* k_tiles = smem + ATT_TILE_BYTES;
    uint8_t* stage_all = smem + (1 + ATT_KSTAGES) * ATT_TILE_BYTES;                        // [8][32][128 B]
    // rel-pos window of the (i-tile, j-tile): 255 offsets x 4 floats, pre-scaled by log2 e
    float4* ewin = reinterpret_cast<float4*>(stage_all + ATT_STAGE_BYTES);                  // [2][256]
    uint32_t* mwin = reinterpret_cast<uint32_t*>(ewin + 2 * ATT_EWIN);                      // [2][4] excluded-key bits
    float* xch = reinterpret_cast<float*>(mwin + 8);                                        // [2][128] half <-> half
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 256);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                      // [KSTAGES]
    uint64_t* k_empty = k_full + ATT_KSTAGES;         // [KSTAGES]
    uint64_t* s_full = k_empty + ATT_KSTAGES;         // [2]
    uint64_t* s_empty = s_full + 2;                   // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(s_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * ATT_BM;
    const int h = blockIdx.y;
    const int n = blockIdx.z;
    const int num_jt = (p.L + ATT_BN - 1) / ATT_BN;
    const int total_it = 2 * num_jt;

    if (warp == ATT_SM_WARPS && lane == 0) {
        tma_prefetch_desc(&tma_qk);
        tma_prefetch_desc(&tma_p);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_KSTAGES; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], ATT_SM_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == ATT_SM_WARPS + 1) {
        tmem_alloc(tmem_holder, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == ATT_SM_WARPS) {                           // TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(q_tile, &tma_qk, q_full, h * 32, i0, n);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < total_it; ++it) {
                const int jt = it >= num_jt ? it - num_jt : it;
                mbar_wait(&k_empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&k_full[stage], ATT_TILE_BYTES);
                tma_load_3d(k_tiles + stage * ATT_TILE_BYTES, &tma_qk, &k_full[stage], p.qd + h * 32,
                            jt * ATT_BN, n);
                if (++stage == ATT_KSTAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == ATT_SM_WARPS + 1) {                // MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(ATT_BN);
            mbar_wait(q_full, 0);
            tc_fence_after();
            const uint64_t dq = umma_desc_k_sw128(smem_u32(q_tile));
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < total_it; ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = static_cast<uint32_t>(it >> 1) & 1u;
                mbar_wait(&s_empty[acc], acc_phase ^ 1u);
                mbar_wait(&k_full[stage], phase);
                tc_fence_after();
                const uint64_t dk = umma_desc_k_sw128(smem_u32(k_tiles + stage * ATT_TILE_BYTES));
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc) * ATT_BN;
                umma_f16(tmem_d, dq, dk, idesc, 0u);            // head-dim columns  0..15
                umma_f16(tmem_d, dq + 2, dk + 2, idesc, 1u);    // head-dim columns 16..31
                umma_commit(&k_empty[stage]);
                umma_commit(&s_full[acc]);
                if (++stage == ATT_KSTAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps (0..7); the two
        // single-thread roles have the highest warp ids, which the SMSP arbiter favours
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
        const int tid = threadIdx.x;                      // 0..255
        const int half = tid >> 7;                        // key columns [64*half, 64*half + 64) of every tile
        const int r = quarter * 32 + lane;                // row inside the query tile
        const int i = i0 + r;
        const bool row_ok = i < p.L;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (row_ok) {
            const __half* pp = p.qkp + (static_cast<long long>(n) * p.L + i) * p.ld + 2 * p.qd + h * 4;
            const uint2 w = *reinterpret_cast<const uint2*>(pp);
            p0 = h2_lo(w.x); p1 = h2_hi(w.x); p2 = h2_lo(w.y); p3 = h2_hi(w.y);
        }
        const float4* Eh = reinterpret_cast<const float4*>(p.E) + static_cast<long long>(h) * (2 * p.L - 1);
        const uint8_t* mrow = p.mask + static_cast<long long>(n) * p.L;
        // P leaves through TMA stores: every warp stages 32 rows x 64 columns (128-byte rows, 128B
        // swizzle) and one lane issues the box store; rows >= L and columns >= Lk are clipped by the map
        uint8_t* stage = stage_all + (tid >> 5) * 4096;
        uint8_t* my = stage + lane * 128;
        const int sw = lane & 7;
        constexpr float LOG2E = 1.4426950408889634f;
        float m_run = -INFINITY, m_l2 = 0.f, l_run = 0.f;
        // mask byte / rel-pos entry this thread stages for iteration `it2` (global loads, prefetched)
        float4 e_cur = make_float4(0.f, 0.f, 0.f, 0.f);
        bool ex_cur = true;
        auto fetch = [&](int it2) {
            const int pass2 = it2 >= num_jt ? 1 : 0;
            const int j02 = (pass2 ? it2 - num_jt : it2) * ATT_BN;
            if (tid < 128) {
                const int j = j02 + tid;
                ex_cur = j < p.L ? (mrow[j] != 0) : true;
            }
            e_cur = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pass2 && tid < 255) {
                const int rel = (j02 - i0) - 127 + tid + (p.L - 1);
                if (rel >= 0 && rel <= 2 * p.L - 2) e_cur = __ldg(Eh + rel);
            }
        };
        fetch(0);

        for (int it = 0; it < total_it; ++it) {
            const int pass = it >= num_jt ? 1 : 0;
            const int jt = pass ? it - num_jt : it;
            const int j0 = jt * ATT_BN;
            const int acc = it & 1;
            const uint32_t acc_phase = static_cast<uint32_t>(it >> 1) & 1u;
            if (it == num_jt) {                           // between the passes: m >= every score of the row
                xch[half * 128 + r] = m_run;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float mr = fmaxf(xch[r], xch[128 + r]);
                const float emax = __ldg(p.E + static_cast<long long>(p.H) * (2 * p.L - 1) * 4 + h);
                const float pn = sqrtf(p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3);
                const float m_est = (mr == -INFINITY ? 0.f : mr) + pn * emax;
                m_l2 = m_est * LOG2E;
            }
            // stage the excluded-key bits (padding mask or beyond L) and, for the second pass, the
            // rel-pos window of this tile (double buffered); both were requested one iteration ago
            float4* ew = ewin + acc * ATT_EWIN;
            uint32_t* mw = mwin + acc * 4;
            if (tid < 128) {
                const uint32_t bits = __ballot_sync(0xffffffffu, ex_cur);
                if (lane == 0) mw[tid >> 5] = bits;
            }
            if (pass && tid < 255)
                ew[tid] = make_float4(e_cur.x * LOG2E, e_cur.y * LOG2E, e_cur.z * LOG2E, e_cur.w * LOG2E);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (it + 1 < total_it) fetch(it + 1);         // in flight during this tile's arithmetic
            mbar_wait(&s_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * ATT_BN + 64 * half) +
                                   (static_cast<uint32_t>(quarter * 32) << 16);
            if (pass == 0) {
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t sr[32];
                    tmem_ld32(taddr + c0, sr);
                    tmem_ld_wait();
                    const uint32_t excl = mw[2 * half + (c0 >> 5)];
                    float cm = -INFINITY;
                    if (excl == 0u) {
#pragma unroll
                        for (int c = 0; c < 32; c += 2)
                            cm = fmaxf(fmaxf(cm, __uint_as_float(sr[c])), __uint_as_float(sr[c + 1]));
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            cm = fmaxf(cm, ((excl >> c) & 1u) ? -INFINITY : __uint_as_float(sr[c]));
                    }
                    m_run = fmaxf(m_run, cm);
                }
            } else {
                const int jh = j0 + 64 * half;            // first key column of this warp
                if (lane == 0) bulk_wait_read<0>();       // the previous box store has read the staging rows
                __syncwarp();
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    if (jh + 16 * g >= p.Lk) break;
                    uint32_t sr[16];
                    tmem_ld16(taddr + 16 * g, sr);
                    tmem_ld_wait();
                    const uint32_t excl = (mw[2 * half + (g >> 1)] >> (16 * (g & 1))) & 0xffffu;
                    const float4* ep = ew + (64 * half + 16 * g - r + 127);
                    float ls = 0.f;
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const float4 ea = ep[c], eb = ep[c + 1];
                        float a = fmaf(__uint_as_float(sr[c]), LOG2E, -m_l2);
                        float b = fmaf(__uint_as_float(sr[c + 1]), LOG2E, -m_l2);
                        a = fmaf(p0, ea.x, a); b = fmaf(p0, eb.x, b);
                        a = fmaf(p1, ea.y, a); b = fmaf(p1, eb.y, b);
                        a = fmaf(p2, ea.z, a); b = fmaf(p2, eb.z, b);
                        a = fmaf(p3, ea.w, a); b = fmaf(p3, eb.w, b);
                        sr[c] = __float_as_uint(fast_exp2(a));
                        sr[c + 1] = __float_as_uint(fast_exp2(b));
                    }
                    if (excl != 0u) {           // rare: a real branch, so unmasked tiles issue no selects
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if ((excl >> c) & 1u) sr[c] = 0u;
                    }
                    uint32_t w[8];
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const float a = __uint_as_float(sr[c]), b = __uint_as_float(sr[c + 1]);
                        w[c >> 1] = pack_h2(a, b);
                        ls += a + b;
                    }
                    l_run += ls;
                    *reinterpret_cast<uint4*>(my + (((2 * g) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(my + (((2 * g + 1) ^ sw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                }
                if (jh < p.Lk) {
                    fence_proxy_async_smem();             // generic-proxy writes -> visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&tma_p, stage, jh, i0 + quarter * 32, n * p.H + h);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[acc]);
        }
        xch[half * 128 + r] = l_run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0 && row_ok)
            p.inv_l[(static_cast<long long>(n) * p.H + h) * p.L + i] = 1.0f / (xch[r] + xch[128 + r]);
        if (lane == 0) bulk_wait_read<0>();               // staging must outlive the stores reading it
    }

    tc_fence_before();
    __syncthreads();
    if (warp == ATT_SM_WARPS + 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

}  // namespace zvb
