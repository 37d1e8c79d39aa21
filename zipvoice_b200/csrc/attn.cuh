// Attention-weights kernel (reference: modules/zipformer.py:1149-1306
// RelPositionMultiheadAttentionWeights.forward, eval path):
//   P[n,h,i,:] = softmax_j( q_i·k_j + p_i·E_h[j-i] , key-padding mask -> -1000 )   (fp32 softmax)
// q·kᵀ runs on tcgen05 (K = 32: two UMMA_K steps per 128x128 score tile, accumulators in
// TMEM, double buffered); the rel-pos bias is a 4-term dot product per score with the
// per-layer table E = linear_pos(pos_emb) staged in shared memory as the (i-tile, j-tile)
// window of 255 relative offsets.  Softmax is split so that every score costs one exp and one
// bias evaluation: pass 1 only takes the row max of q·kᵀ from TMEM (no bias, no exp); with
// m = max_j q·k_j + |p_i|·max_r|E_h[r]| >= every score of the row, pass 2 recomputes the cheap q·kᵀ,
// adds the bias and writes the UNNORMALISED weights 2^12·exp(s - m) as fp16 together with the fp32
// row sum's reciprocal; the three consumers of the layer (NonlinAttention, SelfAttention x2) apply 1/l
// in their GEMM epilogue.  Masked keys get exactly 0 (the reference's -1000 fill underflows as well).
// Pass 2 runs on packed half2: the (q·k - m)·log2e pair is formed in fp32 and converted once, the
// 4-term bias is 1 HMUL2 + 3 HFMA2 against the window stored as column PAIRS (one 16-byte LDS per two
// scores), the exponential is MUFU.EX2.F16 and its result IS the stored fp16 weight -- 7 instructions
// per score instead of 18 with fp32 arithmetic.  The bound m may exceed the true row max by up to
// 2·|p_i|·max|E|; the 2^12 head-room keeps the weights of such rows in fp16's normal range while that
// slack stays below ATT_BOUND_SLACK_L2 (checked per CTA); a CTA with a stronger rel-pos bias takes the
// EXACT row maximum instead: its first pass evaluates the bias too (same packed-half2 code).
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
constexpr int ATT_POS_PAD = 128;     // zero entries on both sides of the rel-pos table (weights.py: POS_PAD)
constexpr int ATT_EWIN_BYTES = 255 * 16;
constexpr float ATT_BOUND_SLACK_L2 = 10.0f;  // log2 units of slack the cheap softmax shift may have (weights >= 2^2)
constexpr int ATT_STAGE_BYTES = ATT_SM_WARPS * 4096; // per softmax warp: 32 rows x 128 B TMA-store staging
constexpr int ATT_MAX_SPLIT = 4;      // key tiles of one (query tile, head, utterance) over up to 4 CTAs of a cluster
constexpr int ATT_XPEER_OFF = (1 + ATT_KSTAGES) * ATT_TILE_BYTES + ATT_STAGE_BYTES + 2 * ATT_EWIN * 16 + 2 * 128 * 4 + 256;
constexpr int ATT_SMEM_BYTES = ATT_XPEER_OFF + 2 * ATT_MAX_SPLIT * 128 * 4 + 1024;   // two CTAs per SM: <= 113 KB each

struct AttnParams {
    int L, Lk, H, N;
    int qd;                          // H * 32: column of head-0 keys inside a qkp row
    const __half* qkp;        // [N*L, ld] = [q | k | p]
    int ld;
    const uint4* Epair;              // [H][2L-1 + 2*ATT_POS_PAD] fp16 column pairs (weights.py: pack_pos_table)
    const float* emax;               // [H] max_r |E[h][r]|_2
    const uint32_t* maskw;           // [N][mask_words] bit j%32 of word j/32 set = key j excluded (padded or >= L)
    int mask_words;                  // 4 * ceil(L / 128)
    __half* P;                // [N][H][L][Lk] unnormalised weights exp(s - m) (written through tma_p)
    float* inv_l;                    // [N][H][L]     1 / row sum
};

// tma_qk: [q | k | p] rows, box 64 x 128; tma_p: P viewed as (Lk, L, N*H), box 64 columns x 32 rows.
// CS > 1 (small grids: single utterances leave half of the SMs idle and every CTA walks 2 x num_jt tile iterations of ~1.2 us):
// a cluster of CS CTAs shares one (query tile, head, utterance), CTA `rank` takes key tiles [rank num_jt / CS, (rank + 1) num_jt /
// CS); the row maxima are exchanged between the passes and the row sums at the end through distributed shared memory
// (st.shared::cluster + a release arrive on the peer's mbarrier), so all CTAs of the cluster use the same shift m and
// rank 0 writes 1 / l.  Host: launch_op picks CS from the grid size and num_jt >= CS.
template <int CS>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_weights_kernel(const __grid_constant__ CUtensorMap tma_qk, const __grid_constant__ CUtensorMap tma_p,
                    const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* q_tile = smem;
    uint8_t* k_tiles = smem + ATT_TILE_BYTES;
    uint8_t* stage_all = smem + (1 + ATT_KSTAGES) * ATT_TILE_BYTES;                        // [8][32][128 B]
    // rel-pos window of the (i-tile, j-tile), pre-scaled by log2 e, as fp16 column pairs: entry e holds
    // {E[e][d], E[e+1][d]} for d = 0..3, so the pair of columns (c, c+1) of row r reads entry c - r + 127
    uint4* ewin = reinterpret_cast<uint4*>(stage_all + ATT_STAGE_BYTES);                    // [2][256]
    float* xch = reinterpret_cast<float*>(ewin + 2 * ATT_EWIN);                             // [2][128] half <-> half
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 256 + 8);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                      // [KSTAGES]
    uint64_t* k_empty = k_full + ATT_KSTAGES;         // [KSTAGES]
    uint64_t* s_full = k_empty + ATT_KSTAGES;         // [2]
    uint64_t* s_empty = s_full + 2;                   // [2]
    uint64_t* e_full = s_empty + 2;                   // [2] rel-pos window landed
    uint64_t* e_empty = e_full + 2;                   // [2] rel-pos window consumed (second pass only)
    uint64_t* x_bar = e_empty + 2;                    // [2] peers' row maxima / row sums landed (CS > 1)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(x_bar + 2);
    float* xpeer = reinterpret_cast<float*>(smem + ATT_XPEER_OFF);     // [2][ATT_MAX_SPLIT][128] written by the peers

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int i0 = (blockIdx.x / CS) * ATT_BM;
    const int h = blockIdx.y;
    const int n = blockIdx.z;
    const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
    const int all_jt = (p.L + ATT_BN - 1) / ATT_BN;
    const int jt_lo = CS > 1 ? static_cast<int>(rank) * all_jt / CS : 0;             // this CTA's key tiles
    const int num_jt = CS > 1 ? (static_cast<int>(rank) + 1) * all_jt / CS - jt_lo : all_jt;
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
            mbar_init(&e_full[s], 1);
            mbar_init(&e_empty[s], ATT_SM_WARPS);
            mbar_init(&x_bar[s], CS > 1 ? 128 * (CS - 1) : 1);
        }
        fence_barrier_init();
    }
    if (warp == ATT_SM_WARPS + 1) {
        tmem_alloc(tmem_holder, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();     // the peers' barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                 // set-up above overlaps the previous kernel's tail
    pdl_launch();

    // The cheap softmax shift m = max_j q.k_j + |p_i| max_r |E[r]| can exceed the true row maximum by up to
    // 2 |p_i| max|E|; the stored weights 2^12 exp(s - m) stay normal fp16 numbers while that slack is below
    // ~2^26.  A CTA whose rows may exceed ATT_BOUND_SLACK_L2 log2 units of slack (strong rel-pos bias)
    // switches -- CTA-uniformly, the rel-pos window in shared memory is common to its eight warps -- to the
    // EXACT row maximum: pass 1 then evaluates the bias as well (packed half2, like pass 2).
    // The decision is taken by the softmax warps from the p rows they load anyway and shared with the TMA
    // producer through `flag` and named barrier 2 (the producer has the Q tile and the first K tiles in
    // flight by then).
    float* flag = xch + 256;                              // [8] per-warp max |p_i|

    if (warp == ATT_SM_WARPS) {                           // TMA producer
        int stage = 0;
        uint32_t phase = 0;
        auto load_k = [&](int it) {
            const int jt = jt_lo + (it >= num_jt ? it - num_jt : it);
            mbar_wait(&k_empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&k_full[stage], ATT_TILE_BYTES);
            tma_load_3d(k_tiles + stage * ATT_TILE_BYTES, &tma_qk, &k_full[stage], p.qd + h * 32,
                        jt * ATT_BN, n);
            if (++stage == ATT_KSTAGES) { stage = 0; phase ^= 1u; }
        };
        const int prefill = total_it < ATT_KSTAGES ? total_it : ATT_KSTAGES;
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(q_tile, &tma_qk, q_full, h * 32, i0, n);
            for (int it = 0; it < prefill; ++it) load_k(it);       // in flight while the softmax warps decide
        }
        __syncwarp();
        asm volatile("bar.sync 2, 288;" ::: "memory");             // whole warp, converged
        if (lane == 0) {
            float pmax = flag[0];
            for (int w = 1; w < ATT_SM_WARPS; ++w) pmax = fmaxf(pmax, flag[w]);
            const bool exact_max = 2.0f * pmax * __ldg(p.emax + h) * 1.4426950408889634f > ATT_BOUND_SLACK_L2;
            const int e_first = exact_max ? 0 : num_jt;            // first iteration that needs the rel-pos window
            const uint4* Eh = p.Epair + static_cast<long long>(h) * (2 * p.L - 1 + 2 * ATT_POS_PAD);
            for (int it = 0; it < total_it; ++it) {
                const int jt = jt_lo + (it >= num_jt ? it - num_jt : it);
                if (it >= prefill) load_k(it);
                if (it >= e_first) {
                    // rel-pos window of this (i-tile, j-tile): 255 consecutive pair entries, first offset
                    // (j0 - i0) - 127, into buffer it & 1 (its own full/empty pair: this thread runs up to
                    // three tiles ahead of the softmax warps, so it cannot share the accumulator hand-shake)
                    const int eb = it & 1;
                    mbar_wait(&e_empty[eb], (static_cast<uint32_t>((it - e_first) >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&e_full[eb], ATT_EWIN_BYTES);
                    bulk_load_1d(ewin + eb * ATT_EWIN, Eh + ((jt * ATT_BN - i0) - 127 + (p.L - 1) + ATT_POS_PAD),
                                 ATT_EWIN_BYTES, &e_full[eb]);
                }
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
        uint32_t ph0 = 0u, ph1 = 0u, ph2 = 0u, ph3 = 0u;          // p_d duplicated into both halves
        if (row_ok) {
            const __half* pp = p.qkp + (static_cast<long long>(n) * p.L + i) * p.ld + 2 * p.qd + h * 4;
            const uint2 w = *reinterpret_cast<const uint2*>(pp);
            p0 = h2_lo(w.x); p1 = h2_hi(w.x); p2 = h2_lo(w.y); p3 = h2_hi(w.y);
            ph0 = __byte_perm(w.x, 0u, 0x1010); ph1 = __byte_perm(w.x, 0u, 0x3232);
            ph2 = __byte_perm(w.y, 0u, 0x1010); ph3 = __byte_perm(w.y, 0u, 0x3232);
        }
        {   // CTA-uniform choice between the cheap and the exact softmax shift (see above)
            float bi = sqrtf(p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3);
#pragma unroll
            for (int q = 16; q > 0; q >>= 1) bi = fmaxf(bi, __shfl_xor_sync(0xffffffffu, bi, q));
            if (lane == 0) flag[warp] = bi;
        }
        asm volatile("bar.sync 2, 288;" ::: "memory");
        float pmax_cta = flag[0];
#pragma unroll
        for (int w = 1; w < ATT_SM_WARPS; ++w) pmax_cta = fmaxf(pmax_cta, flag[w]);
        const bool exact_max = 2.0f * pmax_cta * __ldg(p.emax + h) * 1.4426950408889634f > ATT_BOUND_SLACK_L2;
        const int e_first = exact_max ? 0 : num_jt;       // first iteration that needs the rel-pos window
        // excluded-key bits of this warp's 64 columns of tile jt: words 2*half, 2*half + 1 of the tile's four
        const uint32_t* mwrow = p.maskw + static_cast<long long>(n) * p.mask_words + 2 * half;
        // P leaves through TMA stores: every warp stages 32 rows x 64 columns (128-byte rows, 128B
        // swizzle) and one lane issues the box store; rows >= L and columns >= Lk are clipped by the map
        uint8_t* stage = stage_all + (tid >> 5) * 4096;
        uint8_t* my = stage + lane * 128;
        const int sw = lane & 7;
        constexpr float LOG2E = 1.4426950408889634f;
        float m_run = -INFINITY, m_l2 = 0.f, l_run = 0.f;
        // excluded-key words one tile ahead: two CTAs of ~106 KB leave the L1 almost no capacity, so every global load
        // here is an L2 round trip; requested a whole tile before it is needed
        uint32_t mw0n = __ldg(mwrow + 4 * jt_lo), mw1n = __ldg(mwrow + 4 * jt_lo + 1);
        for (int it = 0; it < total_it; ++it) {
            const int pass = it >= num_jt ? 1 : 0;
            const int jt = jt_lo + (pass ? it - num_jt : it);
            const int j0 = jt * ATT_BN;
            const int acc = it & 1;
            const uint32_t acc_phase = static_cast<uint32_t>(it >> 1) & 1u;
            if (it == num_jt) {                           // between the passes: m >= every score of the row
                xch[half * 128 + r] = m_run;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                float mr = fmaxf(xch[r], xch[128 + r]);
                if (CS > 1) {                             // the row maximum over ALL key tiles: every CTA of the cluster must shift alike
                    if (half == 0) {
#pragma unroll
                        for (uint32_t pr = 0; pr < static_cast<uint32_t>(CS); ++pr)
                            if (pr != rank) {
                                st_f32_remote(xpeer + rank * 128 + r, pr, mr);
                                mbar_arrive_remote_release(&x_bar[0], pr);
                            }
                    }
                    mbar_wait_acquire_cluster(&x_bar[0], 0u);
#pragma unroll
                    for (uint32_t pr = 0; pr < static_cast<uint32_t>(CS); ++pr)
                        if (pr != rank) mr = fmaxf(mr, xpeer[pr * 128 + r]);
                }
                if (exact_max) {                          // exact (fp16) row maximum of the biased scores, log2 units
                    m_l2 = (mr == -INFINITY ? 0.f : mr) - 12.0f;
                } else {
                    const float emax = __ldg(p.emax + h);
                    const float pn = sqrtf(p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3);
                    const float m_est = (mr == -INFINITY ? 0.f : mr) + pn * emax;
                    m_l2 = m_est * LOG2E - 12.0f;        // weights are stored scaled by 2^12
                }
            }
            const uint4* ew = ewin + acc * ATT_EWIN;
            const uint32_t mw0 = mw0n, mw1 = mw1n;
            if (it + 1 < total_it) {
                const int jn = jt_lo + (it + 1 >= num_jt ? it + 1 - num_jt : it + 1);
                mw0n = __ldg(mwrow + 4 * jn);
                mw1n = __ldg(mwrow + 4 * jn + 1);
            }
            if (it >= e_first) mbar_wait(&e_full[acc], static_cast<uint32_t>((it - e_first) >> 1) & 1u);
            mbar_wait(&s_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * ATT_BN + 64 * half) +
                                   (static_cast<uint32_t>(quarter * 32) << 16);
            if (pass == 0 && exact_max) {
                // exact row maximum: (q.k) log2e + bias per column pair in packed half2, as in pass 2
                uint32_t hm = 0xFC00FC00u;                // (-inf, -inf)
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t sr[32];
                    tmem_ld32(taddr + c0, sr);
                    tmem_ld_wait();
                    const uint32_t excl = c0 ? mw1 : mw0;
                    const uint4* ep = ew + (64 * half + c0 - r + 127);
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        const uint4 e4 = ep[c];
                        uint32_t bias = hmul2(ph0, e4.x);
                        bias = hfma2(ph1, e4.y, bias);
                        bias = hfma2(ph2, e4.z, bias);
                        bias = hfma2(ph3, e4.w, bias);
                        uint32_t v = hadd2(bias, pack_h2(__uint_as_float(sr[c]) * LOG2E, __uint_as_float(sr[c + 1]) * LOG2E));
                        if (excl != 0u) {
                            if ((excl >> c) & 1u) v = (v & 0xFFFF0000u) | 0x0000FC00u;
                            if ((excl >> (c + 1)) & 1u) v = (v & 0x0000FFFFu) | 0xFC000000u;
                        }
                        hm = hmax2(hm, v);
                    }
                }
                m_run = fmaxf(m_run, fmaxf(h2_lo(hm), h2_hi(hm)));
            } else if (pass == 0) {
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t sr[32];
                    tmem_ld32(taddr + c0, sr);
                    tmem_ld_wait();
                    const uint32_t excl = c0 ? mw1 : mw0;
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
                if (jh + 64 <= p.Lk && (mw0 | mw1) == 0u) {
                    // interior tile (the common case): all 64 columns inside the padded width, no excluded key --
                    // straight-line code, no per-group bounds or mask selects
                    const f32x2 l2e2 = pack2(LOG2E, LOG2E), nm2 = pack2(-m_l2, -m_l2);
#pragma unroll 1
                    for (int gg = 0; gg < 2; ++gg) {
                        uint32_t sr32[32];
                        tmem_ld32(taddr + 32 * gg, sr32);
                        tmem_ld_wait();
                        const uint4* ep = ew + (64 * half + 32 * gg - r + 127);
                        uint32_t w[16];
#pragma unroll
                        for (int c = 0; c < 32; c += 2) {
                            const uint4 e4 = ep[c];
                            float a, b;                   // (q.k) log2e - m for the column pair: one packed FFMA2
                            unpack2(fma2(pack2(__uint_as_float(sr32[c]), __uint_as_float(sr32[c + 1])), l2e2, nm2), a, b);
                            uint32_t bias = hmul2(ph0, e4.x);
                            bias = hfma2(ph1, e4.y, bias);
                            bias = hfma2(ph2, e4.z, bias);
                            bias = hfma2(ph3, e4.w, bias);
                            w[c >> 1] = ex2_h2(hadd2(bias, pack_h2(a, b)));
                        }
                        // half2 partial row sums per 16 columns (<= 8 x 2^12 per half), then fp32
                        const uint32_t sa = hadd2(hadd2(hadd2(w[0], w[1]), hadd2(w[2], w[3])),
                                                  hadd2(hadd2(w[4], w[5]), hadd2(w[6], w[7])));
                        const uint32_t sb = hadd2(hadd2(hadd2(w[8], w[9]), hadd2(w[10], w[11])),
                                                  hadd2(hadd2(w[12], w[13]), hadd2(w[14], w[15])));
                        l_run += (h2_lo(sa) + h2_hi(sa)) + (h2_lo(sb) + h2_hi(sb));
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            *reinterpret_cast<uint4*>(my + (((4 * gg + q4) ^ sw) << 4)) =
                                make_uint4(w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
                    }
                } else {
#pragma unroll 1
                for (int gg = 0; gg < 2; ++gg) {
                    if (jh + 32 * gg >= p.Lk) break;
                    uint32_t sr32[32];
                    tmem_ld32(taddr + 32 * gg, sr32);            // one exposed TMEM latency per 32 columns
                    tmem_ld_wait();
                    const uint32_t excl32 = gg ? mw1 : mw0;
#pragma unroll
                for (int gl = 0; gl < 2; ++gl) {
                    const int g = 2 * gg + gl;
                    if (jh + 16 * g >= p.Lk) break;
                    const uint32_t* sr = sr32 + 16 * gl;
                    const uint32_t excl = (excl32 >> (16 * gl)) & 0xffffu;
                    const uint4* ep = ew + (64 * half + 16 * g - r + 127);
                    uint32_t w[8];
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const uint4 e4 = ep[c];
                        const float a = fmaf(__uint_as_float(sr[c]), LOG2E, -m_l2);
                        const float b = fmaf(__uint_as_float(sr[c + 1]), LOG2E, -m_l2);
                        uint32_t bias = hmul2(ph0, e4.x);
                        bias = hfma2(ph1, e4.y, bias);
                        bias = hfma2(ph2, e4.z, bias);
                        bias = hfma2(ph3, e4.w, bias);
                        const uint32_t pw = ex2_h2(hadd2(bias, pack_h2(a, b)));
                        w[c >> 1] = pw;
                    }
                    if (excl != 0u) {           // rare: a real branch, so unmasked tiles issue no selects
#pragma unroll
                        for (int c = 0; c < 16; c += 2) {
                            const uint32_t keep = (((excl >> c) & 1u) ? 0u : 0x0000FFFFu) |
                                                  (((excl >> (c + 1)) & 1u) ? 0u : 0xFFFF0000u);
                            w[c >> 1] &= keep;
                        }
                    }
                    // half2 partial row sum of the group (<= 8 x 2^12 per half), then fp32
                    const uint32_t sum2 = hadd2(hadd2(hadd2(w[0], w[1]), hadd2(w[2], w[3])),
                                                hadd2(hadd2(w[4], w[5]), hadd2(w[6], w[7])));
                    l_run += h2_lo(sum2) + h2_hi(sum2);
                    *reinterpret_cast<uint4*>(my + (((2 * g) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(my + (((2 * g + 1) ^ sw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                }
                }
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
            if (lane == 0) {
                mbar_arrive(&s_empty[acc]);
                if (it >= e_first) mbar_arrive(&e_empty[acc]);
            }
        }
        xch[half * 128 + r] = l_run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (CS > 1 && half == 0) {                        // row sums of the other key ranges -> rank 0
            if (rank != 0u) {
                st_f32_remote(xpeer + (ATT_MAX_SPLIT + rank) * 128 + r, 0u, xch[r] + xch[128 + r]);
                mbar_arrive_remote_release(&x_bar[1], 0u);
            } else {
                mbar_wait_acquire_cluster(&x_bar[1], 0u);
            }
        }
        if (half == 0 && row_ok && rank == 0u)
        {
            float l = xch[r] + xch[128 + r];
            if (CS > 1) {
#pragma unroll
                for (int pr = 1; pr < CS; ++pr) l += xpeer[(ATT_MAX_SPLIT + pr) * 128 + r];
            }
            p.inv_l[(static_cast<long long>(n) * p.H + h) * p.L + i] = l > 0.f ? 1.0f / l : 0.f;
        }
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
