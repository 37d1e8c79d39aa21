// sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM), UMMA
// descriptors.  Everything the tensor-core kernels in this tree need, written against the
// PTX ISA directly (no CUTLASS dependency).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace zvb {

#ifndef ZVB_WAIT_TIMEOUT_NS
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU box.
#define ZVB_WAIT_TIMEOUT_NS 4000000000ull
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ------------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the plan is launched with programmatic stream serialization: its CTAs may become resident
// (and run their set-up: barrier init, TMEM allocation, tensor-map prefetch, constant staging) while the
// previous kernel drains.  pdl_wait() blocks until the previous grid has completed and its writes are visible;
// nothing written by a predecessor may be touched before it.  pdl_launch() lets the NEXT grid start launching.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the
// hint expires) instead of burning issue slots of the SMSP it shares with the math warps.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// non-blocking poll (test_wait never suspends the thread)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 63u) == 0u && globaltimer_ns() - t0 > ZVB_WAIT_TIMEOUT_NS) {
            printf("zvb: mbarrier wait timeout block=(%d,%d,%d) thread=%d parity=%u\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, threadIdx.x, parity);
            __trap();
        }
    }
}

// Variants on raw 32-bit shared-window addresses.  A generic pointer into shared memory is converted by
// `cvta.to.shared`, which on sm_100 re-reads %cluster_ctaid and rebuilds the CTA's window base every time
// (S2UR SR_CgaCtaId + 5 uniform ops in SASS): hot loops convert once and keep the u32.
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_u32(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait_u32(bar, parity)) {
        if ((++spins & 63u) == 0u && globaltimer_ns() - t0 > ZVB_WAIT_TIMEOUT_NS) {
            printf("zvb: mbarrier wait timeout block=(%d,%d,%d) thread=%d parity=%u\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128_u32(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128_u32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load global -> shared, completion signalled on `bar` (complete_tx::bytes).
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 1-D bulk copy global -> shared (no tensor map): `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 3-D tiled store shared -> global (bulk async group); out-of-range rows / columns of the box are
// clipped by the tensor map, so partial tiles need no predication.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still have to READ their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// L2 prefetch of a box (no shared memory, no completion): hides the HBM latency of operands that a
// later TMA load of the same box will then find in L2
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask` and
// completes `bytes` on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                               int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Distributed shared memory hand-off between the CTAs of a cluster (attention weights, key split): a float stored into
// the peer's shared memory, published by a release arrive (cluster scope) on the peer's mbarrier; the waiter acquires at
// cluster scope before it reads.
__device__ __forceinline__ void st_f32_remote(const float* local_addr, uint32_t rank, float v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.f32 [ra], %2;\n\t}"
        ::"r"(smem_u32(local_addr)), "r"(rank), "f"(v)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
        if (ok) return;
        if ((++spins & 63u) == 0u && globaltimer_ns() - t0 > ZVB_WAIT_TIMEOUT_NS) {
            printf("zvb: cluster mbarrier wait timeout block=(%d,%d,%d) thread=%d\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}

// CTA-pair (cta_group::2) variants.  A shared::cluster address carries the CTA rank of the pair in
// bit 24; clearing it addresses the same offset in the leader CTA (rank 0).
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
// TMA load executed by either CTA of a pair into its OWN shared memory, completing bytes on the
// LEADER's mbarrier (only the leader's MMA thread waits for operands).
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK),
        "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster.  Relaxed: the arrival only says
// "this warp's tcgen05.ld of the accumulator have completed" (they were waited for and fenced); no generic
// memory is published, and a release here costs a cluster-wide fence (ERRBAR) per warp and tile.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_holder)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// D[tmem of both CTAs, 256 rows] (+)= A[128 rows in each CTA's smem] * B[N/2 rows in each CTA's smem]^T,
// issued by the leader CTA only; descriptors hold the leader's addresses (same offsets in the peer).
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once the pair's MMAs have retired) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}

// ------------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_holder)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate, single-CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask` (smem slot shared by a cluster)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// K-major operand tile in shared memory, rows of 128 bytes (64 fp16), 128B swizzle as written
// by a TMA load with CU_TENSOR_MAP_SWIZZLE_128B into a 1024B-aligned buffer.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused: 1)
//   bits [32,46) stride byte offset >> 4 = 1024 >> 4 (next group of 8 rows)
//   bits [46,48) descriptor version = 1      bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=fp16 (formats, bits 7-9 / 10-12 = 0),
// both K-major, M=m, N=n.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t n, uint32_t m = 128) {
    return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ------------------------------------------------------------------------------- math
// Packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100): two fp32 lanes in one 64-bit register
// pair, one instruction -- halves the issue slots of the CUDA-core inner loops.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Swoosh(x) = softplus(x - c) - 0.08 x - d with softplus(y) = max(y,0) + log1p(exp(-|y|)).
// One MUFU op (ex2) per activation instead of two (ex2 + lg2): log1p on [0,1] is a degree-5
// polynomial (max abs error 1.2e-5), and max(y,0) - 0.08 y is folded into 0.42 y + 0.5 |y|, so the
// whole activation is 1 FMUL + 1 MUFU + 8 FFMA on y = x - c:
//   swoosh = t*P(t) + 0.42 y + 0.5 |y| - (0.08 c + d),   t = 2^(-|y| log2 e)
__device__ __forceinline__ float swoosh_from_offset(float y, float k0) {
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-fabsf(y) * 1.4426950408889634f));
    float p = fmaf(t, 0.031377589387161245f, -0.1341354334221127f);
    p = fmaf(t, p, 0.2878262894239249f);
    p = fmaf(t, p, -0.491347927069251f);
    p = fmaf(t, p, 0.9994349844843187f);
    float r = fmaf(t, p, k0);
    r = fmaf(y, 0.42f, r);
    return fmaf(fabsf(y), 0.5f, r);
}
// Same activation evaluated from x directly (no separate x - c): z = (x - c) log2 e is one FFMA, the
// |.| and the negation ride on the MUFU operand, and the linear terms are re-expressed in x and |z|.
__device__ __forceinline__ float swoosh_direct(float x, float c, float k0) {
    constexpr float L2E = 1.4426950408889634f;
    const float z = fmaf(x, L2E, -c * L2E);
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-fabsf(z)));
    float p = fmaf(t, 0.031377589387161245f, -0.1341354334221127f);
    p = fmaf(t, p, 0.2878262894239249f);
    p = fmaf(t, p, -0.491347927069251f);
    p = fmaf(t, p, 0.9994349844843187f);
    float r = fmaf(t, p, k0 - 0.42f * c);
    r = fmaf(x, 0.42f, r);
    return fmaf(fabsf(z), 0.5f / L2E, r);
}
// swoosh_direct on a pair: 8 FFMA2 + 2 MUFU.  The fp32 FMA pipe issues one 3-register FFMA per two
// cycles and scheduler, an FFMA2 in the same two cycles -- packing doubles what the epilogue warps get out
// of it (tools/microbench/epi_math.cu: 5.8 -> 7.8 activations/clk/SM at 8 warps).
__device__ __forceinline__ void swoosh_direct2(float& x0, float& x1, float c, float k0) {
    constexpr float L2E = 1.4426950408889634f;
    const f32x2 x = pack2(x0, x1);
    const f32x2 z = fma2(x, pack2(L2E, L2E), pack2(-c * L2E, -c * L2E));
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
    const float k = k0 - 0.42f * c;
    q = fma2(t, q, pack2(k, k));
    q = fma2(x, pack2(0.42f, 0.42f), q);
    q = fma2(pack2(a0, a1), pack2(0.5f / L2E, 0.5f / L2E), q);
    unpack2(q, x0, x1);
}
// the same on a packed pair, in and out
__device__ __forceinline__ f32x2 swoosh_x2(f32x2 x, float c, float k0) {
    constexpr float L2E = 1.4426950408889634f;
    const f32x2 z = fma2(x, pack2(L2E, L2E), pack2(-c * L2E, -c * L2E));
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
    const float k = k0 - 0.42f * c;
    q = fma2(t, q, pack2(k, k));
    q = fma2(x, pack2(0.42f, 0.42f), q);
    return fma2(pack2(a0, a1), pack2(0.5f / L2E, 0.5f / L2E), q);
}
// Scheduling-pinned variants: ptxas keeps `asm volatile` statements in program order, so chains interleaved in the
// source stay interleaved in SASS (left alone it serialises every chain to save registers).
__device__ __forceinline__ f32x2 fma2_o(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ex2_neg_abs_o(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(-fabsf(x)));
    return y;
}
// G pairs at once, the G Horner chains interleaved step by step: a single chain is a string of dependent FFMA2s
// behind a MUFU (4-cycle issue-to-use each, ~20 for the MUFU), and an epilogue warp shares its scheduler with only
// three others -- the independent chains are what keeps the FMA pipe fed.
template <int G>
__device__ __forceinline__ void swoosh_x2_group(f32x2* v, float c, float k0) {
    constexpr float L2E = 1.4426950408889634f;
    f32x2 z[G], t[G], q[G];
#pragma unroll
    for (int i = 0; i < G; ++i) z[i] = fma2_o(v[i], pack2(L2E, L2E), pack2(-c * L2E, -c * L2E));
#pragma unroll
    for (int i = 0; i < G; ++i) {
        float z0, z1;
        unpack2(z[i], z0, z1);
        const float t0 = ex2_neg_abs_o(z0);
        const float t1 = ex2_neg_abs_o(z1);
        t[i] = pack2(t0, t1);
    }
#pragma unroll
    for (int i = 0; i < G; ++i)
        q[i] = fma2_o(t[i], pack2(0.031377589387161245f, 0.031377589387161245f), pack2(-0.1341354334221127f, -0.1341354334221127f));
#pragma unroll
    for (int i = 0; i < G; ++i) q[i] = fma2_o(t[i], q[i], pack2(0.2878262894239249f, 0.2878262894239249f));
#pragma unroll
    for (int i = 0; i < G; ++i) q[i] = fma2_o(t[i], q[i], pack2(-0.491347927069251f, -0.491347927069251f));
#pragma unroll
    for (int i = 0; i < G; ++i) q[i] = fma2_o(t[i], q[i], pack2(0.9994349844843187f, 0.9994349844843187f));
    const float k = k0 - 0.42f * c;
#pragma unroll
    for (int i = 0; i < G; ++i) q[i] = fma2_o(t[i], q[i], pack2(k, k));
#pragma unroll
    for (int i = 0; i < G; ++i) q[i] = fma2_o(v[i], pack2(0.42f, 0.42f), q[i]);
#pragma unroll
    for (int i = 0; i < G; ++i) {
        float z0, z1;
        unpack2(z[i], z0, z1);
        v[i] = fma2_o(pack2(fabsf(z0), fabsf(z1)), pack2(0.5f / L2E, 0.5f / L2E), q[i]);
    }
}
// exact GELU, 0.5 x (1 + erf(x / sqrt 2)) (the vocoder's ConvNeXt blocks: torch.nn.GELU default)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
constexpr float SWOOSH_L_C = 4.0f, SWOOSH_L_K0 = -(0.08f * 4.0f + 0.035f);
constexpr float SWOOSH_R_C = 1.0f, SWOOSH_R_K0 = -(0.08f * 1.0f + 0.313261687f);
__device__ __forceinline__ float swoosh_l(float x) {
    // log(1+exp(x-4)) - 0.08x - 0.035 (reference: modules/scaling.py:1189-1195)
    return swoosh_from_offset(x - SWOOSH_L_C, SWOOSH_L_K0);
}
__device__ __forceinline__ float swoosh_r(float x) {
    // log(1+exp(x-1)) - 0.08x - 0.313261687 (reference: modules/scaling.py:1200-1206)
    return swoosh_from_offset(x - SWOOSH_R_C, SWOOSH_R_K0);
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) {
    return __fdividef(1.0f, 1.0f + __expf(-x));
}
// 16-bit storage type of every activation / weight: IEEE fp16 (11-bit significand; the residual stream
// is kept in it, so its rounding -- not bf16's 8 bits -- bounds the per-module accumulation error).
// Conversions saturate to +-65504 instead of producing inf.
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// packed half2 arithmetic on raw 32-bit registers
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
    uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
    uint32_t d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) {
    uint32_t d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {      // 2^x on both halves (MUFU.EX2.F16)
    uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y;
}
__device__ __forceinline__ __half f2h(float v) {
    const uint32_t w = pack_h2(v, 0.0f);
    return __ushort_as_half(static_cast<unsigned short>(w & 0xFFFFu));
}
__device__ __forceinline__ float h2_lo(uint32_t w) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xFFFFu)));
}
__device__ __forceinline__ float h2_hi(uint32_t w) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16)));
}

}  // namespace zvb
