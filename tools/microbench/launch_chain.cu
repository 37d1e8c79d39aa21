// What does one kernel of a dependent chain cost inside a replayed CUDA graph on B200?  The single-utterance sampler (C1) runs
// 398 kernels per decoder forward at ~9 us each for ~1-3 us of streaming work; this measures the floor of such a chain for
// kernels shaped like the library's GEMM (640 threads, 227 KB of dynamic shared memory, 512 TMEM columns, mbarrier set-up,
// griddepcontrol.wait / launch_dependents) with a dependent global round trip of growing realism.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_chain.bin launch_chain.cu && ./launch_chain.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// MODE bit 0: TMEM alloc / dealloc + mbarrier init + block barrier (the GEMM's set-up)
// MODE bit 1: dependent data path: bulk-load 16 KB the previous kernel wrote, wait, bulk-store it to this kernel's output
// MODE bit 2: plain dependent path instead: one 16-byte global load per thread of the previous output, store to own output
template <int MODE>
__global__ void __launch_bounds__(640, 1) chain_kernel(const uint4* __restrict__ in, uint4* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (MODE & 1) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&holder)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    } else if ((MODE & 2) && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((MODE & 2) && !(MODE & 1)) __syncthreads();
    pdl_wait();
    pdl_launch();
    if (MODE & 2) {
        if (threadIdx.x == 0) {
            const uint32_t b = smem_u32(&bar), dst = smem_u32(smem);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(in + blockIdx.x * 1024), "r"(16384), "r"(b) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(b) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + blockIdx.x * 1024), "r"(dst), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else if (MODE & 4) {
        const uint4 v = in[blockIdx.x * 1024 + threadIdx.x];
        out[blockIdx.x * 1024 + threadIdx.x] = v;
    } else if (threadIdx.x == 0) {
        out[blockIdx.x * 1024] = make_uint4(1, 2, 3, 4);
    }
    if (MODE & 1) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 1) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(holder));
        }
    }
    (void)lane;
}

__global__ void tiny_kernel(const uint4* in, uint4* out) {
    pdl_wait();
    pdl_launch();
    if (threadIdx.x == 0) out[blockIdx.x * 1024] = in[blockIdx.x * 1024];
}

template <typename K>
static int run(const char* name, K kern, int threads, int smem, int grid, bool pdl, uint4* a, uint4* b, int smem_alt = -1,
               K kern_alt = nullptr) {
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const int n = 400;
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < n; ++i) {
        cudaLaunchConfig_t cfg = {};
        const bool alt = smem_alt >= 0 && (i & 1);
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = alt ? smem_alt : smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        const uint4* in = (i & 1) ? b : a;
        uint4* out = (i & 1) ? a : b;
        CK(cudaLaunchKernelEx(&cfg, alt ? kern_alt : kern, in, out));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, st));
        CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    printf("%-92s %6.2f us per kernel\n", name, best * 1e3f / n);
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); CK(cudaStreamDestroy(st));
    return 0;
}

int main() {
    uint4 *a, *b;
    const size_t bytes = 148 * 1024 * sizeof(uint4);
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
    const int BIG = 232448 - 1024;   // the kernel also has 12 bytes of static shared memory
    CK(cudaFuncSetAttribute(chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG));
    CK(cudaFuncSetAttribute(chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG));
    CK(cudaFuncSetAttribute(chain_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG));
    CK(cudaFuncSetAttribute(chain_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG));
    CK(cudaFuncSetAttribute(chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG));
    CK(cudaFuncSetAttribute(chain_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG));
    for (int pdl = 0; pdl < 2; ++pdl) {
        printf("--- %s\n", pdl ? "programmatic dependent launch" : "plain stream order");
        if (run("tiny kernel (32 threads, 148 CTAs, no shared memory)", tiny_kernel, 32, 0, 148, pdl, a, b)) return 1;
        if (run("640 threads, 16 KB smem, trivial store", chain_kernel<0>, 640, 16384, 148, pdl, a, b)) return 1;
        if (run("640 threads, 227 KB smem, trivial store", chain_kernel<0>, 640, BIG, 148, pdl, a, b)) return 1;
        if (run("640 threads, 227 KB smem, alternating with 100 KB smem", chain_kernel<0>, 640, BIG, 148, pdl, a, b, 102400, chain_kernel<0>)) return 1;
        if (run("+ TMEM alloc(512)/dealloc, mbarrier init, block barriers", chain_kernel<1>, 640, BIG, 148, pdl, a, b)) return 1;
        if (run("+ dependent 16-byte load/store per thread", chain_kernel<5>, 640, BIG, 148, pdl, a, b)) return 1;
        if (run("+ dependent 16 KB bulk load -> mbarrier -> bulk store (GEMM-like data path)", chain_kernel<3>, 640, BIG, 148, pdl, a, b)) return 1;
        if (run("   same, 20 CTAs", chain_kernel<3>, 640, BIG, 20, pdl, a, b)) return 1;
        if (run("bulk path without the TMEM set-up, 16 KB smem, 128 threads", chain_kernel<2>, 128, 16384 + 1024, 148, pdl, a, b)) return 1;
        if (run("   same, 227 KB smem, 640 threads", chain_kernel<2>, 640, BIG, 148, pdl, a, b)) return 1;
    }
    return 0;
}
