// Host side of the C ABI (include/zipvoice_b200.h): builds a static launch plan for one
// TTSZipformer over (N rows, T frames) -- TMA tensor maps, tile shapes, workspace carving --
// and replays it on a stream.  No allocation, no synchronisation, graph capturable.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zipvoice_b200.h"
#include "attn.cuh"
#include "attn3.cuh"
#include "gemm.cuh"
namespace zvb { constexpr int ACT_SWOOSH_R_ = 2; }
#include "elementwise.cuh"
#include "audio.cuh"

using namespace zvb;
typedef __half h16;

// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local long long g_launches = 0;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) return fail(ZVB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int r_ = (expr);          \
        if (r_ != 0) return r_;   \
    } while (0)

static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ZVB_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
    ++g_launches;
    return 0;
}

static int g_num_sms = 0;
static int g_pdl = 1;             // ZVB_NO_PDL=1: plain stream serialization between the kernels of a plan

// Every kernel goes through here: programmatic dependent launch (ptx.cuh: pdl_wait / pdl_launch).
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
static int g_cluster_ok = 1;      // ZVB_NO_CLUSTER=1 disables the CTA-pair (cta_group::2) GEMM variant
static long long g_pair_min_mtiles = -1;    // ZVB_PAIR_MIN_MTILES: fewest 128-row tiles (all batches) for CTA pairs; default 2 x SMs
static int g_tma_store_ok = 1;
static int g_wide_pref = 1;       // ZVB_WIDE_PREF=0: exact-fit tile widths first
static double g_wide_waste = 0.12; // ZVB_WIDE_WASTE: largest padding share accepted for a 256-column tile when K <= 512
static int g_bn192 = 0;           // ZVB_BN192=1: 192-column tiles for short-K GEMMs (measured: no gain)
static int g_layout_ok = 1;       // ZVB_NO_LAYOUT=1 keeps the default operand-ring / aux split everywhere
static int g_attn_split = 1;      // ZVB_ATTN_SPLIT=0: no key split of the attention weights over a cluster; 3 / 4: force 2 / 4 CTAs (tests)
static int g_attn_tc = 0;         // ZVB_ATTN_TC=1: attention weights with the tensor-core rel-pos bias (attn3.cuh; measured slower)
static int g_dw_mode = 0;          // ZVB_DW_MODE: depthwise-convolution block shapes (elementwise.cuh: DwShape), 0 = measured best
static int g_fuse_prologue = 1;    // ZVB_NO_FUSED_PROLOGUE=1: masks, per-stack time projections and row biases as separate launches
static int g_pv_deep = 0;          // ZVB_PV_DEEP=1: SelfAttention P.V with a 10-stage operand ring over the unused staging area
static int g_pv_bn = 0;            // ZVB_PV_BN=<columns>: tile width of the NonlinAttention P.V GEMM (0 = two equal tiles)
static int g_merge_ff1 = 1;        // ZVB_NO_MERGE=1: feed_forward1 / attention in-projections as separate GEMMs
static int g_small_model = 1;      // ZVB_NO_SMALL_MODEL=1: round 1's tile-width choice for small problems
static int g_pre_b = 1;            // ZVB_NO_PRE_B=1: weight tiles are requested after the dependency wait like the activations
static thread_local bool g_plan_build = false;  // set while a plan is being built: B operands of build_linear / build_gated are model weights
struct PlanBuildScope { PlanBuildScope() { g_plan_build = true; } ~PlanBuildScope() { g_plan_build = false; } };
static int g_small_lean = 1;       // ZVB_SMALL_LEAN=0: the small-problem cost model without the measured epilogue costs
static int g_lean_pad = 1;         // ZVB_NO_LEAN_PAD=1: exact-fit tile widths for projections that are no multiple of 64 wide
static int g_fast_bypass = 1;      // ZVB_NO_FAST_BYPASS=1: generic epilogue for the bypass GEMM (feed_forward2)
static int g_fast_resid = 1;       // ZVB_NO_FAST_RESID=1: generic epilogue for the residual-stream GEMMs
static int g_fast_epi = 1;        // ZVB_NO_FAST_EPI=1: generic epilogue everywhere
static int g_resident_ok = 0;     // ZVB_RESIDENT=1: A-stationary tile order for the K = 512 GEMMs (measured 5-8% SLOWER, profiles/gemm_resident_ab_r2.txt)
static int g_pair_min_kb = 8;     // ZVB_PAIR_MIN_KB: fewest k-blocks for which a CTA pair is used    // ZVB_NO_TMA_STORE=1 keeps the epilogue on per-thread stores
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;


// every instantiation of the GEMM kernel: [epilogue kind][activation][CTA pair][lean path] (nullptr = not built)
typedef void (*gemm_fn_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                          const GemmParams);
static gemm_fn_t gemm_fn(int kind, int act, int cluster, int lean) {
#define ZVB_G(K, A, C, L) return gemm_kernel<K, A, C, L>
    if (kind == EPI_GATED) { if (cluster == 1) ZVB_G(EPI_GATED, ACT_NONE, 1, 0); ZVB_G(EPI_GATED, ACT_NONE, 2, 0); }
    if (lean == 2) {
        if (act != ACT_NONE) return nullptr;
        if (cluster == 1) ZVB_G(EPI_LINEAR, ACT_NONE, 1, 2);
        ZVB_G(EPI_LINEAR, ACT_NONE, 2, 2);
    }
    if (lean == 3) {
        if (act != ACT_NONE) return nullptr;
        if (cluster == 1) ZVB_G(EPI_LINEAR, ACT_NONE, 1, 3);
        ZVB_G(EPI_LINEAR, ACT_NONE, 2, 3);
    }
#define ZVB_G2(A, L) if (act == A && lean == L) { if (cluster == 1) ZVB_G(EPI_LINEAR, A, 1, L); ZVB_G(EPI_LINEAR, A, 2, L); }
    ZVB_G2(ACT_NONE, 0) ZVB_G2(ACT_SWOOSH_L, 0) ZVB_G2(ACT_SWOOSH_R, 0) ZVB_G2(ACT_GELU, 0)
    ZVB_G2(ACT_NONE, 1) ZVB_G2(ACT_SWOOSH_L, 1) ZVB_G2(ACT_SWOOSH_R, 1) ZVB_G2(ACT_GELU, 1)
#undef ZVB_G2
#undef ZVB_G
    return nullptr;
}

// Per-device initialisation: the opt-in shared-memory sizes (cudaFuncSetAttribute) and the SM count belong to
// the CURRENT device, so a process that builds plans on several GPUs initialises each of them once.
#ifndef ZVB_SOURCE_HASH
#define ZVB_SOURCE_HASH "unknown"
#endif
static const char g_source_hash[] = "ZVB_SRC_HASH=" ZVB_SOURCE_HASH;     // zipvoice_b200/build.py greps this marker
constexpr int ZVB_MAX_DEVICES = 64;
static std::mutex g_init_mutex;
static int g_dev_sms[ZVB_MAX_DEVICES] = {};

// Measurement switches (environment), read once per process -- before any host-only sizing call, so that
// zvb_plan_workspace_bytes and zvb_plan_create always agree on the plan's shape.
static std::once_flag g_switch_once;
static void load_switches() {
    std::call_once(g_switch_once, [] {
        if (const char* e = getenv("ZVB_NO_CLUSTER")) g_cluster_ok = atoi(e) == 0;
        if (const char* e = getenv("ZVB_NO_TMA_STORE")) g_tma_store_ok = atoi(e) == 0;
        if (const char* e = getenv("ZVB_PAIR_MIN_KB")) g_pair_min_kb = atoi(e);
        if (const char* e = getenv("ZVB_NO_LAYOUT")) g_layout_ok = atoi(e) == 0;
        if (const char* e = getenv("ZVB_NO_PDL")) g_pdl = atoi(e) == 0;
        if (const char* e = getenv("ZVB_RESIDENT")) g_resident_ok = atoi(e) != 0;
        if (const char* e = getenv("ZVB_NO_FAST_EPI")) g_fast_epi = atoi(e) == 0;
        if (const char* e = getenv("ZVB_NO_FAST_RESID")) g_fast_resid = atoi(e) == 0;
        if (const char* e = getenv("ZVB_NO_FAST_BYPASS")) g_fast_bypass = atoi(e) == 0;
        if (const char* e = getenv("ZVB_NO_LEAN_PAD")) g_lean_pad = atoi(e) == 0;
        if (const char* e = getenv("ZVB_NO_SMALL_MODEL")) g_small_model = atoi(e) == 0;
        if (const char* e = getenv("ZVB_SMALL_LEAN")) g_small_lean = atoi(e) != 0;
        if (const char* e = getenv("ZVB_NO_PRE_B")) g_pre_b = atoi(e) == 0;
        if (const char* e = getenv("ZVB_ATTN_SPLIT")) g_attn_split = atoi(e);
        if (const char* e = getenv("ZVB_PAIR_MIN_MTILES")) g_pair_min_mtiles = atoll(e);
        if (const char* e = getenv("ZVB_NO_MERGE")) g_merge_ff1 = atoi(e) == 0;
        if (const char* e = getenv("ZVB_PV_BN")) g_pv_bn = atoi(e);
        if (const char* e = getenv("ZVB_PV_DEEP")) g_pv_deep = atoi(e) != 0;
        if (const char* e = getenv("ZVB_NO_FUSED_PROLOGUE")) g_fuse_prologue = atoi(e) == 0;
        if (const char* e = getenv("ZVB_ATTN_TC")) g_attn_tc = atoi(e) != 0;
        if (const char* e = getenv("ZVB_BN192")) g_bn192 = atoi(e) != 0;
        if (const char* e = getenv("ZVB_DW_MODE")) g_dw_mode = atoi(e);
        if (const char* e = getenv("ZVB_WIDE_PREF")) g_wide_pref = atoi(e) != 0;
        if (const char* e = getenv("ZVB_WIDE_WASTE")) g_wide_waste = atof(e);
    });
}

static int init_device() {
    std::lock_guard<std::mutex> lock(g_init_mutex);
    int dev = 0, count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(ZVB_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= ZVB_MAX_DEVICES) return fail(ZVB_ERR_INVALID, "device ordinal %d out of range", dev);
    if (g_dev_sms[dev] != 0) { g_num_sms = g_dev_sms[dev]; return 0; }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(ZVB_ERR_NO_DEVICE, "device sm_%d%d is not sm_100 (B200)", prop.major, prop.minor);
    load_switches();
    if (g_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (fn == nullptr || q != cudaDriverEntryPointSuccess)
            return fail(ZVB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
#define ZVB_SMEM_ATTR(k) CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES))
    for (int lean = 0; lean < 4; ++lean)
        for (int act = 0; act < 4; ++act)
            for (int cl = 1; cl <= 2; ++cl)
                if (gemm_fn(EPI_LINEAR, act, cl, lean) != nullptr) ZVB_SMEM_ATTR(gemm_fn(EPI_LINEAR, act, cl, lean));
    ZVB_SMEM_ATTR(gemm_fn(EPI_GATED, ACT_NONE, 1, 0)); ZVB_SMEM_ATTR(gemm_fn(EPI_GATED, ACT_NONE, 2, 0));
#undef ZVB_SMEM_ATTR
    CUDA_TRY(cudaFuncSetAttribute(attn_weights_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(attn_weights_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(attn_weights_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(attn_weights_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<7, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<7, 0>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<7, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<7, 1>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<7, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<7, 2>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<9, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<9, 0>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<9, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<9, 1>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<9, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<9, 2>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<15, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<15, 0>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<15, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<15, 1>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<15, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<15, 2>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<31, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<31, 0>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<31, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<31, 1>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<31, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<31, 2>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<7, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<7, 0>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<7, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<7, 1>()));
    CUDA_TRY(cudaFuncSetAttribute(dwconv_kernel<7, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<7, 2>()));
    g_dev_sms[dev] = prop.multiProcessorCount;
    g_num_sms = prop.multiProcessorCount;
    return 0;
}

// Tensor (dim0 fastest) viewed as 3-D, box = (128 bytes, box1, 1) with 128B swizzle.
// Out-of-range elements are zero-filled on loads and clipped on stores.
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                     uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box1, bool f32 = false) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (stride1_bytes & 15) != 0 || (stride2_bytes & 15) != 0)
        return fail(ZVB_ERR_INVALID, "tensor map: pointer/strides must be 16-byte aligned");
    if (box1 == 0 || box1 > 256 || d0 == 0 || d1 == 0 || d2 == 0)
        return fail(ZVB_ERR_INVALID, "tensor map: bad box/dims (%u, %llu, %llu, %llu)", box1,
                    (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {f32 ? 32u : 64u, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                          const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ZVB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// h16 tensor, box = (32 elements = 64 bytes, box1, 1) with the 64-byte swizzle (attention-weight stores, attn3.cuh)
static int make_tmap_sw64(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                          uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box1) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (stride1_bytes & 15) != 0 || (stride2_bytes & 15) != 0)
        return fail(ZVB_ERR_INVALID, "tensor map: pointer/strides must be 16-byte aligned");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {32u, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ZVB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// h16 tensor, box = (box0 elements, box1, 1), no swizzle (rows of the box are contiguous in shared memory)
static int make_tmap_plain(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                           uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (stride1_bytes & 15) != 0 || (stride2_bytes & 15) != 0)
        return fail(ZVB_ERR_INVALID, "tensor map: pointer/strides must be 16-byte aligned");
    if (box0 == 0 || box0 > 256 || box1 == 0 || box1 > 256 || d0 == 0 || d1 == 0 || d2 == 0)
        return fail(ZVB_ERR_INVALID, "tensor map: bad box/dims");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ZVB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// ------------------------------------------------------------------------------------------ ops
enum OpType { OP_PMASKS, OP_MSMALL, OP_ROWBIAS, OP_CAST, OP_GEMM, OP_ATTN, OP_ATTN_TC, OP_BIASNORM, OP_PREP, OP_DOWN, OP_UP, OP_DWCONV, OP_MASK, OP_MASKW, OP_TSEMB, OP_SMALL,
              OP_LAYERNORM, OP_VOC_MASK, OP_VOC_WINDOW, OP_ISTFT_FRAMES, OP_OLA };

struct Op {
    OpType type;
    // GEMM / ATTN
    CUtensorMap ma, mb, mx, ms, mo;          // A, B, epilogue operand, output, bypass orig
    bool has_mx = false, has_ms = false, has_mo = false;
    GemmParams gp;
    int kind = 0, grid = 0, cluster = 1;
    AttnParams ap;
    Attn3Params ap3;
    // elementwise
    const void *p0 = nullptr, *p1 = nullptr;
    void *o0 = nullptr, *o1 = nullptr, *o2 = nullptr;
    const float *f0 = nullptr, *f1 = nullptr, *f2 = nullptr, *f3 = nullptr;
    long long rows = 0;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0, i4 = 0, i5 = 0;
    float w[4] = {0, 0, 0, 0};
    // profiling metadata: category (zvb_op_category) and algorithmic work (FLOPs for the tensor-core
    // kernels, bytes for the memory-bound ones)
    int cat = ZVB_CAT_OTHER;
    double work = 0.0;
    double bytes = 0.0;              // algorithmic HBM bytes: every operand read once, every result written once
    int shape[4] = {0, 0, 0, 0};     // GEMM: rows, out cols, K, block_n
    // fp16 outputs of the op (pointer, element count): scanned for saturated values when the plan has a
    // saturation counter (zvb_plan_set_saturation_counter)
    const void* scan_ptr[2] = {nullptr, nullptr};
    long long scan_n[2] = {0, 0};
};
static inline void mark_out(Op& op, int i, const void* p, long long n) { op.scan_ptr[i] = p; op.scan_n[i] = n; }

static inline long long pair_min_mtiles() { return g_pair_min_mtiles >= 0 ? g_pair_min_mtiles : 2LL * g_num_sms; }

// Small problems (single utterances: 20 m-tiles at T = 1219, 103 at T = 6563): a tile's time is the operand bytes its SM
// streams plus its epilogue, and the kernel's time is that times the number of WAVES the tiles need on 148 SMs / 74 CTA pairs.
// Round 1 picked multiples of 64 only: N = 512 at 20 m-tiles became 160 tiles of 64 columns = 2 waves where 140 tiles of 80
// columns run in one.  The constants are read off the per-launch critical path of a single-utterance sample (tools/timeline_c1.py,
// a -DZVB_TIMELINE build; gpurun_out/r2d_timeline.log): ~1.5 us until the first operands land + ~1.2 us of hand-offs and drain,
// 128 B x (128 + tile columns) per k-block at ~62 GB/s per SM, and an epilogue that costs 0.5-0.9 us per 64 columns on the lean
// paths (tile width a multiple of 64; the residual form also needs n_out % width == 0) but 2.5 us at <= 80 columns and 5-6 us at
// 144-224 columns on the generic path -- which an earlier version of this model ignored (it put every single-utterance GEMM on
// widths like 80 / 144 / 208 / 224 and hence on the generic epilogue).
// lean_kind: 0 = the op can only use the generic epilogue, 1 = lean plain epilogue possible, 2 = lean residual epilogue possible.
static int pick_block_n_small(int n_out, long long m_tiles, int k_blocks, int lean_kind) {
    int best = 0;
    double best_cost = 1e30;
    for (int bn = 16; bn <= 256; bn += 16) {
        const long long n_tiles = (n_out + bn - 1) / bn;
        if (bn > 16 && (n_tiles - 1) * bn >= n_out) continue;
        const long long slots2 = ((m_tiles + 1) / 2) * n_tiles;
        const bool pair = g_cluster_ok && bn >= 64 && m_tiles >= 2 && slots2 >= g_num_sms / 4 && k_blocks >= g_pair_min_kb &&
                          m_tiles >= pair_min_mtiles();
        const long long units = pair ? slots2 : m_tiles * n_tiles;
        const long long lanes = pair ? g_num_sms / 2 : g_num_sms;
        const long long waves = (units + lanes - 1) / lanes;
        const bool lean = g_small_lean && bn % 64 == 0 && (lean_kind == 1 || (lean_kind == 2 && n_out % bn == 0));
        double tile_us, cost;
        if (g_small_lean) {
            const double epi_us = lean ? 0.3 + 0.0065 * bn : 2.0 + 0.02 * bn;
            // a k-block of a single-CTA tile takes ~0.285 us whatever its width up to 128 columns (85-110 GB/s per SM: ring depth
            // over load latency), a pair's 0.35-0.38 us (62 GB/s per SM): gpurun_out/r2d_timeline6.log / r2d_timeline4.log
            const double kb_us = pair ? 128.0 * (128 + bn / 2) / 62000.0 : std::max(0.285, 128.0 * (128 + bn) / 110000.0);
            tile_us = 2.7 + k_blocks * kb_us + epi_us;
            cost = waves * tile_us;
        } else {                                       // the round-2 model before the timeline measurement (ZVB_SMALL_LEAN=0)
            tile_us = 5.2 + k_blocks * 128.0 * (128 + (pair ? bn / 2 : bn)) / 62000.0 + 0.004 * bn;
            cost = waves * tile_us;
            if (bn % 64 != 0) cost *= 1.03;
        }
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

static int pick_block_n(int n_out, long long m_tiles, int k_blocks, int lean_kind = 0) {
    if (g_small_model && n_out >= 64 && m_tiles * ((n_out + 255) / 256) < 2LL * g_num_sms)
        return pick_block_n_small(n_out, m_tiles, k_blocks, lean_kind);
    // short K (<= 8 k-blocks): the mainloop is bound by operand bytes in flight, and a 40 KB stage (192
    // columns) fits four times into the wide ring where a 48 KB stage (256 columns) fits three times
    if (g_bn192 && k_blocks <= 8 && n_out >= 768 && n_out % 192 == 0 && m_tiles * (n_out / 192) >= g_num_sms) return 192;
    // TMA stores move 64-column (h16) sub-tiles, so tile widths that are multiples of 64 are preferred
    // when they waste < 8% of the MMA work; otherwise the narrowest multiple of 16 that covers n_out.
    if (n_out >= 64) {
        int best = 0;
        double best_waste = 1e9;
        for (int bn = 256; bn >= 64; bn -= 64) {
            const int tiles = (n_out + bn - 1) / bn;
            if (m_tiles * tiles < g_num_sms && bn > 64) continue;       // small problems: more, narrower tiles
            const double waste = (double)tiles * bn / n_out - 1.0;
            // wider tiles move fewer operand bytes per output: a 256-column tile that wastes < 8% beats a
            // narrower exact fit (N = 1920: 8 x 256 at 962 TFLOP/s against 10 x 192 at 840)
            // (K <= 512: up to 12% padding, i.e. N = 1152 as 5 x 256 instead of 6 x 192 -- 192 columns are 24 units for 16
            // epilogue warps; with the lean epilogue 232 -> 203-210 us, ZVB_WIDE_WASTE=0.08 restores the exact fit)
            if (bn == 256 && waste < (k_blocks <= 8 ? g_wide_waste : 0.08) && g_wide_pref) return 256;
            if (waste < best_waste - 1e-9) { best_waste = waste; best = bn; }
        }
        if (best != 0 && best_waste < 0.08) return best;
    }
    int tiles = (n_out + 255) / 256;
    int bn = (((n_out + tiles - 1) / tiles) + 15) / 16 * 16;
    while (m_tiles * tiles < g_num_sms && bn > 64) {
        ++tiles;
        bn = (((n_out + tiles - 1) / tiles) + 15) / 16 * 16;
    }
    return bn;
}

// Output tensor maps for the TMA-store epilogue; eligible when whole 128-byte sub-tiles belong to one tile.
static int setup_tma_store(Op& op, int batches_rows /*rows per batch*/, int nbatch) {
    GemmParams& p = op.gp;
    const bool f32 = p.out_mode == OUT_F32;
    const bool bf = p.out_mode == OUT_H16;
    if (!g_tma_store_ok || !(f32 || bf)) return 0;
    if (op.kind == EPI_LINEAR) {
        if (p.out_col_stride != p.block_n) return 0;
        if (p.num_n_tiles > 1 && p.block_n % (bf ? 64 : 32) != 0) return 0;
    } else if (!bf) {
        return 0;
    }
    if ((p.ldc * (f32 ? 4 : 2)) % 16 != 0) return 0;
    const uint64_t esz = f32 ? 4 : 2;
    TRY(make_tmap(&op.ms, p.out, p.n_out, batches_rows, nbatch, (uint64_t)p.ldc * esz,
                  (uint64_t)p.ldc * esz * batches_rows, GEMM_BLOCK_M, f32));
    op.has_ms = true;
    p.tma_store = 1;
    return 0;
}

static void gp_defaults(GemmParams& p) { memset(&p, 0, sizeof p); p.rows_per_group = 1; }

struct LinearEpi {
    int act = ACT_NONE;
    const h16* resid = nullptr;          // fp16 residual stream tile (TMA aux ring), pitch = ldc
    const float* rowbias = nullptr;
    int rows_per_group = 1;
    const h16* orig = nullptr;           // bypass operand, fp16, pitch = ldc (needs resid and bypass_scale)
    const float* bypass_scale = nullptr;
    int out_mode = OUT_H16;
    int t_L = 0, t_pitch = 0, t_batch_rows = 0, t_hd = 1, t_hp = 1;   // OUT_T_H16
    int block_n = 0;                     // 0 = choose
    const uint8_t* row_mask = nullptr;   // rows with mask != 0 are written as zeros (generic epilogue)
    int act_cols = 0;                    // > 0: activation on output columns < act_cols only
};

// Decides whether the op runs as 2-CTA clusters with a multicast B tile, and the persistent grid.
static void set_grid(Op& op) {
    GemmParams& p = op.gp;
    const long long slots2 = (long long)p.batches * ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    // CTA pairs pay off when the mainloop dominates the tile (measured: K = 1920 GEMMs 320 -> 279 us,
    // P.V 222 -> 195 us).  For K = 512 they were a loss until the peer's accumulator hand-off stopped fencing
    // (relaxed remote arrive) and the epilogue arithmetic was packed; since then 8 k-blocks are worth it
    // (-30% L2->SM operand bytes; 790 -> 770 ms per sample, A/B on one box), fewer make no difference
    // ... and only on large problems: a pair pays a cluster launch, a cluster barrier in its set-up (+0.8 us, tools/timeline_c1.py)
    // and one at its exit, which a GEMM of a few tiles per SM never earns back -- without pairs a single utterance samples in
    // 47.8 ms instead of 51.5, two in 57.7 instead of 61.9, a 60 s dialog in 144-146 ms instead of 147; 16 utterances are even,
    // from 24 on (and the 16 x 2344-frame stereo batch) the pairs' smaller L2 -> SM traffic wins by 2 % (profiles/pair_threshold_r2.txt)
    op.cluster = (g_cluster_ok && p.block_n >= 64 && p.num_m_tiles >= 2 && slots2 >= g_num_sms / 4 &&
                  p.num_k_blocks >= g_pair_min_kb && (long long)p.batches * p.num_m_tiles >= pair_min_mtiles()) ? 2 : 1;
    if (op.cluster == 2) {
        const long long clusters = slots2 < g_num_sms / 2 ? slots2 : g_num_sms / 2;
        op.grid = static_cast<int>(clusters * 2);
    } else {
        const long long tiles = (long long)p.batches * p.num_m_tiles * p.num_n_tiles;
        op.grid = static_cast<int>(tiles < g_num_sms ? tiles : g_num_sms);
    }
    p.ring_bytes = GEMM_OPERAND_BYTES; p.aux_slots = GEMM_AUX_SLOTS; p.stage_depth = 2;
    gemm_ring(p.block_n, op.cluster, p.ring_bytes, &p.stages, &p.stage_bytes);
}

// Splits the 208 KB shared budget between the operand ring and the aux / staging slots (gemm.cuh); called once
// the epilogue operands and the store path of the op are known.
static void gemm_layout(Op& op) {
    GemmParams& p = op.gp;
    // A-stationary: K <= 512 GEMMs with several n-tiles per m-group are bound by the L2 -> SM operand bytes
    // (16 KB of A + the CTA's half of B per k-block and SM, ~64 B/clk against a ~43 B/clk chip-wide L2 cap); with
    // the pair's A tile resident for all n-tiles of its m-group only B streams (~35 B/clk).  Needs the aux-less
    // TMA-store epilogue (two staging buffers) so that 128 KB of A and >= 3 B stages fit.
    if (g_resident_ok && op.cluster == 2 && p.aux_mode == AUX_NONE && p.tma_store && p.num_n_tiles >= 2 &&
        p.num_k_blocks <= 8 && p.a_zn == 0) {
        const int a_bytes = p.num_k_blocks * GEMM_A_BYTES;
        const int b_stage = (p.block_n / 2) * GEMM_BLOCK_K * 2;
        int stages = (GEMM_SHARED_BUDGET - 2 * GEMM_AUX_BYTES - a_bytes) / b_stage;
        if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
        if (stages >= 3) {
            p.a_resident = 1; p.a_bytes = a_bytes; p.stages = stages; p.stage_bytes = b_stage;
            p.ring_bytes = a_bytes + stages * b_stage; p.aux_slots = 2; p.stage_depth = 1;
            return;
        }
    }
    if (!g_layout_ok) return;
    int sb, st;
    gemm_ring(p.block_n, op.cluster, GEMM_OPERAND_BYTES, &st, &sb);
    if (p.aux_mode == AUX_NONE && p.tma_store) {
        // wide ring: one staging buffer per epilogue half, everything else to the operands
        const int ring = GEMM_SHARED_BUDGET - 2 * GEMM_AUX_BYTES;
        int st2, sb2;
        gemm_ring(p.block_n, op.cluster, ring, &st2, &sb2);
        if (st2 > st && p.num_k_blocks > st) {
            p.ring_bytes = ring; p.aux_slots = 2; p.stage_depth = 1; p.stages = st2;
        }
    } else if (p.aux_mode != AUX_NONE && p.tma_store && p.num_k_blocks * 2 < st) {
        // deep aux: the operands of two tiles are all the ring ever holds
        const int need = 2 * p.num_k_blocks;
        int slots = (GEMM_SHARED_BUDGET - need * sb) / GEMM_AUX_BYTES;
        if (slots > GEMM_AUX_SLOTS_MAX) slots = GEMM_AUX_SLOTS_MAX;
        if (p.orig_tma) slots &= ~1;          // operand + `orig` entries travel in pairs
        p.aux_slots = slots; p.stages = need; p.ring_bytes = GEMM_SHARED_BUDGET - slots * GEMM_AUX_BYTES;
    }
}
static inline uint32_t b_box_rows(const Op& op) { return op.gp.block_n / op.cluster; }

// out[M, n_out] = epi(A[M, K] · W[n_out, K]ᵀ + b)
static int build_linear(Op& op, const h16* A, long long M, int lda, const zvb_linear& lin, void* out, int ldc,
                        const LinearEpi& e) {
    op.type = OP_GEMM;
    op.kind = EPI_LINEAR;
    if (lin.k_pitch % 8 != 0 || lda % 8 != 0) return fail(ZVB_ERR_INVALID, "linear: pitches must be multiples of 8");
    const int K = lin.k_pitch < lda ? lin.k_pitch : lda;    // both zero padded beyond in_features
    const long long m_tiles = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    // which lean epilogue the op could take if its tile width allows (the conditions of fast_epi / fast_resid below)
    int lean_kind = 0;
    if (e.out_mode == OUT_H16 && e.row_mask == nullptr && lin.out_features % 8 == 0 && ldc % 8 == 0 &&
        (e.rowbias == nullptr || e.rows_per_group >= GEMM_BLOCK_M)) {
        if (e.resid == nullptr && e.act_cols % 32 == 0 && (reinterpret_cast<uintptr_t>(lin.b) & 15) == 0) lean_kind = g_fast_epi ? 1 : 0;
        else if (e.resid != nullptr && e.act == ACT_NONE && g_fast_resid &&
                 (e.orig == nullptr || (g_fast_bypass && e.rowbias == nullptr)))
            lean_kind = 2;
    }
    int bn = e.block_n ? e.block_n
                       : pick_block_n(lin.out_features, m_tiles, (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K, lean_kind);
    // plain fp16 projections whose width is no multiple of 64 (attention in_proj: 272 columns): tiles of a multiple of 64
    // columns keep them on the lean epilogue (whole 64-column store boxes per tile; columns past n_out are clipped by the
    // store's tensor map and their weight rows are TMA zero fill) -- 2 x 192 instead of 2 x 144
    if (e.block_n == 0 && g_fast_epi && g_lean_pad && bn % 64 != 0 && lin.out_features > 128 && e.out_mode == OUT_H16 &&
        m_tiles * ((lin.out_features + 255) / 256) >= 2LL * g_num_sms &&
        e.resid == nullptr && e.rowbias == nullptr && e.row_mask == nullptr && lin.out_features % 8 == 0) {
        const int tiles = (lin.out_features + 255) / 256;
        bn = (((lin.out_features + tiles - 1) / tiles) + 63) / 64 * 64;
    }
    GemmParams& p = op.gp;
    gp_defaults(p);
    p.M = static_cast<int>(M);
    p.n_out = lin.out_features;
    p.num_k_blocks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    p.block_n = bn;
    p.num_m_tiles = static_cast<int>(m_tiles);
    p.num_n_tiles = (lin.out_features + bn - 1) / bn;
    p.batches = 1;
    p.out_mode = e.out_mode; p.out = out; p.ldc = ldc;
    p.out_col_stride = bn; p.n_valid = bn;
    p.bias = lin.b;
    p.rowbias = e.rowbias; p.rows_per_group = e.rows_per_group; p.ld_rowbias = lin.out_features;
    p.bypass_scale = e.bypass_scale;
    p.act = e.act;
    p.act_cols = e.act_cols;
    p.row_mask = e.row_mask;
    p.t_L = e.t_L; p.t_pitch = e.t_pitch; p.t_batch_rows = e.t_batch_rows; p.t_hd = e.t_hd; p.t_hp = e.t_hp;
    set_grid(op);
    TRY(make_tmap(&op.ma, A, K, M, 1, (uint64_t)lda * 2, (uint64_t)lda * 2 * M, GEMM_BLOCK_M));
    TRY(make_tmap(&op.mb, lin.w, K, lin.rows, 1, (uint64_t)lin.k_pitch * 2, (uint64_t)lin.k_pitch * 2 * lin.rows,
                  b_box_rows(op)));
    if (e.resid != nullptr) {
        if (e.out_mode != OUT_H16 || ldc % 8 != 0)
            return fail(ZVB_ERR_INVALID, "linear: a residual needs an fp16 output with a pitch that is a multiple of 8");
        p.aux_mode = AUX_ADD_H16; p.aux_zb = 0;
        TRY(make_tmap(&op.mx, e.resid, ldc, M, 1, (uint64_t)ldc * 2, (uint64_t)ldc * 2 * M, GEMM_BLOCK_M));
        op.has_mx = true;
    }
    TRY(setup_tma_store(op, (int)M, 1));
    // bypass: `orig` rides through the aux ring next to the residual sub-tiles
    if (e.orig != nullptr) {
        if (p.aux_mode != AUX_ADD_H16 || e.bypass_scale == nullptr || lin.out_features % 64 != 0 ||
            (reinterpret_cast<uintptr_t>(e.bypass_scale) & 15) != 0)
            return fail(ZVB_ERR_INVALID, "linear: bypass needs a residual, a 16-byte aligned scale and out_features %% 64 == 0");
        TRY(make_tmap(&op.mo, e.orig, ldc, M, 1, (uint64_t)ldc * 2, (uint64_t)ldc * 2 * M, GEMM_BLOCK_M));
        op.has_mo = true;
        p.orig_tma = 1;
    }
    gemm_layout(op);
    p.pre_b = (g_pre_b && g_plan_build) ? 1 : 0;        // B = model weights: requested before the dependency wait
    p.fast_epi = (g_fast_epi && p.tma_store && p.aux_mode == AUX_NONE && p.out_mode == OUT_H16 &&
                  (p.rowbias == nullptr || p.rows_per_group >= GEMM_BLOCK_M) && p.act_cols % 32 == 0 &&
                  p.rowscale == nullptr && p.row_mask == nullptr && bn % 64 == 0 && lin.out_features % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(lin.b) & 15) == 0) ? 1 : 0;
    {
        const bool base = g_fast_resid && e.act == ACT_NONE && p.tma_store && p.aux_mode == AUX_ADD_H16 && p.out_mode == OUT_H16 &&
                          p.rowscale == nullptr && p.row_mask == nullptr && bn % 64 == 0 && lin.out_features % bn == 0;
        p.fast_resid = 0;
        if (base && !p.orig_tma && (p.rowbias == nullptr || p.rows_per_group >= GEMM_BLOCK_M)) p.fast_resid = 1;
        else if (base && p.orig_tma && p.rowbias == nullptr && g_fast_bypass) p.fast_resid = 2;      // + bypass (LEAN 3)
    }
    if (e.out_mode == OUT_H16) mark_out(op, 0, out, M * ldc);
    else if (e.out_mode == OUT_T_H16 && e.t_L > 0) mark_out(op, 0, out, (M / e.t_L) * (long long)e.t_batch_rows * e.t_pitch);
    op.shape[0] = (int)M; op.shape[1] = lin.out_features; op.shape[2] = lin.in_features; op.shape[3] = bn;
    op.cat = ZVB_CAT_GEMM_LINEAR;
    op.work = 2.0 * (double)M * lin.out_features * lin.in_features;
    op.bytes = 2.0 * ((double)M * lin.in_features + (double)lin.out_features * lin.in_features) +
               (double)M * lin.out_features * ((e.out_mode == OUT_F32 ? 4.0 : 2.0) + (e.resid ? 2.0 : 0.0) + (e.orig ? 2.0 : 0.0));
    return 0;
}

// gated projection (weights packed per 256-row tile as [128 | 128]); n_out = gated outputs
static int build_gated(Op& op, const h16* A, long long M, int lda, const zvb_linear& lin, int n_out, int gate_mode,
                       void* out, int ldc, const uint8_t* row_mask, const LinearEpi& e) {
    op.type = OP_GEMM;
    op.kind = EPI_GATED;
    const int K = lin.k_pitch < lda ? lin.k_pitch : lda;
    GemmParams& p = op.gp;
    gp_defaults(p);
    p.M = static_cast<int>(M);
    p.n_out = n_out;
    p.num_k_blocks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    p.block_n = 256;
    p.num_m_tiles = static_cast<int>((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
    p.num_n_tiles = (n_out + 127) / 128;
    if (lin.rows != p.num_n_tiles * 256) return fail(ZVB_ERR_INVALID, "gated linear: expected %d packed rows, got %d", p.num_n_tiles * 256, lin.rows);
    p.batches = 1;
    p.out_mode = e.out_mode; p.out = out; p.ldc = ldc;
    p.out_col_stride = 128; p.n_valid = 256;
    p.bias = lin.b;
    p.gate_mode = gate_mode;
    p.row_mask = row_mask;
    p.t_L = e.t_L; p.t_pitch = e.t_pitch; p.t_batch_rows = e.t_batch_rows; p.t_hd = e.t_hd; p.t_hp = e.t_hp;
    set_grid(op);
    TRY(make_tmap(&op.ma, A, K, M, 1, (uint64_t)lda * 2, (uint64_t)lda * 2 * M, GEMM_BLOCK_M));
    TRY(make_tmap(&op.mb, lin.w, K, lin.rows, 1, (uint64_t)lin.k_pitch * 2, (uint64_t)lin.k_pitch * 2 * lin.rows,
                  b_box_rows(op)));
    TRY(setup_tma_store(op, (int)M, 1));
    gemm_layout(op);
    p.pre_b = (g_pre_b && g_plan_build) ? 1 : 0;        // B = model weights: requested before the dependency wait
    if (e.out_mode == OUT_H16) mark_out(op, 0, out, M * ldc);
    else if (e.out_mode == OUT_T_H16 && e.t_L > 0) mark_out(op, 0, out, (M / e.t_L) * (long long)e.t_batch_rows * e.t_pitch);
    op.shape[0] = (int)M; op.shape[1] = 2 * n_out; op.shape[2] = lin.in_features; op.shape[3] = 256;
    op.cat = ZVB_CAT_GEMM_GATED;
    op.work = 2.0 * (double)M * (2.0 * n_out) * lin.in_features;
    op.bytes = 2.0 * ((double)M * lin.in_features + 2.0 * n_out * lin.in_features + (double)M * n_out);
    return 0;
}

// out[n*L+i, :] = P[n,h] · V  with V given transposed: Vt[n][rows][Lk].
//   per_head != 0 (SelfAttention): head h uses Vt rows [h*hp, h*hp+hd) -> out cols [h*hd, (h+1)*hd)
//   per_head == 0 (NonlinAttention): head 0 weights, all `hd` value columns, out *= mul (h16 [N*L, ldm])
static int build_pv(Op& op, const h16* P, const float* inv_l, const h16* Vt, void* out, int ldc, int N, int H, int L,
                    int Lk, int hd, int hp, int per_head, const h16* mul, int ldm) {
    op.type = OP_GEMM;
    op.kind = EPI_LINEAR;
    GemmParams& p = op.gp;
    gp_defaults(p);
    p.M = L;
    p.num_k_blocks = (Lk + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    p.num_m_tiles = (L + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    p.batches = N;
    p.b_zb = 1;
    p.a_zb = H;
    p.out_mode = OUT_H16; p.out = out; p.ldc = ldc;
    p.rowscale = inv_l; p.rs_zb = H; p.rs_zn = per_head ? 1 : 0;
    int vt_rows;
    if (per_head) {
        if (hp % 16 != 0 || hd > hp) return fail(ZVB_ERR_INVALID, "pv: head pad must be a multiple of 16");
        p.block_n = hp; p.num_n_tiles = H; p.a_zn = 1;
        p.n_out = H * hd; p.out_col_stride = hd; p.n_valid = hd;
        vt_rows = H * hp;
    } else {
        const int tiles = (hd + 255) / 256;
        p.block_n = (((hd + tiles - 1) / tiles) + 15) / 16 * 16;
        // single utterances: 2 x 10 m-tiles x 2 n-tiles = 40 CTAs streamed 20 k-blocks of 40 KB each (9.2 us of the launch's
        // 14 us, tools/timeline_c1.py); narrower tiles put the same bytes on three times as many SMs
        if (g_small_model && g_small_lean && (long long)N * p.num_m_tiles * tiles < g_num_sms)
            p.block_n = pick_block_n_small(hd, (long long)N * p.num_m_tiles, p.num_k_blocks, 0);
        if (g_pv_bn > 0 && g_pv_bn % 16 == 0 && g_pv_bn <= 256) p.block_n = g_pv_bn;      // ZVB_PV_BN: measurement override
        p.num_n_tiles = (hd + p.block_n - 1) / p.block_n; p.a_zn = 0;
        p.n_out = hd; p.out_col_stride = p.block_n; p.n_valid = p.block_n;
        vt_rows = hd;
        if (mul != nullptr) {
            if (ldm % 8 != 0) return fail(ZVB_ERR_INVALID, "pv: multiplier pitch must be a multiple of 8");
            p.aux_mode = AUX_MUL_H16; p.aux_zb = 1;
            TRY(make_tmap(&op.mx, mul, ldm, L, N, (uint64_t)ldm * 2, (uint64_t)ldm * 2 * L, GEMM_BLOCK_M));
            op.has_mx = true;
        }
    }
    set_grid(op);
    TRY(make_tmap(&op.ma, P, Lk, L, (uint64_t)N * H, (uint64_t)Lk * 2, (uint64_t)Lk * 2 * L, GEMM_BLOCK_M));
    TRY(make_tmap(&op.mb, Vt, Lk, vt_rows, N, (uint64_t)Lk * 2, (uint64_t)Lk * 2 * vt_rows, b_box_rows(op)));
    if (!per_head) TRY(setup_tma_store(op, L, N));
    gemm_layout(op);
    // inside a plan the weights P were written by the attention kernel at least two launches earlier (a launch starts
    // only after its predecessor passed its own dependency wait), so their tiles may be requested before the wait; V^T is
    // the predecessor's output and is not
    p.pre_b = (g_pre_b && g_plan_build) ? 2 : 0;
    // SelfAttention: the kernel only streams P (one 18 KB stage per k-block, ring-depth bound) and its 12-column head outputs
    // always leave through the direct per-thread stores (never a multiple of 8 columns: store_unit's staged path is not taken),
    // so the staging area is unused and the ring may take all of it but one slot: 10 stages instead of 8
    if (g_pv_deep && per_head && (hd & 7) != 0 && p.aux_mode == AUX_NONE && !p.tma_store && op.cluster == 1) {
        const int ring = GEMM_SHARED_BUDGET - GEMM_AUX_BYTES;
        int st, sb;
        gemm_ring(p.block_n, op.cluster, ring, &st, &sb);
        if (st > p.stages) { p.ring_bytes = ring; p.aux_slots = 1; p.stage_depth = 1; p.stages = st; p.stage_bytes = sb; }
    }
    mark_out(op, 0, out, (long long)N * L * ldc);
    op.shape[0] = N * L; op.shape[1] = per_head ? H * hd : hd; op.shape[2] = L; op.shape[3] = p.block_n;
    op.cat = ZVB_CAT_GEMM_PV;
    op.work = 2.0 * (double)N * (per_head ? H : 1) * (double)L * L * hd;
    op.bytes = 2.0 * ((double)N * (per_head ? H : 1) * L * Lk + (double)N * vt_rows * Lk +
                      (double)N * L * (per_head ? H * hd : hd) * (mul != nullptr ? 2.0 : 1.0));
    return 0;
}

static inline int attn_mask_words(int L) { return 4 * ((L + ATT_BN - 1) / ATT_BN); }

// pos_table: the layout of weights.py:pack_pos_table for this L; maskw: [N][attn_mask_words(L)] (mask_words_op)
static int build_attn(Op& op, const h16* qkp, int ld, const void* pos_table, const uint32_t* maskw, h16* P,
                      float* inv_l, int N, int H, int L, int Lk) {
    op.type = OP_ATTN;
    AttnParams& a = op.ap;
    a.L = L; a.Lk = Lk; a.H = H; a.N = N; a.qd = H * 32;
    a.qkp = qkp; a.ld = ld; a.P = P; a.inv_l = inv_l;
    if ((reinterpret_cast<uintptr_t>(pos_table) & 15) != 0) return fail(ZVB_ERR_INVALID, "attn: pos table must be 16-byte aligned");
    a.Epair = reinterpret_cast<const uint4*>(pos_table);
    a.emax = reinterpret_cast<const float*>(a.Epair + static_cast<size_t>(H) * (2 * L - 1 + 2 * ATT_POS_PAD));
    a.maskw = maskw; a.mask_words = attn_mask_words(L);
    if (ld % 8 != 0 || Lk % 8 != 0) return fail(ZVB_ERR_INVALID, "attn: pitches must be multiples of 8");
    TRY(make_tmap(&op.ma, qkp, ld, L, N, (uint64_t)ld * 2, (uint64_t)ld * 2 * L, ATT_BM));
    // P as (Lk, L, N*H): every softmax warp stores 32-row x 64-column boxes
    TRY(make_tmap(&op.ms, P, Lk, L, (uint64_t)N * H, (uint64_t)Lk * 2, (uint64_t)Lk * 2 * L, 32));
    op.has_ms = true;
    mark_out(op, 0, P, (long long)N * H * L * Lk);
    op.cat = ZVB_CAT_ATTN_WEIGHTS;
    // q.k (K = 32) + rel-pos (4-dim dot against 2L-1 offsets), reference FLOP model SURVEY.md §8d
    op.work = (double)N * H * (2.0 * L * L * 32 + 2.0 * L * (2.0 * L - 1) * 4);
    op.bytes = 2.0 * ((double)N * L * (H * 68) + (double)N * H * L * Lk) + 4.0 * N * H * L;     // q, k, p read once (not the row pitch)
    return 0;
}

// entries per copy of the tensor-core rel-pos table (weights.py: pack_pos_table_tc uses the same formula)
static inline int attn_tc_lz(int L) { return (2 * L + 264 + 1) / 2 * 2; }

static int build_attn_tc(Op& op, const h16* qkp, int ld, const void* pos_table_tc, const uint32_t* maskw, h16* P,
                         float* inv_l, int N, int H, int L, int Lk) {
    op.type = OP_ATTN_TC;
    Attn3Params& a = op.ap3;
    a.L = L; a.Lk = Lk; a.H = H; a.N = N; a.qd = H * 32;
    a.qkp = qkp; a.ld = ld; a.P = P; a.inv_l = inv_l;
    if (pos_table_tc == nullptr || (reinterpret_cast<uintptr_t>(pos_table_tc) & 15) != 0)
        return fail(ZVB_ERR_INVALID, "attn: tensor-core pos table missing or not 16-byte aligned");
    a.Z = reinterpret_cast<const uint2*>(pos_table_tc);
    a.LZ = attn_tc_lz(L);
    a.emax = reinterpret_cast<const float*>(a.Z + static_cast<size_t>(H) * 2 * a.LZ);
    a.maskw = maskw; a.mask_words = attn_mask_words(L);
    a.dbg = getenv("ZVB_ATTN_DBG") ? atoi(getenv("ZVB_ATTN_DBG")) : 0;
    if (ld % 8 != 0 || Lk % 8 != 0) return fail(ZVB_ERR_INVALID, "attn: pitches must be multiples of 8");
    TRY(make_tmap(&op.ma, qkp, ld, L, N, (uint64_t)ld * 2, (uint64_t)ld * 2 * L, A3_BM));
    TRY(make_tmap_sw64(&op.ms, P, Lk, L, (uint64_t)N * H, (uint64_t)Lk * 2, (uint64_t)Lk * 2 * L, 32));
    op.has_ms = true;
    mark_out(op, 0, P, (long long)N * H * L * Lk);
    op.cat = ZVB_CAT_ATTN_WEIGHTS;
    op.work = (double)N * H * (2.0 * L * L * 32 + 2.0 * L * (2.0 * L - 1) * 4);
    op.bytes = 2.0 * ((double)N * L * (H * 68) + (double)N * H * L * Lk) + 4.0 * N * H * L;     // q, k, p read once (not the row pitch)
    return 0;
}

static Op mask_words_op(const uint8_t* mask, uint32_t* out, int N, int L) {
    Op op; op.type = OP_MASKW; op.p0 = mask; op.o0 = out; op.i0 = N; op.i1 = L; op.i2 = attn_mask_words(L);
    return op;
}

static int dw_tile_frames(int K, int mode) {
    if (K > 15 && mode == 2) return DwShape<31, 2>::TT;
    return DW_TT;
}

static int build_dwconv(Op& d, const h16* x, h16* out, const float* w, const float* b, int N, int L, int C, int K, int act = 1) {
    d = Op();
    d.type = OP_DWCONV; d.p0 = x; d.o0 = out; d.f0 = w; d.f1 = b;
    d.i0 = N; d.i1 = L; d.i2 = C; d.i3 = K; d.i4 = act;
    if (act == 0 && K != 7) return fail(ZVB_ERR_INVALID, "dwconv without activation is built for 7 taps only");
    if (K != 7 && K != 9 && K != 15 && K != 31)
        return fail(ZVB_ERR_INVALID, "depthwise kernel size %d not built (7, 9, 15, 31)", K);
    if (C % 8 != 0) return fail(ZVB_ERR_INVALID, "dwconv: channels must be a multiple of 8");
    // x as (C, L, N); a box = 64 channels x (128 + K - 1) frames, rows outside [0, L) are zero-filled
    d.i5 = g_dw_mode;                  // the tile height is part of the tensor map: the shape is fixed at plan creation
    TRY(make_tmap_plain(&d.ma, x, C, L, N, (uint64_t)C * 2, (uint64_t)C * 2 * L, 64, dw_tile_frames(K, g_dw_mode) + K - 1));
    mark_out(d, 0, out, (long long)N * L * C);
    d.cat = ZVB_CAT_DWCONV;
    d.shape[0] = N * L; d.shape[1] = C; d.shape[2] = K; d.shape[3] = act;
    d.work = 2.0 * 2.0 * (double)N * L * C;
    d.bytes = d.work;
    return 0;
}

template <int K, int ACT, int MODE>
static void launch_dwconv_mode(const Op& op, cudaStream_t st) {
    const int N = op.i0, L = op.i1, C = op.i2;
    const int groups = (C + 63) / 64;
    const int tiles = N * ((L + DwShape<K, MODE>::TT - 1) / DwShape<K, MODE>::TT);
    int per_group = (DwShape<K, MODE>::MINB * g_num_sms) / groups;         // resident blocks per SM
    if (per_group < 1) per_group = 1;
    if (per_group > tiles) per_group = tiles;
    launch_k(dwconv_kernel<K, ACT, MODE>, dim3(per_group, groups), dim3(DwShape<K, MODE>::THREADS), dw_smem_bytes<K, MODE>(), st,
             op.ma, (h16*)op.o0, op.f0, op.f1, L, C, N);
}
template <int K, int ACT = 1>
static void launch_dwconv(const Op& op, cudaStream_t st) {
    if (op.i5 == 0) launch_dwconv_mode<K, ACT, 0>(op, st);
    else if (op.i5 == 2) launch_dwconv_mode<K, ACT, 2>(op, st);
    else launch_dwconv_mode<K, ACT, 1>(op, st);
}

// Cluster size of the attention-weights kernel's key split (attn.cuh) for `base` = query tiles x heads x utterances CTAs
// with `q_tiles` key tiles each.
static int attn_split_choice(long long base, int q_tiles) {
    int cs = 1;
    if (!g_attn_split) return cs;
    if (q_tiles >= 4 && base * 4 <= g_num_sms + g_num_sms / 8) cs = 4;
    else if (q_tiles >= 2 && base * 2 <= g_num_sms + g_num_sms / 8) cs = 2;
    else if (q_tiles >= 8) {
        // tail of the last wave: two CTAs per SM = 296 slots; a 60 s dialog has 416 CTAs of 104 tile iterations each
        // (two waves for 1.4 waves of work), as pairs 832 half-length CTAs in three (measured: 2.41 -> 2.16 ms per
        // forward; on grids of many waves the split only adds its hand-offs: C5 +16%, C3 +31% when forced)
        const long long slots = 2LL * g_num_sms;
        const long long w1 = (base + slots - 1) / slots, w2 = (2 * base + slots - 1) / slots;
        if (10 * w2 <= 16 * w1) cs = 2;           // <= 0.8 of the unsplit wave count
    }
    if (g_attn_split > 1 && q_tiles >= g_attn_split) cs = g_attn_split == 3 ? 2 : g_attn_split;   // tests: force 2 (=3) or 4
    return cs;
}

static int launch_op(const Op& op, cudaStream_t st) {
    switch (op.type) {
        case OP_GEMM: {
            if (op.grid <= 0) return 0;
            const CUtensorMap& mx = op.has_mx ? op.mx : op.ma;
            const CUtensorMap& ms = op.has_ms ? op.ms : op.ma;
            const CUtensorMap& mo = op.has_mo ? op.mo : op.ma;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(op.grid);
            cfg.blockDim = dim3(GEMM_THREADS);
            cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
            cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = op.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
            const int lean = op.kind == EPI_GATED ? 0 : (op.gp.fast_epi ? 1 : op.gp.fast_resid == 2 ? 3 : op.gp.fast_resid ? 2 : 0);
            gemm_fn_t fn = gemm_fn(op.kind, op.gp.act, op.cluster, lean);
            if (fn == nullptr) return fail(ZVB_ERR_INVALID, "gemm: no kernel for kind %d act %d lean %d", op.kind, op.gp.act, lean);
            cudaError_t e = cudaLaunchKernelEx(&cfg, fn, op.ma, op.mb, mx, ms, mo, op.gp);
            if (e != cudaSuccess) return fail(ZVB_ERR_CUDA, "launch gemm: %s", cudaGetErrorString(e));
            return check_launch("gemm");
        }
        case OP_ATTN: {
            const int q_tiles = (op.ap.L + ATT_BM - 1) / ATT_BM;          // also the number of key tiles
            // small grids (single utterances): the key tiles of a query tile are split over a cluster of 2 or 4 CTAs while the
            // grid still fits the SMs about once (attn.cuh; ZVB_ATTN_SPLIT=0 keeps one CTA per query tile)
            const int cs = attn_split_choice((long long)q_tiles * op.ap.H * op.ap.N, q_tiles);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(q_tiles * cs, op.ap.H, op.ap.N);
            cfg.blockDim = dim3(ATT_THREADS);
            cfg.dynamicSmemBytes = ATT_SMEM_BYTES;
            cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, cs == 4 ? attn_weights_kernel<4> : cs == 2 ? attn_weights_kernel<2> : attn_weights_kernel<1>,
                                               op.ma, op.ms, op.ap);
            if (e != cudaSuccess) return fail(ZVB_ERR_CUDA, "launch attn_weights: %s", cudaGetErrorString(e));
            return check_launch("attn_weights");
        }
        case OP_ATTN_TC: {
            dim3 grid((op.ap3.L + A3_BM - 1) / A3_BM, op.ap3.H, op.ap3.N);
            launch_k(attn_weights_tc_kernel, dim3(grid), dim3(A3_THREADS), A3_SMEM_BYTES, st, op.ma, op.ms, op.ap3);
            return check_launch("attn_weights_tc");
        }
        case OP_BIASNORM: {
            const int blocks = static_cast<int>((op.rows + 7) / 8);
            if (op.i0 <= 512)
                launch_k(biasnorm_bypass_kernel<2>, dim3(blocks), dim3(256), 0, st, 
                    (const h16*)op.p0, (const h16*)op.p1, (h16*)op.o0, (h16*)op.o1, op.f3, op.i1,
                    op.f0, op.f1, op.f2, op.rows, op.i0);
            else
                launch_k(biasnorm_bypass_kernel<4>, dim3(blocks), dim3(256), 0, st, 
                    (const h16*)op.p0, (const h16*)op.p1, (h16*)op.o0, (h16*)op.o1, op.f3, op.i1,
                    op.f0, op.f1, op.f2, op.rows, op.i0);
            return check_launch("biasnorm_bypass");
        }
        case OP_PREP: {
            const long long n = ((op.rows + 3) / 4) * (op.i0 / 8);
            launch_k(stream_prep_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const h16*)op.p0, (h16*)op.o0, op.f0, op.i1,
                                                                           op.rows, op.i0);
            return check_launch("stream_prep");
        }
        case OP_DOWN: {
            const long long n = (long long)op.i0 * ((op.i2 + 1) / 2) * (op.i4 / 8);
            launch_k(downsample_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 
                (const h16*)op.p0, (h16*)op.o0, op.i0, op.i1, op.i2, op.i3, op.w[0], op.w[1], op.w[2], op.w[3], op.i4);
            return check_launch("downsample");
        }
        case OP_UP: {
            const long long n = (long long)op.i0 * op.i2 * (op.i4 / 8);
            launch_k(upsample_combine_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 
                (const h16*)op.p0, (const h16*)op.p1, (h16*)op.o0, op.f0, op.i0, op.i1, op.i2, op.i3, op.i4);
            return check_launch("upsample_combine");
        }
        case OP_DWCONV: {
            const int K = op.i3;
            if (K == 7 && op.i4 == 0) launch_dwconv<7, 0>(op, st);
            else if (K == 7) launch_dwconv<7>(op, st);
            else if (K == 9) launch_dwconv<9>(op, st);
            else if (K == 15) launch_dwconv<15>(op, st);
            else if (K == 31) launch_dwconv<31>(op, st);
            else return fail(ZVB_ERR_INVALID, "depthwise kernel size %d not built (7, 9, 15, 31)", K);
            return check_launch("dwconv");
        }
        case OP_MASK: {
            const int n = op.i0 * op.i2;
            launch_k(stride_mask_kernel, dim3((n + 255) / 256), dim3(256), 0, st, (const uint8_t*)op.p0, (uint8_t*)op.o0, op.i0, op.i1,
                                                                 op.i2, op.i3);
            return check_launch("stride_mask");
        }
        case OP_MASKW: {
            const int n = op.i0 * op.i2;
            launch_k(mask_words_kernel, dim3((n + 255) / 256), dim3(256), 0, st, (const uint8_t*)op.p0, (uint32_t*)op.o0, op.i0, op.i1, op.i2);
            return check_launch("mask_words");
        }
        case OP_TSEMB: {
            const int n = op.i0 * (op.i1 / 2);
            launch_k(timestep_embedding_kernel, dim3((n + 127) / 128), dim3(128), 0, st, op.f0, (float*)op.o0, op.i0, op.i1);
            return check_launch("timestep_embedding");
        }
        case OP_SMALL: {
            const long long warps = (long long)op.i0 * op.i2;
            launch_k(small_linear_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, st, 
                op.f0, op.f1, op.f2, op.f3, (float*)op.o0, op.i0, op.i1, op.i2, op.i3, op.i4);
            return check_launch("small_linear");
        }
        case OP_PMASKS: {
            const MaskJobs& mj = *static_cast<const MaskJobs*>(op.p1);
            int maxw = 1;
            for (int i = 0; i < mj.n; ++i) maxw = mj.j[i].nwords > maxw ? mj.j[i].nwords : maxw;
            launch_k(prologue_masks_kernel, dim3((op.i0 * maxw + 255) / 256, mj.n), dim3(256), 0, st, (const uint8_t*)op.p0, mj, op.i0, op.i1);
            return check_launch("prologue_masks");
        }
        case OP_MSMALL: {
            const SmallJobs& sj = *static_cast<const SmallJobs*>(op.p1);
            const long long warps = (long long)op.i0 * op.i2;
            launch_k(multi_small_linear_kernel, dim3((unsigned)((warps * 32 + 255) / 256), sj.n), dim3(256), 0, st, op.f0, sj, op.i0, op.i1, op.i2);
            return check_launch("multi_small_linear");
        }
        case OP_ROWBIAS: {
            const RowBiasJobs& rj = *static_cast<const RowBiasJobs*>(op.p1);
            if (op.i0 <= RB_NCHUNK)
                launch_k(rowbias_small_kernel, dim3((op.i2 + 7) / 8, rj.n), dim3(256), 0, st, rj, op.i0, op.i1, op.i3, op.i2, op.i4);
            else
                launch_k(rowbias_all_kernel, dim3((op.i2 + RB_COLS - 1) / RB_COLS, rj.n, (op.i0 + RB_NCHUNK - 1) / RB_NCHUNK), dim3(RB_COLS), 0, st,
                         rj, op.i0, op.i1, op.i3, op.i2, op.i4);
            return check_launch("rowbias_all");
        }
        case OP_CAST: {
            const long long n = op.rows * op.i1;
            launch_k(cast_pad_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const float*)op.p0, (h16*)op.o0, op.rows,
                     op.i0, op.i1);
            return check_launch("cast_pad");
        }
        case OP_LAYERNORM: {
            const int blocks = static_cast<int>((op.rows + 7) / 8);
            launch_k(layernorm_kernel<4>, dim3(blocks), dim3(256), 0, st, (const h16*)op.p0, (h16*)op.o0, op.f0, op.f1,
                     (const uint8_t*)op.p1, op.rows, op.i0, op.w[0]);
            return check_launch("layernorm");
        }
        case OP_VOC_MASK: {
            const int n = op.i0 * op.i1;
            launch_k(voc_mask_kernel, dim3((n + 255) / 256), dim3(256), 0, st, (const int*)op.p0, (uint8_t*)op.o0, op.i0, op.i1);
            return check_launch("voc_mask");
        }
        case OP_VOC_WINDOW: {
            const long long n = (long long)op.i0 * op.i1 * op.i4;
            launch_k(voc_window_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const float*)op.p0, (const int*)op.p1,
                     (h16*)op.o0, op.i0, op.i1, op.i2, op.i3, op.i4, op.w[0]);
            return check_launch("voc_window");
        }
        case OP_ISTFT_FRAMES: {
            long long blocks = op.rows < 8LL * g_num_sms ? op.rows : 8LL * g_num_sms;
            if (blocks < 1) blocks = 1;
            launch_k(voc_istft_frames_kernel, dim3((unsigned)blocks), dim3(AUD_THREADS), (size_t)AUD_FFT_SMEM, st, (const float*)op.p0,
                     op.i0, (const uint8_t*)op.p1, op.f0, (float*)op.o0, op.rows);
            return check_launch("voc_istft_frames");
        }
        case OP_OLA: {
            const long long n = (long long)op.i0 * op.i2 * (op.i1 - 1);
            if (n <= 0) return 0;
            launch_k(voc_overlap_add_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const float*)op.p0,
                     (const int*)op.p1, op.f0, (float*)op.o0, op.i0, op.i1, op.i2, op.i5);
            return check_launch("voc_overlap_add");
        }
    }
    return fail(ZVB_ERR_INVALID, "unknown op");
}

// ------------------------------------------------------------------------------------------ plan
struct zvb_plan {
    int N = 0, T = 0;
    int D = 0, in_dim = 0, out_dim = 0, xin_pitch = 0, has_time = 0, has_g = 0;
    zvb_io io{};
    std::vector<Op> ops;
    unsigned long long* sat_counter = nullptr;      // device counter (caller owned), see zvb_plan_set_saturation_counter
    // argument blocks of the fused prologue kernels (ops point at them: the plan lives on the heap and is never moved)
    MaskJobs mjobs{};
    SmallJobs sjobs{};
    RowBiasJobs rjobs{};
};

struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<uint8_t*>(b)) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~static_cast<size_t>(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

static inline int round8(int x) { return (x + 7) / 8 * 8; }

static Op small_op(const float* in, const float* W, const float* b, const float* addend, float* out, int N, int K,
                   int O, int act_in, int act_out) {
    Op op; op.type = OP_SMALL;
    op.f0 = in; op.f1 = W; op.f2 = b; op.f3 = addend; op.o0 = out;
    op.i0 = N; op.i1 = K; op.i2 = O; op.i3 = act_in; op.i4 = act_out;
    return op;
}

// Builds the plan; with ws == nullptr only measures the workspace.
//
// Residual stream: ONE fp16 tensor per stage, read by the next tensor-core kernel as its A operand and
// by the next residual-adding epilogue through the TMA aux ring (4 bytes of HBM traffic per element and
// update, against 10 for an fp32 stream with a 16-bit operand shadow).
static int build_plan(const zvb_model* m, int N, int T, void* ws, size_t* bytes_out, zvb_plan* plan) {
    PlanBuildScope weights_are_b;
    if (m == nullptr || m->abi_version != ZVB_ABI_VERSION) return fail(ZVB_ERR_INVALID, "model description: ABI version mismatch");
    if (N <= 0 || T <= 0) return fail(ZVB_ERR_INVALID, "N and T must be positive");
    const int D = m->dim, H = m->num_heads, dv = m->value_head_dim;
    if (D % 64 != 0 || D > 1024) return fail(ZVB_ERR_INVALID, "dim %d unsupported (multiple of 64, <= 1024)", D);
    if (m->num_stacks <= 0 || m->num_stacks > ZVB_MAX_STACKS) return fail(ZVB_ERR_INVALID, "bad stack count");
    const bool building = ws != nullptr;
    const int hp = (dv + 15) / 16 * 16;
    const int attn_w = H * (2 * 32 + 4);
    const int nah = m->na_hidden;
    const int ffmax = std::max(std::max(m->ff_dims[0] + H * (2 * 32 + 4), m->ff_dims[1]), m->ff_dims[2]);
    const int xin_pitch = round8(m->in_dim);
    const long long M = (long long)N * T;

    Carver c(ws);
    h16* xin = c.take<h16>(M * xin_pitch);
    float* tbuf = c.take<float>(N);
    float* gbuf = c.take<float>(N);
    uint8_t* mask = c.take<uint8_t>(M);
    float* out = c.take<float>(M * m->out_dim);
    // time embedding chain
    const int td = m->time_dim;
    // guidance_scale_embed has its own input width (reference: modules/zipformer.py:128, 233-238)
    const int gd = m->guidance_dim > 0 ? m->guidance_dim : td;
    float *te0 = nullptr, *teg = nullptr, *te1 = nullptr, *te2 = nullptr, *te3 = nullptr;
    float* temb[ZVB_MAX_STACKS] = {};
    if (td > 0) {
        te0 = c.take<float>((size_t)N * td);
        teg = c.take<float>((size_t)N * gd);
        te1 = c.take<float>((size_t)N * 2 * td);
        te2 = c.take<float>((size_t)N * td);
        te3 = c.take<float>((size_t)N * td);
        for (int s = 0; s < m->num_stacks; ++s) temb[s] = c.take<float>((size_t)N * D);
    }
    h16* cur0 = c.take<h16>(M * D);             // full-rate stream (ping-pong)
    h16* cur1 = c.take<h16>(M * D);
    h16* S[2] = {c.take<h16>(M * D), c.take<h16>(M * D)};        // layer inputs inside a stack
    const bool merged = g_merge_ff1 && m->layers[0].ff1_attn.w != nullptr;
    h16* St[2] = {nullptr, nullptr};                             // src + temb (only when the projections are not merged)
    if (!merged) { St[0] = c.take<h16>(M * D); St[1] = c.take<h16>(M * D); }
    h16* R[2] = {c.take<h16>(M * D), c.take<h16>(M * D)};        // stream inside a layer
    h16* qkp = c.take<h16>(M * attn_w);
    // merged feed_forward1 / attention in-projection (zvb_layer::ff1_attn): fp16 time embedding per stack and the
    // per-utterance row bias W1 * temb (fp32, pitch = merged width; the attention columns stay zero)
    const int mw = m->ff_dims[0] + attn_w;
    // fused prologue (g_fuse_prologue): every layer's row bias is computed up front by ONE kernel into its own buffer;
    // otherwise per layer by a 128-row GEMM over the fp16 time embedding into one shared buffer
    const bool fused_rb = merged && td > 0 && g_fuse_prologue && D <= 512 && D % 8 == 0 && m->num_layers <= RB_MAX_JOBS;
    h16* tembh[ZVB_MAX_STACKS] = {};
    float* rb1 = nullptr;
    std::vector<float*> rbl(m->num_layers > 0 ? m->num_layers : 1, nullptr);
    if (merged && td > 0) {
        if (fused_rb) {
            for (int l = 0; l < m->num_layers; ++l) rbl[l] = c.take<float>((size_t)N * mw);
        } else {
            for (int s = 0; s < m->num_stacks; ++s) tembh[s] = c.take<h16>((size_t)N * D);
            rb1 = c.take<float>((size_t)N * mw);
        }
    }
    h16* hid = c.take<h16>(M * ffmax);
    h16* nay = c.take<h16>(M * nah);
    h16* pvna = c.take<h16>(M * nah);
    h16* pvsa = c.take<h16>(M * H * dv);
    h16* glu = c.take<h16>(M * D);
    h16* cv = c.take<h16>(M * D);
    const int Lk0 = round8(T);
    h16* P = c.take<h16>((size_t)N * H * T * Lk0);
    float* invl = c.take<float>((size_t)N * H * T);
    // per-resolution buffers (pads of the transposed V stay zero for the lifetime of the plan)
    h16 *vtna[5] = {}, *vtsa[5] = {};
    uint8_t* mask_ds[5] = {};
    uint32_t* maskw_ds[5] = {};
    for (int ds = 1; ds <= 4; ds *= 2) {
        bool used = false;
        for (int s = 0; s < m->num_stacks; ++s) used |= m->stacks[s].downsample == ds;
        if (!used) continue;
        const int L = (T + ds - 1) / ds, Lk = round8(L);
        vtna[ds] = c.take<h16>((size_t)N * nah * Lk);
        vtsa[ds] = c.take<h16>((size_t)N * H * hp * Lk);
        mask_ds[ds] = ds == 1 ? mask : c.take<uint8_t>((size_t)N * L);
        maskw_ds[ds] = c.take<uint32_t>((size_t)N * attn_mask_words(L));
    }
    if (bytes_out) *bytes_out = (c.off + 255) & ~static_cast<size_t>(255);
    if (!building) return 0;

    plan->N = N; plan->T = T; plan->D = D; plan->in_dim = m->in_dim; plan->out_dim = m->out_dim;
    plan->xin_pitch = xin_pitch; plan->has_time = td > 0; plan->has_g = m->use_guidance_embed;
    plan->io.xin = xin; plan->io.t = tbuf; plan->io.g = gbuf; plan->io.mask = mask; plan->io.out = out;
    plan->io.xin_pitch = xin_pitch;
    std::vector<Op>& ops = plan->ops;

    // strided masks + excluded-key bit words
    if (g_fuse_prologue) {
        MaskJobs& mj = plan->mjobs;
        mj.n = 0;
        for (int ds = 1; ds <= 4; ds *= 2) {
            if (maskw_ds[ds] == nullptr) continue;
            MaskJob& jb = mj.j[mj.n++];
            jb.ds = ds; jb.Ld = (T + ds - 1) / ds; jb.nwords = attn_mask_words(jb.Ld);
            jb.strided = ds == 1 ? nullptr : mask_ds[ds];
            jb.words = maskw_ds[ds];
        }
        Op op; op.type = OP_PMASKS; op.p0 = mask; op.p1 = &plan->mjobs; op.i0 = N; op.i1 = T;
        ops.push_back(op);
    } else {
        for (int ds = 2; ds <= 4; ds *= 2) {
            if (mask_ds[ds] == nullptr) continue;
            Op op; op.type = OP_MASK; op.p0 = mask; op.o0 = mask_ds[ds];
            op.i0 = N; op.i1 = T; op.i2 = (T + ds - 1) / ds; op.i3 = ds;
            ops.push_back(op);
        }
        for (int ds = 1; ds <= 4; ds *= 2)
            if (maskw_ds[ds] != nullptr) ops.push_back(mask_words_op(mask_ds[ds], maskw_ds[ds], N, (T + ds - 1) / ds));
    }
    // time embedding chain (reference: modules/zipformer.py:267-278, 676-680, 727-729)
    if (td > 0) {
        Op e; e.type = OP_TSEMB; e.f0 = tbuf; e.o0 = te0; e.i0 = N; e.i1 = td;
        ops.push_back(e);
        const float* t_in = te0;
        if (m->use_guidance_embed) {
            if (m->guidance_w == nullptr || gd % 2 != 0) return fail(ZVB_ERR_INVALID, "guidance embedding: weight missing or odd width %d", gd);
            Op eg; eg.type = OP_TSEMB; eg.f0 = gbuf; eg.o0 = teg; eg.i0 = N; eg.i1 = gd;
            ops.push_back(eg);
            ops.push_back(small_op(teg, m->guidance_w, nullptr, te0, te2, N, gd, td, 0, 0));   // te2 = te0 + Wg*emb(g)
            t_in = te2;
        }
        ops.push_back(small_op(t_in, m->time0_w, m->time0_b, nullptr, te1, N, td, 2 * td, 0, ACT_SWOOSH_R_));
        ops.push_back(small_op(te1, m->time2_w, m->time2_b, nullptr, te3, N, 2 * td, td, 0, 0));
        if (g_fuse_prologue && m->num_stacks <= 8) {       // the per-stack projections in one launch
            SmallJobs& sj = plan->sjobs;
            sj.n = m->num_stacks;
            for (int s = 0; s < m->num_stacks; ++s) sj.j[s] = SmallJob{m->stacks[s].time_w, m->stacks[s].time_b, temb[s]};
            Op op; op.type = OP_MSMALL; op.f0 = te3; op.p1 = &plan->sjobs; op.i0 = N; op.i1 = td; op.i2 = D;
            ops.push_back(op);
        } else {
            for (int s = 0; s < m->num_stacks; ++s)
                ops.push_back(small_op(te3, m->stacks[s].time_w, m->stacks[s].time_b, nullptr, temb[s], N, td, D,
                                       ACT_SWOOSH_R_, 0));
        }
        if (fused_rb) {                                     // W1 * temb of every layer in one launch
            RowBiasJobs& rj = plan->rjobs;
            rj.n = 0;
            for (int s = 0; s < m->num_stacks; ++s)
                for (int j = 0; j < m->stacks[s].num_layers; ++j) {
                    const int l = m->stacks[s].first_layer + j;
                    rj.j[rj.n++] = RowBiasJob{static_cast<const h16*>(m->layers[l].ff1_attn.w), temb[s], rbl[l]};
                }
            Op op; op.type = OP_ROWBIAS; op.p1 = &plan->rjobs; op.i0 = N; op.i1 = D; op.i2 = m->ff_dims[0];
            op.i3 = m->layers[0].ff1_attn.k_pitch; op.i4 = mw;
            ops.push_back(op);
        } else if (merged) {
            for (int s = 0; s < m->num_stacks; ++s) {
                Op cst; cst.type = OP_CAST; cst.p0 = temb[s]; cst.o0 = tembh[s]; cst.rows = N; cst.i0 = D; cst.i1 = D;
                ops.push_back(cst);
            }
        }
    }
    // in_proj (reference: modules/zipformer.py:264-265) -> stream
    {
        Op op; LinearEpi e;
        TRY(build_linear(op, xin, M, xin_pitch, m->in_proj, cur0, D, e));
        ops.push_back(op);
    }
    h16* cur = cur0;
    h16* cur_alt = cur1;

    for (int s = 0; s < m->num_stacks; ++s) {
        const zvb_stack& stk = m->stacks[s];
        const int ds = stk.downsample;
        if (ds != 1 && ds != 2 && ds != 4) return fail(ZVB_ERR_INVALID, "downsample %d unsupported", ds);
        if (stk.num_layers <= 0) return fail(ZVB_ERR_INVALID, "empty stack");
        const int L = (T + ds - 1) / ds, Lk = round8(L);
        const long long Ms = (long long)N * L;
        const float* tb = td > 0 ? temb[s] : nullptr;
        int si = 0;
        if (ds != 1) {   // ds == 1 runs the layers directly on `cur` as the first layer's src
            Op op; op.type = OP_DOWN; op.p0 = cur; op.o0 = S[0];
            op.i0 = N; op.i1 = T; op.i2 = L; op.i3 = ds; op.i4 = D;
            for (int k = 0; k < 4; ++k) op.w[k] = stk.ds_weights[k];
            op.cat = ZVB_CAT_RESAMPLE; op.work = 2.0 * ((double)M + (double)Ms) * D;
            mark_out(op, 0, S[0], Ms * D);
            ops.push_back(op);
        }
        const h16* src = ds == 1 ? cur : S[0];
        if (tb != nullptr && !merged) {   // time-embedded copy of the stack input: St[0] = src + temb
            Op op; op.type = OP_PREP; op.p0 = src; op.o0 = St[0];
            op.f0 = tb; op.i0 = D; op.i1 = L; op.rows = Ms;
            op.cat = ZVB_CAT_ELEMENTWISE; op.work = (double)Ms * D * (2.0 + 2.0);
            mark_out(op, 0, St[0], Ms * D);
            ops.push_back(op);
        }
        const h16* srct = (tb != nullptr && !merged) ? St[0] : src;
        for (int j = 0; j < stk.num_layers; ++j) {
            const zvb_layer& ly = m->layers[stk.first_layer + j];
            const bool last = j == stk.num_layers - 1;
            Op op;
            LinearEpi e;
            auto stream_epi = [&](const h16* resid) {
                LinearEpi x; x.resid = resid;
                return x;
            };
            auto t_epi = [&](int batch_rows, int hd_, int hp_) {
                LinearEpi x; x.out_mode = OUT_T_H16; x.t_L = L; x.t_pitch = Lk; x.t_batch_rows = batch_rows;
                x.t_hd = hd_; x.t_hp = hp_;
                return x;
            };
            const h16* qkp_l = qkp;
            int qkp_ld = attn_w, hid_ld = m->ff_dims[0];
            if (merged) {
                // 1+2. ONE GEMM over the layer input for feed_forward1.in_proj (SwooshL) and the attention projections (no
                // activation): hid = [SwooshL(W1 src + W1 temb + b1) | q k p].  feed_forward1 runs on src + temb
                // (reference: zipformer.py:532-536): the time embedding enters as the per-utterance row bias W1 temb.
                const float* rb_l = fused_rb ? rbl[stk.first_layer + j] : rb1;
                if (tb != nullptr && !fused_rb) {
                    zvb_linear w1 = ly.ff1_attn;
                    w1.b = nullptr; w1.out_features = m->ff_dims[0]; w1.rows = m->ff_dims[0];
                    e = LinearEpi(); e.out_mode = OUT_F32;
                    TRY(build_linear(op, tembh[s], N, D, w1, rb1, mw, e)); op.cat = ZVB_CAT_OTHER; ops.push_back(op);
                }
                e = LinearEpi(); e.act = ACT_SWOOSH_L; e.act_cols = m->ff_dims[0];
                if (tb != nullptr) { e.rowbias = rb_l; e.rows_per_group = L; }
                TRY(build_linear(op, src, Ms, D, ly.ff1_attn, hid, mw, e)); ops.push_back(op);
                qkp_l = hid + m->ff_dims[0]; qkp_ld = mw; hid_ld = mw;
            } else {
                // 1. attention projections (on the un-time-embedded input)
                e = LinearEpi();
                TRY(build_linear(op, src, Ms, D, ly.attn_in, qkp, attn_w, e)); ops.push_back(op);
            }
            if (g_attn_tc && ly.pos_table_tc != nullptr)
                TRY(build_attn_tc(op, qkp_l, qkp_ld, ly.pos_table_tc, maskw_ds[ds], P, invl, N, H, L, Lk));
            else
                TRY(build_attn(op, qkp_l, qkp_ld, ly.pos_table, maskw_ds[ds], P, invl, N, H, L, Lk));
            ops.push_back(op);
            // 2. feed_forward1 on src + temb:  R0 = src + temb + FF1(src + temb)
            if (!merged) {
                e = LinearEpi(); e.act = ACT_SWOOSH_L;
                TRY(build_linear(op, srct, Ms, D, ly.ff_in[0], hid, m->ff_dims[0], e)); ops.push_back(op);
            }
            e = stream_epi(src); e.rowbias = tb; e.rows_per_group = L;
            TRY(build_linear(op, hid, Ms, hid_ld, ly.ff_out[0], R[0], D, e)); ops.push_back(op);
            // 3. nonlin attention
            e = t_epi(nah, 1, 1);
            TRY(build_gated(op, R[0], Ms, D, ly.na_sx, nah, GATE_TANH_SX, vtna[ds], 0, nullptr, e)); ops.push_back(op);
            e = LinearEpi();
            TRY(build_linear(op, R[0], Ms, D, ly.na_y, nay, nah, e)); ops.push_back(op);
            TRY(build_pv(op, P, invl, vtna[ds], pvna, nah, N, H, L, Lk, nah, nah, 0, nay, nah)); ops.push_back(op);
            e = stream_epi(R[0]);
            TRY(build_linear(op, pvna, Ms, nah, ly.na_out, R[1], D, e)); ops.push_back(op);
            // 4. self_attn1 (+ temb for the conv module that follows)
            e = t_epi(H * hp, dv, hp);
            TRY(build_linear(op, R[1], Ms, D, ly.sa_in[0], vtsa[ds], 0, e)); ops.push_back(op);
            TRY(build_pv(op, P, invl, vtsa[ds], pvsa, H * dv, N, H, L, Lk, dv, hp, 1, nullptr, 0)); ops.push_back(op);
            e = stream_epi(R[1]); e.rowbias = tb; e.rows_per_group = L;
            TRY(build_linear(op, pvsa, Ms, H * dv, ly.sa_out[0], R[0], D, e)); ops.push_back(op);
            // 5. conv_module1
            e = LinearEpi();
            TRY(build_gated(op, R[0], Ms, D, ly.conv_in[0], D, GATE_GLU_XS, glu, D, mask_ds[ds], e)); ops.push_back(op);
            { Op d; TRY(build_dwconv(d, glu, cv, ly.dw_w[0], ly.dw_b[0], N, L, D, stk.conv_kernel)); ops.push_back(d); }
            e = stream_epi(R[0]);
            TRY(build_linear(op, cv, Ms, D, ly.conv_out[0], R[1], D, e)); ops.push_back(op);
            // 6. feed_forward2 + bypass_mid
            e = LinearEpi(); e.act = ACT_SWOOSH_L;
            TRY(build_linear(op, R[1], Ms, D, ly.ff_in[1], hid, m->ff_dims[1], e)); ops.push_back(op);
            e = stream_epi(R[1]); e.orig = src; e.bypass_scale = ly.bypass_mid_scale;
            TRY(build_linear(op, hid, Ms, m->ff_dims[1], ly.ff_out[1], R[0], D, e)); ops.push_back(op);
            // 7. self_attn2 (+ temb)
            e = t_epi(H * hp, dv, hp);
            TRY(build_linear(op, R[0], Ms, D, ly.sa_in[1], vtsa[ds], 0, e)); ops.push_back(op);
            TRY(build_pv(op, P, invl, vtsa[ds], pvsa, H * dv, N, H, L, Lk, dv, hp, 1, nullptr, 0)); ops.push_back(op);
            e = stream_epi(R[0]); e.rowbias = tb; e.rows_per_group = L;
            TRY(build_linear(op, pvsa, Ms, H * dv, ly.sa_out[1], R[1], D, e)); ops.push_back(op);
            // 8. conv_module2
            e = LinearEpi();
            TRY(build_gated(op, R[1], Ms, D, ly.conv_in[1], D, GATE_GLU_XS, glu, D, mask_ds[ds], e)); ops.push_back(op);
            { Op d; TRY(build_dwconv(d, glu, cv, ly.dw_w[1], ly.dw_b[1], N, L, D, stk.conv_kernel)); ops.push_back(d); }
            e = stream_epi(R[1]);
            TRY(build_linear(op, cv, Ms, D, ly.conv_out[1], R[0], D, e)); ops.push_back(op);
            // 9. feed_forward3
            e = LinearEpi(); e.act = ACT_SWOOSH_L;
            TRY(build_linear(op, R[0], Ms, D, ly.ff_in[2], hid, m->ff_dims[2], e)); ops.push_back(op);
            e = stream_epi(R[0]);
            TRY(build_linear(op, hid, Ms, m->ff_dims[2], ly.ff_out[2], R[1], D, e)); ops.push_back(op);
            // 10. BiasNorm + bypass -> next layer input and its time-embedded copy
            h16* nsrc = (last && ds == 1) ? cur_alt : S[si ^ 1];     // never aliases `src`
            h16* nsrct = (!last && tb != nullptr && !merged) ? St[si ^ 1] : nullptr;
            { Op b; b.type = OP_BIASNORM; b.p0 = R[1]; b.p1 = src; b.o0 = nsrc; b.o1 = nsrct;
              b.f0 = ly.norm_bias; b.f1 = ly.norm_log_scale; b.f2 = ly.bypass_scale; b.f3 = tb;
              b.i0 = D; b.i1 = L; b.rows = Ms;
              b.cat = ZVB_CAT_BIASNORM;
              b.work = (double)Ms * D * (3 * 2.0 + (nsrct != nullptr ? 2.0 : 0.0));
              mark_out(b, 0, nsrc, Ms * D);
              if (nsrct != nullptr) mark_out(b, 1, nsrct, Ms * D);
              ops.push_back(b); }
            src = nsrc;
            srct = nsrct != nullptr ? nsrct : nsrc;
            si ^= 1;
        }
        if (ds == 1) {
            std::swap(cur, cur_alt);        // the last layer wrote into cur_alt
        } else {
            Op op; op.type = OP_UP; op.p0 = cur; op.p1 = src; op.o0 = cur_alt;
            op.f0 = stk.out_combiner_scale;
            op.i0 = N; op.i1 = T; op.i2 = L; op.i3 = ds; op.i4 = D;
            op.cat = ZVB_CAT_RESAMPLE; op.work = 2.0 * (2.0 * (double)M + (double)Ms) * D;
            mark_out(op, 0, cur_alt, M * D);
            ops.push_back(op);
            std::swap(cur, cur_alt);
        }
    }
    // out_proj (reference: modules/zipformer.py:291)
    {
        Op op; LinearEpi e; e.out_mode = OUT_F32;
        TRY(build_linear(op, cur, M, D, m->out_proj, out, m->out_dim, e));
        ops.push_back(op);
    }
    return 0;
}


// ------------------------------------------------------------------------------------------ vocoder plan (§8 f1)
struct zvb_vocoder_plan {
    int N = 0, T = 0, n_mels = 0, hop = 0;
    std::vector<Op> ops;
    int i_mask = -1, i_window = -1, i_ola = -1;      // ops whose operands are the call's own buffers
};

// ops of one `vocoder.decode` over N utterances of <= T frames; with ws == nullptr only measures the workspace
static int build_vocoder(const zvb_vocoder* v, int N, int T, void* ws, size_t* bytes_out, zvb_vocoder_plan* plan) {
    PlanBuildScope weights_are_b;
    if (v == nullptr || v->abi_version != ZVB_ABI_VERSION) return fail(ZVB_ERR_INVALID, "vocoder description: ABI version mismatch");
    if (N <= 0 || T <= 1) return fail(ZVB_ERR_INVALID, "vocoder: N must be positive and T at least 2");
    if (v->n_fft != AUD_NFFT || v->kernel != 7) return fail(ZVB_ERR_INVALID, "vocoder: built for n_fft 1024 and 7-tap convolutions");
    if (v->dim % 256 != 0 || v->dim > 1024) return fail(ZVB_ERR_INVALID, "vocoder: dim %d unsupported (multiple of 256, <= 1024)", v->dim);
    if (v->n_layers <= 0 || v->n_layers > ZVB_VOC_MAX_LAYERS) return fail(ZVB_ERR_INVALID, "vocoder: bad layer count");
    if (v->hop <= 0 || AUD_NFFT % v->hop != 0) return fail(ZVB_ERR_INVALID, "vocoder: hop must divide n_fft");
    if (v->head.out_features != AUD_NFFT + 2) return fail(ZVB_ERR_INVALID, "vocoder: head must have n_fft + 2 outputs");
    const int D = v->dim, I = v->intermediate;
    const long long M = (long long)N * T;
    const int ldA = v->embed.k_pitch;
    if (ldA < v->kernel * v->n_mels || ldA % 8 != 0) return fail(ZVB_ERR_INVALID, "vocoder: embed pitch %d too small", ldA);
    const int ldS = (AUD_NFFT + 2 + 3) / 4 * 4;
    Carver c(ws);
    uint8_t* mask = c.take<uint8_t>(M);
    h16* A0 = c.take<h16>(M * ldA);
    h16* X[2] = {c.take<h16>(M * D), c.take<h16>(M * D)};
    h16* Y = c.take<h16>(M * D);
    h16* Y2 = c.take<h16>(M * D);
    h16* Hd = c.take<h16>(M * I);
    float* S = c.take<float>(M * ldS);
    float* frames = c.take<float>(M * AUD_NFFT);
    if (bytes_out) *bytes_out = (c.off + 255) & ~static_cast<size_t>(255);
    if (ws == nullptr) return 0;
    plan->N = N; plan->T = T; plan->n_mels = v->n_mels; plan->hop = v->hop;
    std::vector<Op>& ops = plan->ops;
    auto ln_op = [&](const h16* x, h16* out, const float* w, const float* b, const uint8_t* m) {
        Op o; o.type = OP_LAYERNORM; o.p0 = x; o.o0 = out; o.f0 = w; o.f1 = b; o.p1 = m; o.rows = M; o.i0 = D; o.w[0] = 1e-6f;
        o.cat = ZVB_CAT_BIASNORM; o.work = 2.0 * 2.0 * (double)M * D;
        return o;
    };
    { Op o; o.type = OP_VOC_MASK; o.o0 = mask; o.i0 = N; o.i1 = T; plan->i_mask = (int)ops.size(); ops.push_back(o); }
    { Op o; o.type = OP_VOC_WINDOW; o.o0 = A0; o.i0 = N; o.i1 = T; o.i2 = v->n_mels; o.i3 = v->kernel; o.i4 = ldA; o.w[0] = 1.0f;
      o.cat = ZVB_CAT_ELEMENTWISE; o.work = (double)M * (v->n_mels * 4.0 + ldA * 2.0);
      plan->i_window = (int)ops.size(); ops.push_back(o); }
    { Op op; LinearEpi e; TRY(build_linear(op, A0, M, ldA, v->embed, Y, D, e)); ops.push_back(op); }
    ops.push_back(ln_op(Y, X[0], v->norm_w, v->norm_b, mask));
    int cur = 0;
    for (int l = 0; l < v->n_layers; ++l) {
        const zvb_voc_layer& ly = v->layers[l];
        { Op d; TRY(build_dwconv(d, X[cur], Y, ly.dw_w, ly.dw_b, N, T, D, v->kernel, 0)); ops.push_back(d); }
        ops.push_back(ln_op(Y, Y2, ly.ln_w, ly.ln_b, nullptr));
        { Op op; LinearEpi e; e.act = ACT_GELU; TRY(build_linear(op, Y2, M, D, ly.pw1, Hd, I, e)); ops.push_back(op); }
        { Op op; LinearEpi e; e.resid = X[cur]; e.row_mask = mask;
          TRY(build_linear(op, Hd, M, I, ly.pw2, X[cur ^ 1], D, e)); ops.push_back(op); }
        cur ^= 1;
    }
    ops.push_back(ln_op(X[cur], Y2, v->final_w, v->final_b, nullptr));
    { Op op; LinearEpi e; e.out_mode = OUT_F32; TRY(build_linear(op, Y2, M, D, v->head, S, ldS, e)); ops.push_back(op); }
    { Op o; o.type = OP_ISTFT_FRAMES; o.p0 = S; o.i0 = ldS; o.p1 = mask; o.f0 = v->window; o.o0 = frames; o.rows = M;
      o.cat = ZVB_CAT_OTHER; o.work = (double)M * (ldS + AUD_NFFT) * 4.0; ops.push_back(o); }
    { Op o; o.type = OP_OLA; o.p0 = frames; o.f0 = v->window; o.i0 = N; o.i1 = T; o.i2 = v->hop; o.i5 = 0;
      o.cat = ZVB_CAT_OTHER; o.work = (double)M * AUD_NFFT * 4.0 + (double)N * v->hop * (T - 1) * 4.0;
      plan->i_ola = (int)ops.size(); ops.push_back(o); }
    return 0;
}

// one CUDA event pair per op (synchronises): shared by zvb_decoder_profile / zvb_vocoder_profile
static int profile_ops(const std::vector<Op>& ops, cudaStream_t st, int max_ops, float* ms, int* category, double* work,
                       double* bytes, int* shapes, int* num_ops) {
    const int n = static_cast<int>(ops.size());
    *num_ops = n;
    if (n > max_ops) return fail(ZVB_ERR_INVALID, "profile buffers too small: %d ops", n);
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(ev[0], st));
    int rc = 0;
    for (int i = 0; i < n && rc == 0; ++i) {
        rc = launch_op(ops[i], st);
        if (rc == 0 && cudaEventRecord(ev[i + 1], st) != cudaSuccess) rc = fail(ZVB_ERR_CUDA, "event record");
    }
    if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(ZVB_ERR_CUDA, "profile: stream sync failed");
    for (int i = 0; i < n && rc == 0; ++i) {
        cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
        category[i] = ops[i].cat;
        work[i] = ops[i].work;
        if (bytes != nullptr) bytes[i] = ops[i].bytes > 0.0 ? ops[i].bytes : ops[i].work;
        if (shapes != nullptr)
            for (int k = 0; k < 4; ++k) shapes[4 * i + k] = ops[i].shape[k];
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* zvb_last_error(void) { return g_err.c_str(); }
int zvb_abi_version(void) { return ZVB_ABI_VERSION; }
long long zvb_launch_count(void) { return g_launches; }

int zvb_plan_workspace_bytes(const zvb_model* model, int N, int T, size_t* bytes) {
    load_switches();
    if (bytes == nullptr) return fail(ZVB_ERR_INVALID, "bytes is null");
    return build_plan(model, N, T, nullptr, bytes, nullptr);
}

int zvb_plan_create(const zvb_model* model, int N, int T, void* workspace, size_t workspace_bytes, zvb_plan** plan) {
    if (plan == nullptr || workspace == nullptr) return fail(ZVB_ERR_INVALID, "null argument");
    TRY(init_device());
    size_t need = 0;
    TRY(build_plan(model, N, T, nullptr, &need, nullptr));
    if (workspace_bytes < need) return fail(ZVB_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
    zvb_plan* p = new zvb_plan();
    int r = build_plan(model, N, T, workspace, nullptr, p);
    if (r != 0) { delete p; return r; }
    *plan = p;
    return 0;
}

void zvb_plan_destroy(zvb_plan* plan) { delete plan; }

int zvb_plan_io(const zvb_plan* plan, zvb_io* io) {
    if (plan == nullptr || io == nullptr) return fail(ZVB_ERR_INVALID, "null argument");
    *io = plan->io;
    return 0;
}

int zvb_decoder_forward(zvb_plan* plan, void* stream) {
    if (plan == nullptr) return fail(ZVB_ERR_INVALID, "null plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (const Op& op : plan->ops) {
        TRY(launch_op(op, st));
        if (plan->sat_counter != nullptr)
            for (int i = 0; i < 2; ++i)
                if (op.scan_ptr[i] != nullptr && op.scan_n[i] > 0) {
                    const long long vec = op.scan_n[i] / 8;
                    long long blocks = (vec + 255) / 256;
                    if (blocks > 8 * g_num_sms) blocks = 8 * g_num_sms;
                    if (blocks < 1) blocks = 1;
                    launch_k(count_saturated_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (const h16*)op.scan_ptr[i],
                             op.scan_n[i], plan->sat_counter);
                    TRY(check_launch("count_saturated"));
                }
    }
    return 0;
}

int zvb_plan_set_saturation_counter(zvb_plan* plan, unsigned long long* counter) {
    if (plan == nullptr) return fail(ZVB_ERR_INVALID, "null plan");
    plan->sat_counter = counter;
    return 0;
}

const char* zvb_source_hash(void) { return g_source_hash + 13; }

int zvb_decoder_profile(zvb_plan* plan, void* stream, int max_ops, float* ms, int* category, double* work,
                        double* bytes, int* shapes, int* num_ops) {
    if (plan == nullptr || ms == nullptr || category == nullptr || work == nullptr || num_ops == nullptr)
        return fail(ZVB_ERR_INVALID, "null argument");
    return profile_ops(plan->ops, static_cast<cudaStream_t>(stream), max_ops, ms, category, work, bytes, shapes, num_ops);
}

int zvb_decoder_forward_f32(zvb_plan* plan, const float* x, const float* t, const uint8_t* mask, const float* g,
                            float* out, void* stream) {
    if (plan == nullptr || x == nullptr || mask == nullptr || out == nullptr) return fail(ZVB_ERR_INVALID, "null argument");
    if (plan->has_time && t == nullptr) return fail(ZVB_ERR_INVALID, "this network needs t");
    if (plan->has_g && g == nullptr) return fail(ZVB_ERR_INVALID, "this network needs guidance_scale");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long M = (long long)plan->N * plan->T;
    const long long n = M * plan->xin_pitch;
    launch_k(cast_pad_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x, (h16*)plan->io.xin, M, plan->in_dim, plan->xin_pitch);
    TRY(check_launch("cast_pad"));
    if (plan->has_time) CUDA_TRY(cudaMemcpyAsync(plan->io.t, t, sizeof(float) * plan->N, cudaMemcpyDeviceToDevice, st));
    if (plan->has_g) CUDA_TRY(cudaMemcpyAsync(plan->io.g, g, sizeof(float) * plan->N, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(plan->io.mask, mask, (size_t)M, cudaMemcpyDeviceToDevice, st));
    TRY(zvb_decoder_forward(plan, stream));
    CUDA_TRY(cudaMemcpyAsync(out, plan->io.out, sizeof(float) * M * plan->out_dim, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// fills a device float array with one value (captured as a kernel so graphs stay replayable)
__global__ void fill_kernel(float* p, const float* src, int idx, int n) {
    pdl_wait();
    pdl_launch();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = src[idx];
}
__global__ void copy_scaled_kernel(float* dst, const float* src, float scale, int n) {
    pdl_wait();
    pdl_launch();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] * scale;
}

int zvb_sample(zvb_plan* plan, float* x, const float* text, const float* speech, const uint8_t* mask,
               const float* guidance, const float* ts, const float* ts_host, int num_step, int mode, int B, int F,
               int Ft, float* vrec, void* stream) {
    if (plan == nullptr || x == nullptr || text == nullptr || speech == nullptr || mask == nullptr || ts == nullptr ||
        ts_host == nullptr)
        return fail(ZVB_ERR_INVALID, "null argument");
    if (mode < 0 || mode > 2) return fail(ZVB_ERR_INVALID, "mode must be 0, 1 or 2");
    const int N = mode == 1 ? 2 * B : B;
    if (N != plan->N) return fail(ZVB_ERR_INVALID, "plan was built for N=%d rows, call needs %d", plan->N, N);
    if (2 * F + Ft != plan->in_dim || F != plan->out_dim)
        return fail(ZVB_ERR_INVALID, "feature dims (%d,%d) do not match the plan (%d,%d)", F, Ft, plan->in_dim, plan->out_dim);
    if (mode != 0 && guidance == nullptr) return fail(ZVB_ERR_INVALID, "guidance is null");
    if ((mode == 2) != (plan->has_g != 0)) return fail(ZVB_ERR_INVALID, "mode %d does not match the network (guidance embed %d)", mode, plan->has_g);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int T = plan->T;
    const long long per_utt = (long long)T * F;
    const size_t mbytes = (size_t)B * T;
    CUDA_TRY(cudaMemcpyAsync(plan->io.mask, mask, mbytes, cudaMemcpyDeviceToDevice, st));
    if (mode == 1) CUDA_TRY(cudaMemcpyAsync(plan->io.mask + mbytes, mask, mbytes, cudaMemcpyDeviceToDevice, st));
    if (mode == 2) {
        launch_k(copy_scaled_kernel, dim3((N + 127) / 128), dim3(128), 0, st, plan->io.g, guidance, 1.0f, N);
        TRY(check_launch("copy_guidance"));
    }
    const long long n_in = (long long)N * T * plan->xin_pitch;
    const long long n_x = (long long)B * per_utt;
    for (int step = 0; step < num_step; ++step) {
        const float t = ts_host[step];
        const int drop_speech = t > 0.5f ? 1 : 0;                 // reference: solver.py:90-98
        launch_k(assemble_input_kernel, dim3((unsigned)((n_in + 255) / 256)), dim3(256), 0, st, 
            x, text, speech, (h16*)plan->io.xin, B, T, F, Ft, plan->xin_pitch, mode == 1, drop_speech);
        TRY(check_launch("assemble_input"));
        launch_k(fill_kernel, dim3((N + 127) / 128), dim3(128), 0, st, plan->io.t, ts, step, N);
        TRY(check_launch("fill_t"));
        TRY(zvb_decoder_forward(plan, stream));
        const float gscale = (mode == 1 && !drop_speech) ? 2.0f : 1.0f;
        launch_k(cfg_euler_kernel, dim3((unsigned)((n_x + 255) / 256)), dim3(256), 0, st, 
            x, plan->io.out, guidance, gscale, ts, step, vrec ? vrec + (long long)step * n_x : nullptr, B, per_utt,
            mode == 1);
        TRY(check_launch("cfg_euler"));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ test entry points
int zvb_test_linear(const void* A, int M, int K, int lda, const void* W, const float* bias, int n_out, int k_pitch,
                    int block_n, int act, const void* resid, const void* orig, const float* bypass_scale, void* out,
                    int ldc, int out_mode, void* stream) {
    TRY(init_device());
    zvb_linear lin{W, bias, n_out, K, k_pitch, n_out};
    Op op; LinearEpi e; e.act = act; e.resid = (const h16*)resid; e.orig = (const h16*)orig; e.bypass_scale = bypass_scale;
    e.out_mode = out_mode;
    e.block_n = block_n;
    TRY(build_linear(op, (const h16*)A, M, lda, lin, out, ldc, e));
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_attn_weights(const void* qkp, int ld, const void* pos_table, const uint8_t* mask, void* scratch, void* P,
                          float* inv_l, int N, int H, int L, int Lk, void* stream) {
    TRY(init_device());
    if (scratch == nullptr) return fail(ZVB_ERR_INVALID, "attn: scratch (N * 4 * ceil(L/128) words) is null");
    TRY(launch_op(mask_words_op(mask, (uint32_t*)scratch, N, L), static_cast<cudaStream_t>(stream)));
    Op op;
    TRY(build_attn(op, (const h16*)qkp, ld, pos_table, (const uint32_t*)scratch, (h16*)P, inv_l, N, H, L, Lk));
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_attn_weights_tc(const void* qkp, int ld, const void* pos_table_tc, const uint8_t* mask, void* scratch, void* P,
                             float* inv_l, int N, int H, int L, int Lk, void* stream) {
    TRY(init_device());
    if (scratch == nullptr) return fail(ZVB_ERR_INVALID, "attn: scratch (N * 4 * ceil(L/128) words) is null");
    TRY(launch_op(mask_words_op(mask, (uint32_t*)scratch, N, L), static_cast<cudaStream_t>(stream)));
    Op op;
    TRY(build_attn_tc(op, (const h16*)qkp, ld, pos_table_tc, (const uint32_t*)scratch, (h16*)P, inv_l, N, H, L, Lk));
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_pv(const void* P, const float* inv_l, const void* Vt, void* out, int N, int H, int L, int Lk, int hd, int hp,
                int per_head, const void* mul, void* stream) {
    TRY(init_device());
    Op op;
    const int ldc = per_head ? H * hd : hd;
    TRY(build_pv(op, (const h16*)P, inv_l, (const h16*)Vt, out, ldc, N, H, L, Lk, hd, hp, per_head, (const h16*)mul, hd));
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_gated(const void* A, int M, int K, int lda, const void* W, const float* bias, int rows, int n_out,
                   int k_pitch, int gate_mode, const uint8_t* row_mask, void* out, int ldc, void* stream) {
    TRY(init_device());
    zvb_linear lin{W, bias, n_out, K, k_pitch, rows};
    Op op; LinearEpi e;
    TRY(build_gated(op, (const h16*)A, M, lda, lin, n_out, gate_mode, out, ldc, row_mask, e));
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_biasnorm_bypass(const void* src, const void* orig, void* out, void* out_t,
                             const float* temb, int rows_per_group, const float* nbias, const float* log_scale,
                             const float* bscale, long long rows, int C, void* stream) {
    TRY(init_device());
    if (C % 8 != 0 || C > 1024) return fail(ZVB_ERR_INVALID, "biasnorm: C must be a multiple of 8, <= 1024");
    Op b; b.type = OP_BIASNORM; b.p0 = src; b.p1 = orig; b.o0 = out; b.o1 = out_t;
    b.f0 = nbias; b.f1 = log_scale; b.f2 = bscale; b.f3 = temb; b.i0 = C; b.i1 = rows_per_group; b.rows = rows;
    return launch_op(b, static_cast<cudaStream_t>(stream));
}

int zvb_test_dwconv(const void* x, void* out, const float* wt, const float* bias, int N, int L, int C, int K,
                    void* stream) {
    TRY(init_device());
    Op d;
    TRY(build_dwconv(d, (const h16*)x, (h16*)out, wt, bias, N, L, C, K));
    return launch_op(d, static_cast<cudaStream_t>(stream));
}

int zvb_test_linear_t(const void* A, int M, int K, int lda, const void* W, const float* bias, int n_out, int k_pitch,
                      void* out, int t_L, int t_pitch, int t_batch_rows, int t_hd, int t_hp, void* stream) {
    TRY(init_device());
    zvb_linear lin{W, bias, n_out, K, k_pitch, n_out};
    Op op; LinearEpi e; e.out_mode = OUT_T_H16;
    e.t_L = t_L; e.t_pitch = t_pitch; e.t_batch_rows = t_batch_rows; e.t_hd = t_hd; e.t_hp = t_hp;
    TRY(build_linear(op, (const h16*)A, M, lda, lin, out, 0, e));
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_downsample(const void* src, void* out, int N, int L, int ds, const float* w4, int C, void* stream) {
    TRY(init_device());
    if (C % 8 != 0 || (ds != 1 && ds != 2 && ds != 4) || w4 == nullptr) return fail(ZVB_ERR_INVALID, "downsample: bad arguments");
    Op op; op.type = OP_DOWN; op.p0 = src; op.o0 = out;
    op.i0 = N; op.i1 = L; op.i2 = (L + ds - 1) / ds; op.i3 = ds; op.i4 = C;
    for (int k = 0; k < 4; ++k) op.w[k] = w4[k];
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_upsample_combine(const void* orig, const void* y, void* out, const float* scale, int N, int L, int ds, int C,
                              void* stream) {
    TRY(init_device());
    if (C % 8 != 0 || (ds != 1 && ds != 2 && ds != 4)) return fail(ZVB_ERR_INVALID, "upsample: bad arguments");
    Op op; op.type = OP_UP; op.p0 = orig; op.p1 = y; op.o0 = out; op.f0 = scale;
    op.i0 = N; op.i1 = L; op.i2 = (L + ds - 1) / ds; op.i3 = ds; op.i4 = C;
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_stream_prep(const void* x, void* xt, const float* temb, int rows_per_group, long long rows, int C, void* stream) {
    TRY(init_device());
    if (C % 8 != 0) return fail(ZVB_ERR_INVALID, "stream_prep: C must be a multiple of 8");
    Op op; op.type = OP_PREP; op.p0 = x; op.o0 = xt; op.f0 = temb; op.i0 = C; op.i1 = rows_per_group; op.rows = rows;
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_assemble_input(const float* x, const float* text, const float* speech, void* xin, int B, int T, int F, int Ft,
                            int ldx, int cfg, int drop_speech, void* stream) {
    TRY(init_device());
    if (ldx < 2 * F + Ft) return fail(ZVB_ERR_INVALID, "assemble: pitch %d < %d", ldx, 2 * F + Ft);
    const long long n_in = (long long)(cfg ? 2 * B : B) * T * ldx;
    launch_k(assemble_input_kernel, dim3((unsigned)((n_in + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             x, text, speech, (h16*)xin, B, T, F, Ft, ldx, cfg, drop_speech);
    return check_launch("assemble_input");
}

int zvb_test_small_linear(const float* in, const float* W, const float* bias, const float* addend, float* out, int N, int K,
                          int O, int act_in, int act_out, void* stream) {
    TRY(init_device());
    return launch_op(small_op(in, W, bias, addend, out, N, K, O, act_in, act_out), static_cast<cudaStream_t>(stream));
}

int zvb_test_timestep_embedding(const float* t, float* out, int N, int dim, void* stream) {
    TRY(init_device());
    if (dim % 2 != 0) return fail(ZVB_ERR_INVALID, "timestep embedding: odd width");
    Op e; e.type = OP_TSEMB; e.f0 = t; e.o0 = out; e.i0 = N; e.i1 = dim;
    return launch_op(e, static_cast<cudaStream_t>(stream));
}

int zvb_test_masks(const uint8_t* mask, int N, int T, int ds, uint8_t* strided, uint32_t* words, void* stream) {
    TRY(init_device());
    const int L = (T + ds - 1) / ds;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint8_t* m = mask;
    if (ds != 1) {
        if (strided == nullptr) return fail(ZVB_ERR_INVALID, "masks: strided output is null");
        Op op; op.type = OP_MASK; op.p0 = mask; op.o0 = strided; op.i0 = N; op.i1 = T; op.i2 = L; op.i3 = ds;
        TRY(launch_op(op, st));
        m = strided;
    }
    if (words != nullptr) TRY(launch_op(mask_words_op(m, words, N, L), st));
    return 0;
}

int zvb_test_cfg_euler(float* x, const float* v, const float* guidance, float gscale, const float* ts, int step, int B,
                       long long per_utt, int cfg, void* stream) {
    const long long n = (long long)B * per_utt;
    launch_k(cfg_euler_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        x, v, guidance, gscale, ts, step, nullptr, B, per_utt, cfg);
    return check_launch("cfg_euler");
}

// ------------------------------------------------------------------------------------------ f3 / f1 entry points
int zvb_fbank(const float* wav, const int32_t* lens, int B, int s_pitch, const float* window, const float* fb,
              const int32_t* fb_range, int n_mels, int hop, float scale, float* out, int T, void* stream) {
    TRY(init_device());
    if (wav == nullptr || lens == nullptr || window == nullptr || fb == nullptr || fb_range == nullptr || out == nullptr)
        return fail(ZVB_ERR_INVALID, "fbank: null argument");
    if (B <= 0 || T <= 0 || n_mels <= 0 || hop <= 0 || s_pitch <= 0) return fail(ZVB_ERR_INVALID, "fbank: bad sizes");
    const long long total = (long long)B * T;
    long long blocks = total < 8LL * g_num_sms ? total : 8LL * g_num_sms;
    launch_k(fbank_kernel, dim3((unsigned)blocks), dim3(AUD_THREADS), (size_t)AUD_FFT_SMEM, static_cast<cudaStream_t>(stream), wav,
             (const int*)lens, B, s_pitch, window, fb, reinterpret_cast<const int2*>(fb_range), n_mels, hop, scale, out, T);
    return check_launch("fbank");
}

int zvb_vocoder_workspace_bytes(const zvb_vocoder* voc, int N, int T, size_t* bytes) {
    load_switches();
    if (bytes == nullptr) return fail(ZVB_ERR_INVALID, "bytes is null");
    return build_vocoder(voc, N, T, nullptr, bytes, nullptr);
}

int zvb_vocoder_create(const zvb_vocoder* voc, int N, int T, void* workspace, size_t workspace_bytes, zvb_vocoder_plan** plan) {
    if (plan == nullptr || workspace == nullptr) return fail(ZVB_ERR_INVALID, "null argument");
    TRY(init_device());
    size_t need = 0;
    TRY(build_vocoder(voc, N, T, nullptr, &need, nullptr));
    if (workspace_bytes < need) return fail(ZVB_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
    zvb_vocoder_plan* p = new zvb_vocoder_plan();
    int r = build_vocoder(voc, N, T, workspace, nullptr, p);
    if (r != 0) { delete p; return r; }
    *plan = p;
    return 0;
}

void zvb_vocoder_destroy(zvb_vocoder_plan* plan) { delete plan; }

int zvb_vocoder_decode(zvb_vocoder_plan* plan, const float* mel, const int32_t* lens, float scale, int clamp, float* wav,
                       void* stream) {
    if (plan == nullptr || mel == nullptr || lens == nullptr || wav == nullptr) return fail(ZVB_ERR_INVALID, "null argument");
    Op& m = plan->ops[plan->i_mask];
    m.p0 = lens;
    Op& w = plan->ops[plan->i_window];
    w.p0 = mel; w.p1 = lens; w.w[0] = scale;
    Op& o = plan->ops[plan->i_ola];
    o.p1 = lens; o.o0 = wav; o.i5 = clamp;
    for (const Op& op : plan->ops) TRY(launch_op(op, static_cast<cudaStream_t>(stream)));
    return 0;
}

int zvb_vocoder_profile(zvb_vocoder_plan* plan, void* stream, int max_ops, float* ms, int* category, double* work,
                        double* bytes, int* num_ops) {
    if (plan == nullptr || ms == nullptr || category == nullptr || work == nullptr || num_ops == nullptr)
        return fail(ZVB_ERR_INVALID, "null argument");
    if (plan->ops[plan->i_window].p0 == nullptr) return fail(ZVB_ERR_INVALID, "vocoder profile: call zvb_vocoder_decode once first");
    return profile_ops(plan->ops, static_cast<cudaStream_t>(stream), max_ops, ms, category, work, bytes, nullptr, num_ops);
}

int zvb_test_dwconv_linear(const void* x, void* out, const float* wt, const float* bias, int N, int L, int C, int K,
                           void* stream) {
    TRY(init_device());
    Op d;
    TRY(build_dwconv(d, (const h16*)x, (h16*)out, wt, bias, N, L, C, K, 0));
    return launch_op(d, static_cast<cudaStream_t>(stream));
}

int zvb_test_layernorm(const void* x, void* out, const float* w, const float* b, const uint8_t* mask, long long rows, int C,
                       float eps, void* stream) {
    TRY(init_device());
    if (C % 256 != 0 || C > 1024) return fail(ZVB_ERR_INVALID, "layernorm: C must be a multiple of 256, <= 1024");
    Op o; o.type = OP_LAYERNORM; o.p0 = x; o.o0 = out; o.f0 = w; o.f1 = b; o.p1 = mask; o.rows = rows; o.i0 = C; o.w[0] = eps;
    return launch_op(o, static_cast<cudaStream_t>(stream));
}

int zvb_test_linear_masked(const void* A, int M, int K, int lda, const void* W, const float* bias, int n_out, int k_pitch,
                           int act, const void* resid, const uint8_t* row_mask, void* out, int ldc, void* stream) {
    TRY(init_device());
    zvb_linear lin{W, bias, n_out, K, k_pitch, n_out};
    Op op; LinearEpi e; e.act = act; e.resid = (const h16*)resid; e.row_mask = row_mask;
    TRY(build_linear(op, (const h16*)A, M, lda, lin, out, ldc, e));
    return launch_op(op, static_cast<cudaStream_t>(stream));
}

int zvb_test_istft(const float* S, int ld, const int32_t* lens, const float* window, float* frames, uint8_t* mask, float* wav,
                   int N, int T, int hop, int clamp, void* stream) {
    TRY(init_device());
    if (T < 2 || hop <= 0 || AUD_NFFT % hop != 0 || ld < AUD_NFFT + 2) return fail(ZVB_ERR_INVALID, "istft: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { Op o; o.type = OP_VOC_MASK; o.p0 = lens; o.o0 = mask; o.i0 = N; o.i1 = T; TRY(launch_op(o, st)); }
    { Op o; o.type = OP_ISTFT_FRAMES; o.p0 = S; o.i0 = ld; o.p1 = mask; o.f0 = window; o.o0 = frames; o.rows = (long long)N * T;
      TRY(launch_op(o, st)); }
    { Op o; o.type = OP_OLA; o.p0 = frames; o.p1 = lens; o.f0 = window; o.o0 = wav; o.i0 = N; o.i1 = T; o.i2 = hop; o.i5 = clamp;
      TRY(launch_op(o, st)); }
    return 0;
}

int zvb_debug_launch_shape(long long rows, int n_out, int k, int lean_kind, int num_sms, long long attn_ctas, int q_tiles,
                           int* block_n, int* pair, int* attn_split) {
    if (rows <= 0 || n_out <= 0 || k <= 0 || num_sms <= 0) return fail(ZVB_ERR_INVALID, "launch shape: sizes must be positive");
    std::lock_guard<std::mutex> lock(g_init_mutex);
    load_switches();
    const int saved = g_num_sms;
    g_num_sms = num_sms;
    const long long m_tiles = (rows + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    const int k_blocks = (k + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    const int bn = pick_block_n(n_out, m_tiles, k_blocks, lean_kind);
    const long long n_tiles = (n_out + bn - 1) / bn;
    const long long slots2 = ((m_tiles + 1) / 2) * n_tiles;
    if (block_n != nullptr) *block_n = bn;
    if (pair != nullptr)      // the conditions of set_grid
        *pair = (g_cluster_ok && bn >= 64 && m_tiles >= 2 && slots2 >= g_num_sms / 4 && k_blocks >= g_pair_min_kb &&
                 m_tiles >= pair_min_mtiles()) ? 1 : 0;
    if (attn_split != nullptr) *attn_split = attn_split_choice(attn_ctas, q_tiles);
    g_num_sms = saved;
    return 0;
}

#ifdef ZVB_TIMELINE
// Debug build only (tools/timeline_c1.py): copies the per-launch stamps of CTA 0 of every GEMM launch and resets the counter.
int zvb_debug_timeline(unsigned long long* out, int max_rows) {
    unsigned int n = 0;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(&n, g_tl_n, sizeof n));
    int rows = static_cast<int>(n < (unsigned)TL_MAX ? n : (unsigned)TL_MAX);
    rows = rows < max_rows ? rows : max_rows;
    if (rows > 0) CUDA_TRY(cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 20 * rows));
    n = 0;
    CUDA_TRY(cudaMemcpyToSymbol(g_tl_n, &n, sizeof n));
    return rows;
}
#endif

}  // extern "C"
