// common.cuh -- shared device/host helpers for the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "mli_b200.h"

namespace mli {

constexpr int kPage = MLI_PAGE_BLOCK_SIZE;  // positions per KV page (reference include/constants.h:12)

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* file, int line);
void count_launch(int n = 1);

#define MLI_CUDA(expr)                                              \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return mli::cuda_fail(_e, __FILE__, __LINE__); \
    } while (0)

#define MLI_LAUNCH_CHECK()                                          \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return mli::cuda_fail(_e, __FILE__, __LINE__); \
        mli::count_launch();                                        \
    } while (0)

#define MLI_REQUIRE(cond, msg)                                      \
    do {                                                            \
        if (!(cond)) {                                              \
            mli::set_error(std::string("argument error: ") + (msg)); \
            return MLI_ERR_ARG;                                     \
        }                                                           \
    } while (0)

// every C-ABI entry point that takes a context makes the context's device current first (one process
// may drive several GPUs, each through its own context)
#define MLI_ENTER(ctx, msg)                                          \
    do {                                                            \
        MLI_REQUIRE((ctx) != nullptr, msg);                         \
        int _dev = -1;                                              \
        if (cudaGetDevice(&_dev) != cudaSuccess || _dev != (ctx)->device) \
            MLI_CUDA(cudaSetDevice((ctx)->device));                 \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- page addressing (reference include/utils.h:32-60) -------------------------------------
// fp32 format (the reference's, API-visible): element (row r, position j, sub-row off, column c) =
//   page_table[r*W + j/16][(j%16)*3*d + off*d + c]
// compact format (MLI_OPT_KV_FORMAT = 1, opt-in): a position is [inp fp32 x d | K bf16 x d | V bf16 x d]
// = 8*d bytes instead of 12*d; K starts d floats into the position in BOTH formats, K|V stay adjacent.
__device__ __host__ __forceinline__ size_t page_pos_floats(int d, int kv_bf16) {
    return kv_bf16 ? 2 * (size_t)d : 3 * (size_t)d;
}
// start of sub-row `off` (0 inp, 1 K, 2 V) of position j; for bf16 K / V the result is the byte
// address of a bf16 row, still typed float*
__device__ __forceinline__ float* page_row_ptr(float* page, int j, int d, int off, int kv_bf16 = 0) {
    float* pos = page + (size_t)(j & (kPage - 1)) * page_pos_floats(d, kv_bf16);
    if (!kv_bf16) return pos + (size_t)off * d;
    return pos + (off == 0 ? 0 : (off == 1 ? d : d + d / 2));
}

// ---- PTX wrappers: mbarrier + bulk async copy (TMA, non-tensor form) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    // make barrier initialisation visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Waits in these kernels are bounded by one pipeline stage (microseconds).  A wait that lasts
// ~2 s can only be a protocol bug; trap so the launch fails loudly instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("mli: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x);
            __trap();
        }
    }
}

// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// L2 eviction-priority policies for streaming (read-once) and for resident (re-read every step) data
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// the same bulk copy with an L2 cache hint
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                              uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ int atom_add_acq_rel_gpu(int* addr, int v) {
    int old;
    asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// programmatic dependent launch: everything before griddep_wait() may overlap the tail of the
// previous kernel in the stream; nothing that kernel wrote may be read before it.  Both are no-ops
// for a kernel launched without the attribute.  EVERY kernel of a PDL chain must execute the wait
// on every path (completion of a kernel must imply completion of its predecessors).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// optional in-graph step timeline (tools/step_timeline.py): trace[0] = steps seen so far,
// trace[1] = capacity in steps, trace[8 + 8*step + slot] = globaltimer (ns) at which kernel `slot`
// of that step had its dependencies satisfied (i.e. its predecessor had completed).  The scheduler
// (slot 0) opens a new step.  trace[2] != 0 (tools/sched_timing.py): the scheduler also records 16
// phase stamps per step after the step table, at trace[8 + 8*capacity + 16*step + phase].
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int slot) {
    if (trace == nullptr) return;
    if ((blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x) != 0) return;
    unsigned long long step = trace[0];
    if (slot == 0) trace[0] = step + 1; else step -= 1;
    if (step < trace[1]) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        trace[8 + 8 * step + slot] = t;
    }
}

// Where a kernel releases its dependents.  The heavy kernels (GEMMs, attention, encoder) do it LATE,
// once their streaming / main loop is over, so that only the dependent's prologue overlaps (with this
// kernel's tail).  Releasing right after the own wait was measured to be slower: the dependent's CTAs
// then sit resident through the whole predecessor and start late themselves (a GEMM that had been
// parked behind the 11 us scheduler ran 5 us longer; late release: 88.7 -> 86.9 us per engine step).
// MLI_EARLY_TRIGGER restores the early release for A/B runs (tools/step_timeline.py).
#ifdef MLI_EARLY_TRIGGER
#define GRIDDEP_TRIGGER_EARLY() mli::griddep_launch_dependents()
#define GRIDDEP_TRIGGER_LATE() ((void)0)
#else
#define GRIDDEP_TRIGGER_EARLY() ((void)0)
#define GRIDDEP_TRIGGER_LATE() mli::griddep_launch_dependents()
#endif

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace mli

// ---- context (shared by all translation units) ---------------------------------------------
namespace mli {
enum WsSlot {
    WS_ATTN_META = 0,  // item lists / per-row prefix for the fused attention
    WS_ATTN_PART,      // split-KV partial accumulators
    WS_ATTN_CNT,       // per-row arrival counters of the fused attention (zero between launches)
    WS_TILES,          // M-tile lists for prefill / encoder
    WS_LOGITS,         // [B,V] logits when the caller does not want them
    WS_QOUT,           // q_output when the caller passes NULL
    WS_ATTN_OUT,       // attention_result when the caller passes NULL
    WS_QKT,            // dense-path score scratch
    WS_TC_CTR,         // tile counters of the persistent tcgen05 GEMM launches (zero between launches)
    WS_NUM_SLOTS
};
}  // namespace mli

struct mli_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    int gemm_mode = 1;          // MLI_OPT_GEMM_MODE
    int tc_available = 0;       // tcgen05 GEMM path usable on this device
    int attn_chunk_pages = 0;   // 0 = auto
    int attn_ctas_per_sm = 0;   // 0 = auto
    int attn_kernel = 0;        // MLI_OPT_ATTN_KERNEL
    int attn_min_dyn = 4096;    // MLI_OPT_ATTN_MIN_DYN
    int attn_lengths_final = 0; // set by the engine around its attention launch: lengths[] may be read ahead of the dependency wait
    void* ws[mli::WS_NUM_SLOTS] = {};
    size_t ws_bytes[mli::WS_NUM_SLOTS] = {};
    unsigned long long* trace = nullptr;  // in-graph step timeline buffer (device), else NULL
    void* tc_dbg = nullptr;     // device buffer for GEMM phase stamps (tools/gemm_timing.py), else NULL
    // last tcgen05 GEMM launch of each kind (latest-token stage, prefill stage, logits, merged engine step):
    // kernel (0 static grid, 1 persistent, 2 CTA pairs, -1 none yet), K split, tile width, decode tile width, K passes
    int tc_last_plan[4][5] = {{-1, 0, 0, 0, 0}, {-1, 0, 0, 0, 0}, {-1, 0, 0, 0, 0}, {-1, 0, 0, 0, 0}};
    int opt_pdl = 1;            // MLI_OPT_PDL
    int kv_bf16 = 0;            // MLI_OPT_KV_FORMAT: 1 = compact pages (K, V in bf16)
    bool use_pdl = false;       // launch_kernel() adds programmatic stream serialization (set by the engine)
    int ws_frozen = 0;          // number of live CUDA graphs (engines) that captured ws pointers
    // when set, the fused decode-attention main kernel is bracketed by these events (profiling)
    cudaEvent_t attn_ev_start = nullptr, attn_ev_stop = nullptr;
    cudaEvent_t gemm_ev_start = nullptr, gemm_ev_stop = nullptr;   // same for the step's merged GEMM
    // kernel -> dynamic shared memory this context's DEVICE has been configured for.  The attribute is
    // per device, so it is tracked per context (one process may drive several GPUs), not per process.
    std::vector<std::pair<const void*, size_t>> smem_attr;
};

namespace mli {
// device buffer of at least `bytes` for `slot` (grows with a synchronous re-allocation; growing
// while frozen is an error because captured graphs hold the old pointer)
int ws_get(mli_ctx* ctx, int slot, size_t bytes, void** out);
// same, and the buffer is zero-filled whenever it is (re)allocated
int ws_get_zeroed(mli_ctx* ctx, int slot, size_t bytes, void** out);
// cudaFuncAttributeMaxDynamicSharedMemorySize >= bytes for `func` on the context's device
int ensure_dyn_smem_impl(mli_ctx* ctx, const void* func, size_t bytes);
template <typename F>
int ensure_dyn_smem(mli_ctx* ctx, F* func, size_t bytes) {
    return ensure_dyn_smem_impl(ctx, reinterpret_cast<const void*>(func), bytes);
}

// launch on the context's stream; with ctx->use_pdl the kernel is allowed to start while its
// predecessor drains (programmatic dependent launch) -- only for kernels that call griddep_wait()
// before touching anything a predecessor produces or still reads
template <typename... KArgs, typename... Args>
int launch_kernel(mli_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->use_pdl ? 1 : 0;
    MLI_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    count_launch();
    return 0;
}
}  // namespace mli
