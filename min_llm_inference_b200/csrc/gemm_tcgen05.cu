// gemm_tcgen05.cu -- the dense contractions of the path on the 5th-gen tensor cores
// (MLI_OPT_GEMM_MODE = 0): latest-token QKV projection, prefill K/V and the logits GEMM.
//
// All three are  D[f, n] = sum_k W[f, k] * X[n, k]  with
//     f  = output feature (rows of the K-major, pre-transposed weight matrix [Wk^T; Wq^T; Wv^T],
//          or rows of emb_table for the logits)           -> UMMA M = 128 (TMEM lanes)
//     n  = activation row (batch row / prompt position)   -> UMMA N = 16..256 (TMEM columns)
//     k  = emb_dim                                        -> 32 floats (= one 128-byte swizzle row)
//          per pipeline stage, UMMA_K = 8 for tf32
// "swap-AB": the (small, ragged) batch dimension is the MMA N, so decode batches of any size fill
// the 128-lane accumulator.  Inputs are fp32; the tensor cores take tf32, so every operand is
// split  x = hi + lo  (hi = rna_tf32(x), lo = rna_tf32(x - hi)) and each k-step issues three
// MMAs, hi*hi + lo*hi + hi*lo, into an fp32 accumulator in TMEM (3xTF32; relative error ~5e-7,
// against 4.9e-4 for plain tf32 -- SURVEY App. C).  Weights are split/transposed once per
// registered weight set and arrive by TMA; activation rows are gathered from the KV pages and split
// inside the kernel (no staging pass through HBM).  At decode sizes these GEMMs are bound by the
// operand bytes one SM can ingest, so the tile is 128 features x up to 256 rows (weights read once
// per 256 rows) and K is split across a thread-block cluster whose partial tiles are summed
// through distributed shared memory in a fixed order.  The epilogue writes straight into the KV
// pages / q_output / logits (the reference's gather + 3 sgemm + save_to_page_table scatter,
// src/kernels/paged_attention_cublas.cu:45-99, is one kernel).
#include "common.cuh"
#include "kernels.h"

#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace mli {

namespace {

constexpr int kBM = 128;       // features per tile (TMEM lanes)
constexpr int kBK = 32;        // fp32 per stage row = 128 B = one swizzle row
constexpr int kUmmaK = 8;      // tf32 MMA K
// The tensor core adds each MMA into the fp32 accumulator with truncation, so one long chain over
// K drifts by ~K/16 ulp (measured 2.4e-5 relative at K = 2048 on all-positive data).  The K loop is
// therefore cut into chains of at most 512 (split-K across the cluster first, then up to 4 TMEM
// accumulators per CTA) which are added with ordinary fp32 adds.

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// the same with an L2 cache hint (the split weights are re-read every step: evict-last)
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                                 uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version[46,48)=1 | layout[61,64)=2)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(const void* smem_tile) {
    const uint64_t addr = (uint64_t)(smem_u32(smem_tile) & 0x3FFFF) >> 4;
    return addr | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=tf32 [7,10)=2, b=tf32 [10,13)=2,
// K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float to_tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// the same rounding (to nearest, ties away from zero, low 13 mantissa bits cleared) with two full-rate integer
// instructions: cvt.rna.tf32.f32 was measured to dominate the activation converters (32 conversions per thread and
// k-block: ~1000 cycles per k-block for a warp, tools/ncu_jobs.py prefill_stamps).  Bit-identical for finite values.
__device__ __forceinline__ float to_tf32_rna_int(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(const float4& v, float4& h, float4& l) {
    h = make_float4(to_tf32_rna_int(v.x), to_tf32_rna_int(v.y), to_tf32_rna_int(v.z), to_tf32_rna_int(v.w));
    l = make_float4(to_tf32_rna_int(v.x - h.x), to_tf32_rna_int(v.y - h.y), to_tf32_rna_int(v.z - h.z),
                    to_tf32_rna_int(v.w - h.w));
}

enum TcMode { TC_LATEST = 0, TC_PREFILL = 1, TC_LOGITS = 2, TC_STEP = 3 };

struct TcArgs {
    int mode;
    int K;                 // emb_dim
    int d;                 // features per weight matrix (emb_dim) -- LATEST / PREFILL
    int n_rows;            // activation rows covered by the grid (host bound)
    const int* n_tiles;    // PREFILL: device count of 64-row tiles (rows = 64 * *n_tiles)
    const TileDesc* tiles; // PREFILL
    float* const* page_table;
    const int* lengths;
    float* q_out;          // LATEST
    float* score;          // LOGITS [n_rows][V]
    const float* dense_src;  // LOGITS: activation rows [n_rows][K]
    const int* act;        // LOGITS (optional) / STEP: compact list of active rows
    const int* counts;     // [0] = active rows, [1] = granules (device-side)
    const TileDesc* gran;  // STEP: 16-position prefill granules of the new rows
    const int* gran_bound; // STEP: per-row prompt length bounding the granules (chunked prefill); NULL = lengths
    int use_gran;
    int kv_bf16;           // compact page format: K and V rows are stored as bf16
    int defer;             // LOGITS: every split rank stores its partial plane (no cross-CTA reduce)
    size_t defer_stride;   // floats between partial planes
    int V, W, B;
    int bn;                // activation rows per tile: multiple of 16, <= 256 (the UMMA N)
    int n_stages;          // smem pipeline depth
    int n_acc;             // TMEM accumulators the CTA's K range is cut into
    int acc_stride;        // TMEM columns between accumulators (bn rounded up to 32)
    int tmem_cols;         // power of two >= n_acc * acc_stride
    long long* dbg;        // optional phase stamps
    unsigned long long* trace;  // optional step timeline
    int trace_slot;
    // persistent launches (bulk plans: split == 1, no deferred reduce): grid = one CTA per SM, work items
    // (feature tile, activation tile) handed out by an atomic counter -- see tc_item()
    int bn_decode;         // pair kernel: tile width (rows) of a launch WITHOUT prefill granules, chosen on the host so
                           // that the items fill the CTA pairs in whole rounds; 0 = 256
    int dyn;               // 1 = persistent / dynamic items
    int m_tiles;           // feature tiles of the operand
    int* ctr;              // [2]: next item, CTAs finished (both back at zero when the launch ends)
    int kv_group;          // pair kernel: feature pairs per group in the order of the prefill items (0 = all)
    int n_pass;            // pair kernel: K is walked in n_pass passes of K / n_pass (each through the TMEM accumulators,
                           // fp32 chains <= 1024); pass p > 0 adds its sums to what pass p - 1 stored.  <= 1: one pass
};

// optional phase stamps (clock64 of one thread per role) for tools/gemm_timing.py: [cta][8]
//   0 start, 1 setup done, 2 first MMA issued, 3 converters done, 4 accumulators complete,
//   5 after exchange barrier, 6 after reduce, 7 end
#define TC_STAMP(slot)                                                                         \
    do {                                                                                       \
        if (args.dbg != nullptr && lane == 0 && (warp == 4 || (slot) == 2))                    \
            args.dbg[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = clock64(); \
    } while (0)

constexpr int kMaxTcStages = 4;
constexpr int kMaxBN = 256;
constexpr int kWBytes = kBM * kBK * 4;          // one 128 x 32 fp32 weight tile: 16 KB
constexpr int kTcConvThreads = 256;             // warps 4..11
constexpr int kTcThreadsV2 = 128 + kTcConvThreads;

// ---- cluster / DSMEM helpers -----------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
    float4 v;
    // not volatile: ordering against the writers comes from the cluster barrier around the reduce
    asm("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "r"(addr));
    return v;
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// source row of activation n (nullptr = row absent) and where its 128 outputs of this CTA go
struct RowIO {
    const float* src;
    float* dst;
};
constexpr int kGranM = kPage;   // positions per granule (STEP mode)

// rows the launch really has (device-side counts where the mode has them)
__device__ __forceinline__ int tc_n_valid(const TcArgs& args) {
    int n = args.n_rows;
    if (args.mode == TC_PREFILL) {
        n = min(n, *args.n_tiles * kTileM);
    } else if (args.mode == TC_STEP) {
        const int pad = (args.counts[0] + 15) & ~15;
        n = min(n, pad + (args.use_gran ? args.counts[1] * kGranM : 0));
    } else if (args.mode == TC_LOGITS && args.counts != nullptr) {
        n = min(n, args.counts[0]);
    }
    return n;
}

// Persistent launches: the work of a launch is the list of (activation tile, feature tile) pairs that have
// something to compute, feature tile fastest (the CTAs working at the same time share the activation rows
// through L2).  Activation tiles that hold active rows (STEP: the first ceil(pad / bn) tiles) need every
// feature tile; tiles of prefill positions need K and V features only (no q for prefill positions), so a
// launch that is mostly prefill keeps every SM busy instead of idling the third that owns q features.
struct TcItems {
    int a_tiles;   // activation tiles that need all feature tiles
    int kv_tiles;  // feature tiles a pure prefill tile needs
    int n_items;
    int p_tiles;   // pair kernel: pure prefill tiles (n_tiles - a_tiles)
};
__device__ __forceinline__ TcItems tc_items(const TcArgs& args, int n_valid) {
    TcItems t;
    const int n_tiles = (n_valid + args.bn - 1) / args.bn;
    if (args.mode == TC_STEP) {
        const int pad = (args.counts[0] + 15) & ~15;
        t.a_tiles = min(n_tiles, (pad + args.bn - 1) / args.bn);
        t.kv_tiles = 2 * (args.d / kBM);
    } else {
        t.a_tiles = n_tiles;
        t.kv_tiles = args.m_tiles;
    }
    t.n_items = t.a_tiles * args.m_tiles + (n_tiles - t.a_tiles) * t.kv_tiles;
    return t;
}
__device__ __forceinline__ void tc_item(const TcArgs& args, const TcItems& t, int item, int* mt, int* nt) {
    const int full = t.a_tiles * args.m_tiles;
    if (item < full) {
        *nt = item / args.m_tiles;
        *mt = item % args.m_tiles;
    } else {
        const int u = item - full;
        *nt = t.a_tiles + u / t.kv_tiles;
        const int k = u % t.kv_tiles, per = args.d / kBM;
        *mt = (k < per) ? k : k + per;   // K features, then V features ([Wk^T; Wq^T; Wv^T]: skip the q block)
    }
}

__device__ __forceinline__ RowIO row_io(const TcArgs& args, int n, int n_valid, int mat, int f0) {
    RowIO io{nullptr, nullptr};
    if (n >= n_valid) return io;
    if (args.mode == TC_LOGITS) {
        const int r = args.act ? args.act[n] : n;
        io.src = args.dense_src + (size_t)r * args.K;
        io.dst = args.score + (size_t)r * args.V + f0;
        return io;
    }
    int r, j;
    bool latest = true;
    if (args.mode == TC_LATEST) {
        r = n;
        j = args.lengths[r] - 1;
        if (j < 0) return io;
    } else if (args.mode == TC_STEP) {
        const int n_act = args.counts[0];
        const int pad = (n_act + 15) & ~15;
        if (n < pad) {
            if (n >= n_act) return io;
            r = args.act[n];
            j = args.lengths[r] - 1;
            if (j < 0) return io;
        } else {
            // earlier positions of a new row; its position L-1 is the row's "latest" entry above
            const TileDesc t = args.gran[(n - pad) / kGranM];
            r = t.row;
            j = t.j0 + ((n - pad) % kGranM);
            if (j >= (args.gran_bound ? args.gran_bound[r] : args.lengths[r]) - 1) return io;
            latest = false;
            if (mat == 1) return io;   // no q for prefill positions
        }
    } else {
        const TileDesc t = args.tiles[n / kTileM];
        r = t.row;
        j = t.j0 + (n % kTileM);
        if (j >= args.lengths[r]) return io;
    }
    (void)latest;
    float* page = args.page_table[(size_t)r * args.W + j / kPage];
    io.src = page_row_ptr(page, j, args.d, 0, args.kv_bf16);
    if (mat == 1) {
        io.dst = args.q_out + (size_t)r * args.d + f0;
    } else {
        // K / V row of the position; in the compact format it is a bf16 row: f0 elements = f0/2 floats
        float* row = page_row_ptr(page, j, args.d, mat == 0 ? 1 : 2, args.kv_bf16);
        io.dst = args.kv_bf16 ? row + f0 / 2 : row + f0;
    }
    return io;
}

// ---------------------------------------------------------------------------------------------
// Cluster split-K 3xTF32 GEMM.  grid = (feature tiles, activation-tile walkers, split), cluster =
// (1, 1, split): the `split` CTAs of a cluster share one 128 x bn output tile and each contracts a
// contiguous slice of K; partial tiles are exchanged through distributed shared memory and summed
// in a fixed order (rank 0, 1, ...), so results do not depend on scheduling.
//
//   warp 0      TMA producer: pre-split weight tiles W_hi | W_lo (cp.async.bulk.tensor, 128B swizzle)
//   warp 1      tcgen05.mma issuer (one lane), commits to the stage / accumulator mbarriers
//   warp 2      TMEM allocator
//   warps 4-11  activation path: gather fp32 rows straight from the KV pages (or the dense
//               attention result), split them into tf32 hi | lo in registers and store them in the
//               UMMA 128B-swizzled K-major layout -- no staging pass through HBM; afterwards the
//               same warps drain TMEM (tcgen05.ld) into the partial tile / the destination rows
// ---------------------------------------------------------------------------------------------
template <bool kDyn>
__global__ void __launch_bounds__(kTcThreadsV2, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   TcArgs args) {
    extern __shared__ unsigned char tc_smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TC_STAMP(0);
    const int bn = args.bn;
    const int nst = args.n_stages;
    const int split = gridDim.z;
    const int krank = blockIdx.z;   // == rank inside the (1,1,split) cluster
    const int x_bytes = bn * kBK * 4;
    const int stage_bytes = 2 * kWBytes + 2 * x_bytes;

    unsigned char* base = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* ctrl = base + (size_t)nst * stage_bytes;
    uint64_t* full_w = reinterpret_cast<uint64_t*>(ctrl);
    uint64_t* full_x = full_w + kMaxTcStages;
    uint64_t* empty_bar = full_x + kMaxTcStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxTcStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    const float** src_tab = reinterpret_cast<const float**>(ctrl + 256);
    float** dst_tab = reinterpret_cast<float**>(ctrl + 256 + kMaxBN * 8);
    float* part = reinterpret_cast<float*>(base);   // [bn][128] partial tile, aliases the stages

    // which features does this CTA produce?
    // LATEST: operand rows are [Wk^T; Wq^T; Wv^T] -> mat 0 = K, 1 = q, 2 = V
    // PREFILL: operand rows are [Wk^T; Wv^T]        -> mat 0 = K, 1 -> 2 = V
    constexpr bool dyn = kDyn;   // persistent launch with dynamic work items (args.dyn), see tc_items()
    __shared__ int s_next_item[2];
    int m0 = blockIdx.x * kBM;
    int mat = 0, f0 = m0;
    auto set_features = [&](int m_tile) {
        m0 = m_tile * kBM;
        mat = 0;
        f0 = m0;
        if (args.mode != TC_LOGITS) {
            mat = m0 / args.d;
            f0 = m0 % args.d;
            if (args.mode == TC_PREFILL && mat == 1) mat = 2;
        }
    };
    set_features((int)blockIdx.x);
    // persistent launch: this CTA's first item is its own index (known without looking at the counter, so
    // the weight tiles of that item can be requested before the dependency wait); the scheduler's counts are
    // final long before this kernel can start (see below)
    // (the item geometry lives in shared memory, not in registers: the converter warps are at the register limit)
    __shared__ TcItems s_items;
    int item = (int)blockIdx.x, nt_dyn = 0, n_items = 0;
    if (dyn) {
        const TcItems items = tc_items(args, tc_n_valid(args));
        if (tid == 0) s_items = items;   // published by the set-up barrier below
        n_items = items.n_items;
        if (item < n_items) {
            int mt;
            tc_item(args, items, item, &mt, &nt_dyn);
            set_features(mt);
        }
    }

    // The extra activation-tile walkers of a step launch (blockIdx.y >= 1) only ever see tiles past the
    // first; when all active rows fit the first tile those hold prefill positions only, which need no
    // q: their q-feature clusters leave before allocating anything.  (The scheduler's counts are final
    // long before this kernel can start, so they may be read ahead of griddepcontrol.wait.)
    if (!dyn && args.mode == TC_STEP && blockIdx.y >= 1 && mat == 1 && ((args.counts[0] + 15) & ~15) <= bn) return;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_a_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full_w[s], 1);
            mbar_init(&full_x[s], kTcConvThreads / 32);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, (uint32_t)args.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_ptr_smem;

    TC_STAMP(1);
    const int num_kb = args.K / kBK;
    const int kb_per = num_kb / split;          // host guarantees divisibility
    const int kb0 = krank * kb_per, kb1 = kb0 + kb_per;
    const int kb_per_acc = (kb_per + args.n_acc - 1) / args.n_acc;
    // The weights are static: the producer starts filling the pipeline with weight tiles right away,
    // before this kernel is allowed to look at anything its predecessor wrote.
    const int n_pre = (dyn && item >= n_items) ? 0 : min(nst, kb_per);
    const uint64_t w_policy = l2_policy_evict_last();
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < n_pre; ++i) {
            unsigned char* st = base + (size_t)i * stage_bytes;
            mbar_expect_tx(&full_w[i], 2 * kWBytes);
            tma_load_2d_hint(st, &map_a_hi, (kb0 + i) * kBK, m0, &full_w[i], w_policy);
            tma_load_2d_hint(st + kWBytes, &map_a_lo, (kb0 + i) * kBK, m0, &full_w[i], w_policy);
        }
    }
    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
    griddep_wait();
    GRIDDEP_TRIGGER_EARLY();
    trace_stamp(args.trace, args.trace_slot);

    const int n_valid = tc_n_valid(args);

    uint32_t it = 0;        // pipeline counter (runs on across tiles; identical in every role)
    uint32_t tile_iter = 0;   // tiles this CTA really processed
    for (int nt = dyn ? nt_dyn : (int)blockIdx.y;;) {
        if (dyn ? (item >= n_items) : (nt * bn >= n_valid)) break;
        // persistent launch: claim the NEXT item now; the barrier below publishes it (two slots: a thread may
        // still be reading this iteration's slot when thread 0 claims the one after)
        // (a persistent launch never skips a tile, so tile_iter counts the loop iterations)
        if (dyn && tid == 0) s_next_item[tile_iter & 1] = (int)gridDim.x + atomicAdd(&args.ctr[0], 1);
        const int n0 = nt * bn;
        const int n_eff = min(bn, ((n_valid - n0) + 15) & ~15);   // UMMA N of this tile
        // ---- row tables of the tile ----
        bool live = false;
        if (tid < bn) {
            const RowIO io = row_io(args, n0 + tid, n_valid, mat, f0);
            src_tab[tid] = io.src;
            dst_tab[tid] = io.dst;
            live = io.dst != nullptr;
        }
        // A tile none of whose rows wants this CTA's features (q features over a tile of prefill
        // positions) is skipped.  The decision depends only on (feature tile, activation tile), so
        // it is the same in every CTA of the cluster.
        const bool any_live = __syncthreads_or(live);
        // static launches skip a tile nobody wants; a persistent launch only lists tiles with work (a tile of
        // prefill granules that all lie past their row's length is computed for nothing, which is rare and
        // keeps the weight tiles requested ahead of the wait valid)
        if (!any_live && !dyn) {
            nt += gridDim.y;
            continue;
        }

        if (warp == 0) {
            // ===================== TMA producer (weights) =====================
            if (lane == 0) {
                uint32_t i = it;
                for (int kb = kb0; kb < kb1; ++kb, ++i) {
                    if (tile_iter == 0 && kb - kb0 < n_pre) continue;   // issued before the wait
                    const int s = i % nst;
                    mbar_wait(&empty_bar[s], ((i / nst) & 1) ^ 1);
                    unsigned char* st = base + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full_w[s], 2 * kWBytes);
                    tma_load_2d_hint(st, &map_a_hi, kb * kBK, m0, &full_w[s], w_policy);
                    tma_load_2d_hint(st + kWBytes, &map_a_lo, kb * kBK, m0, &full_w[s], w_policy);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                const uint32_t idesc = make_idesc_tf32(kBM, n_eff);
                uint32_t i = it;
                for (int kb = kb0; kb < kb1; ++kb, ++i) {
                    const int s = i % nst;
                    const uint32_t ph = (i / nst) & 1;
                    mbar_wait(&full_w[s], ph);
                    mbar_wait(&full_x[s], ph);
                    tc_fence_after();
                    unsigned char* st = base + (size_t)s * stage_bytes;
                    const uint64_t a_hi = make_kmajor_sw128_desc(st);
                    const uint64_t a_lo = make_kmajor_sw128_desc(st + kWBytes);
                    const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * kWBytes);
                    const uint64_t b_lo = make_kmajor_sw128_desc(st + 2 * kWBytes + x_bytes);
                    const int rel = kb - kb0;
                    const uint32_t acc = tmem_acc + (uint32_t)((rel / kb_per_acc) * args.acc_stride);
                    const bool first_kb = (rel % kb_per_acc) == 0;
                    if (kb == kb0) TC_STAMP(2);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint64_t koff = (uint64_t)((k * kUmmaK * 4) >> 4);  // 32 B per k-step
                        umma_tf32(acc, a_lo + koff, b_hi + koff, idesc, (first_kb && k == 0) ? 0u : 1u);
                        umma_tf32(acc, a_hi + koff, b_lo + koff, idesc, 1);
                        umma_tf32(acc, a_hi + koff, b_hi + koff, idesc, 1);
                    }
                    umma_commit(&empty_bar[s]);   // smem stage reusable once these MMAs have read it
                }
                umma_commit(tmem_full_bar);       // accumulators of this tile complete
            }
            __syncwarp();
        } else if (warp >= 4) {
            // ===================== activation gather + tf32 split =====================
            const int c = tid - 128;            // 0..255
            const int chunk = c & 7;            // 16-byte chunk of the 128-byte k-slice
            const int rbase = c >> 3;           // rows rbase, rbase + 32, ...
            const float4* rp[kMaxBN / 32];
#pragma unroll
            for (int i = 0; i < kMaxBN / 32; ++i) {
                const int r = rbase + 32 * i;
                const float* p = (r < n_eff) ? src_tab[r] : nullptr;
                rp[i] = p ? reinterpret_cast<const float4*>(p) + chunk : nullptr;
            }
            // two register sets (even / odd k-blocks): the loads of k-block kb+2 are in flight while
            // k-block kb+1 is converted, so global latency is off the per-stage critical path.  (The pair kernel's
            // scheme -- two warp groups on alternate k-blocks, nothing in flight when a stage is published -- was
            // measured here too: with only K/split = 8 k-blocks per CTA it loses, 20.0 vs 17.9 us per launch.)
            float4 ra[kMaxBN / 32], rb[kMaxBN / 32];
            auto load_set = [&](float4 (&dst)[kMaxBN / 32], int kb) {
#pragma unroll
                for (int i = 0; i < kMaxBN / 32; ++i)
                    dst[i] = rp[i] ? ldg_stream(rp[i] + kb * (kBK / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto store_set = [&](const float4 (&src)[kMaxBN / 32], uint32_t i2) {
                const int s = i2 % nst;
                mbar_wait(&empty_bar[s], ((i2 / nst) & 1) ^ 1);
                unsigned char* xh = base + (size_t)s * stage_bytes + 2 * kWBytes;
                unsigned char* xl = xh + x_bytes;
#pragma unroll
                for (int i = 0; i < kMaxBN / 32; ++i) {
                    const int r = rbase + 32 * i;
                    if (r < n_eff) {
                        // (cvt.rna here: the integer rounding of the pair kernel costs this one, which holds eight
                        // rows per thread, its last registers -- 88 bytes of spills; measured on the configs[1] step:
                        // 75.4 instead of 72.3 us per iteration)
                        const float4 v = src[i];
                        const float4 h = make_float4(to_tf32_rna(v.x), to_tf32_rna(v.y), to_tf32_rna(v.z),
                                                     to_tf32_rna(v.w));
                        const float4 l = make_float4(to_tf32_rna(v.x - h.x), to_tf32_rna(v.y - h.y),
                                                     to_tf32_rna(v.z - h.z), to_tf32_rna(v.w - h.w));
                        const int off = r * 128 + ((chunk ^ (r & 7)) << 4);   // 128B swizzle
                        *reinterpret_cast<float4*>(xh + off) = h;
                        *reinterpret_cast<float4*>(xl + off) = l;
                    }
                }
                // every thread orders its own stores for the async proxy; one arrival per warp
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_x[s]);
            };
            load_set(ra, kb0);
            if (kb0 + 1 < kb1) load_set(rb, kb0 + 1);
            uint32_t i2 = it;
            for (int kb = kb0; kb < kb1; kb += 2, i2 += 2) {
                store_set(ra, i2);
                if (kb + 2 < kb1) load_set(ra, kb + 2);
                if (kb + 1 < kb1) {
                    store_set(rb, i2 + 1);
                    if (kb + 3 < kb1) load_set(rb, kb + 3);
                }
            }
            TC_STAMP(3);

            // ===================== epilogue: TMEM -> partial tile / destination rows =====================
            mbar_wait(tmem_full_bar, tile_iter & 1);
            tc_fence_after();
            TC_STAMP(4);
            const int q = warp & 3;                  // TMEM lane quadrant this warp may read
            const int half = (warp - 4) >> 2;        // two warps per quadrant split the columns
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            for (int c0 = half * 32; c0 < n_eff; c0 += 64) {
                float v[32];
                tmem_ld32(tmem_acc + lane_base + (uint32_t)c0, v);
                for (int a = 1; a < args.n_acc; ++a) {
                    float u[32];
                    tmem_ld32(tmem_acc + lane_base + (uint32_t)(a * args.acc_stride + c0), u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += u[j];
                }
                // transpose through shared memory: [row][feature] so that rows leave as 512-byte runs
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < n_eff) part[(size_t)(c0 + j) * kBM + q * 32 + lane] = v[j];
            }
            tc_fence_before();
        }
        // ---- split-K exchange: rank z sums rows z, z + split, ... of all partial tiles ----
        cluster_sync_all();
        TC_STAMP(5);
        {
            // rank z stores rows z, z + split, ... summed over all ranks' partial tiles (fixed order);
            // with a deferred reduce (or no split) every CTA stores all rows of its own tile
            const bool own_only = (split == 1) || args.defer;
            const int row_step = own_only ? 1 : split;
            const int row_first = own_only ? 0 : krank;
            const int n_src = own_only ? 1 : split;
            const size_t plane = args.defer ? (size_t)krank * args.defer_stride : 0;
            const int f4 = tid & 31;
            // compact page format: K and V features leave as bf16 (round to nearest even), q as fp32
            const bool to_bf16 = args.kv_bf16 && args.mode != TC_LOGITS && mat != 1;
            auto store_row = [&](float* p, const float4& v) {
                if (to_bf16) {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                    uint2 u;
                    u.x = *reinterpret_cast<const uint32_t*>(&lo);
                    u.y = *reinterpret_cast<const uint32_t*>(&hi);
                    reinterpret_cast<uint2*>(p)[f4] = u;
                } else {
                    reinterpret_cast<float4*>(p)[f4] = v;
                }
            };
            if (own_only) {
                for (int n = tid >> 5; n < n_eff; n += kTcThreadsV2 / 32) {
                    float* p = dst_tab[n];
                    if (p == nullptr) continue;
                    store_row(p + plane, *reinterpret_cast<const float4*>(part + (size_t)n * kBM + f4 * 4));
                }
            } else {
                uint32_t raddr[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) raddr[r] = dsmem_addr(part + f4 * 4, (uint32_t)(r < split ? r : 0));
                for (int n = row_first + row_step * (tid >> 5); n < n_eff; n += row_step * (kTcThreadsV2 / 32)) {
                    float* p = dst_tab[n];
                    if (p == nullptr) continue;
                    float4 t[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        if (r < n_src) t[r] = ld_dsmem_f4(raddr[r] + (uint32_t)n * (kBM * 4));
                    float4 sum = t[0];
#pragma unroll
                    for (int r = 1; r < 8; ++r)
                        if (r < n_src) { sum.x += t[r].x; sum.y += t[r].y; sum.z += t[r].z; sum.w += t[r].w; }
                    store_row(p, sum);
                }
            }
        }
        TC_STAMP(6);
        // the partial tile aliases the pipeline stages: order these generic-proxy accesses before
        // the next tile's TMA (async proxy) writes
        asm volatile("fence.proxy.async;" ::: "memory");
        cluster_sync_all();   // partial tiles and row tables may be reused
        tc_fence_after();
        it += (uint32_t)kb_per;
        ++tile_iter;
        if (dyn) {
            // (the slot of the iteration that just ended; it is next written two iterations from now)
            item = s_next_item[(tile_iter - 1) & 1];
            if (item < n_items) {
                int mt;
                tc_item(args, s_items, item, &mt, &nt);
                set_features(mt);
            }
        } else {
            nt += gridDim.y;
        }
    }
    if (dyn && tid == 0) {
        // the last CTA out re-arms both counters for the next launch
        __threadfence();
        if (atomicAdd(&args.ctr[1], 1) == (int)gridDim.x - 1) {
            args.ctr[0] = 0;
            args.ctr[1] = 0;
            __threadfence();
        }
    }
    GRIDDEP_TRIGGER_LATE();
    if (tile_iter == 0 && warp == 0 && lane == 0) {
        // nothing to do after all: the prefetched weight tiles must land before the CTA may exit
        for (int i = 0; i < n_pre; ++i) mbar_wait(&full_w[i], 0);
    }
    tc_fence_before();
    __syncthreads();
    TC_STAMP(7);
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, (uint32_t)args.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// Bulk plan on CTA PAIRS (tcgen05 cta_group::2).  Same work items as the persistent launch above, but
// an item is (activation tile of 256 rows, PAIR of feature tiles): the two CTAs of a cluster sit on
// the two SMs of a TPC and issue ONE 256 x 256 x 8 MMA per k-step (the leader's thread issues it for
// both).  Each CTA keeps its own 128 features of the weight tile (A operand, by TMA) and converts only
// HALF of the activation rows (B operand: rows [rank * n/2, (rank + 1) * n/2) of the tile); the tensor
// cores read the other half from the peer's shared memory.  Per 32-wide k-block a CTA therefore moves
// 96 KB of MMA operands + 32 KB of converter stores + 32 KB of TMA writes through its shared memory
// instead of 144 + 64 + 32 KB, which takes the main loop from shared-memory-bound (1900 cycles per
// k-block, measured 1800) to MMA-bound (1570), and halves the gather / split work per CTA.
//   full_w / full_x of the LEADER collect both CTAs (the peer's TMA completes its bytes on the
//   leader's barrier: cp.async.bulk.tensor.cta_group::2; the peer's converter warps arrive remotely);
//   empty / accumulator-complete barriers are signalled in both CTAs by multicast commits.
// ---------------------------------------------------------------------------------------------
constexpr int kPairStages = 3;
constexpr int kPairHalf = kMaxBN / 2;                           // activation rows a CTA converts
constexpr int kPairXBytes = kPairHalf * kBK * 4;                // 16 KB (each of hi, lo)
constexpr int kPairStageBytes = 2 * kWBytes + 2 * kPairXBytes;  // 64 KB

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// executed by both CTAs of the pair; the bytes complete on the LEADER's barrier (peer bit of the shared
// window address cleared, as cute::SM100_TMA_2SM_LOAD does)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                                uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {
    const uint32_t ra = dsmem_addr(bar, cta_rank);
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void st_dsmem_i32(uint32_t addr, int v) {
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// items in units of feature-tile PAIRS
__device__ __forceinline__ TcItems tc_items_pair(const TcArgs& args, int n_valid, int bn) {
    TcItems t;
    const int n_tiles = (n_valid + bn - 1) / bn;
    const int m_units = args.m_tiles / 2;
    if (args.mode == TC_STEP) {
        const int pad = (args.counts[0] + 15) & ~15;
        t.a_tiles = min(n_tiles, (pad + bn - 1) / bn);
        t.kv_tiles = args.d / kBM;          // 2 matrices x (d / 128) / 2 pairs
    } else {
        t.a_tiles = n_tiles;
        t.kv_tiles = m_units;
    }
    t.p_tiles = n_tiles - t.a_tiles;
    t.n_items = t.a_tiles * m_units + t.p_tiles * t.kv_tiles;
    return t;
}
// Items of the prefill tiles are ordered in GROUPS of kv_group feature pairs: inside a group the activation tile is
// the slow index, so the ~74 pairs of the GPU work on kv_group weight tiles at a time.  At emb_dim 4096 a weight
// tile of a pair is 8 MB (hi + lo) and the K, V operand 268 MB: walking all 32 feature pairs per activation tile
// streamed the whole operand from HBM once per 256 positions (3.8 TB over the configs[3] prefill); with groups the
// operand is read once per group pass and the activations kv_tiles / kv_group times.  Up to emb_dim 2048 the
// operand fits the L2 and kv_group = kv_tiles (the plain order).
__device__ __forceinline__ void tc_item_pair(const TcArgs& args, const TcItems& t, int item, int* mu, int* nt) {
    const int m_units = args.m_tiles / 2;
    const int full = t.a_tiles * m_units;
    if (item < full) {
        *nt = item / m_units;
        *mu = item % m_units;
    } else {
        const int u = item - full;
        const int G = (args.kv_group > 0 && args.kv_group < t.kv_tiles) ? args.kv_group : t.kv_tiles;
        const int per_group = G * t.p_tiles;
        const int g = u / per_group, r = u % per_group;
        *nt = t.a_tiles + r / G;
        const int k = g * G + r % G, per = args.d / (2 * kBM);
        *mu = (k < per) ? k : k + per;   // K pairs, then V pairs (the q block is skipped)
    }
}

__global__ void __launch_bounds__(kTcThreadsV2, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                        TcArgs args) {
    extern __shared__ unsigned char tc_smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TC_STAMP(0);
    constexpr int nst = kPairStages;
    constexpr int stage_bytes = kPairStageBytes;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_pairs = (int)gridDim.x >> 1;

    unsigned char* base = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* ctrl = base + (size_t)nst * stage_bytes;
    uint64_t* full_w = reinterpret_cast<uint64_t*>(ctrl);
    uint64_t* full_x = full_w + kMaxTcStages;
    uint64_t* empty_bar = full_x + kMaxTcStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxTcStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    int* s_next_item = reinterpret_cast<int*>(tmem_ptr_smem + 2);   // [2], same offset in both CTAs
    const float** src_tab = reinterpret_cast<const float**>(ctrl + 256);
    float** dst_tab = reinterpret_cast<float**>(ctrl + 256 + kMaxBN * 8);
    float* part = reinterpret_cast<float*>(base);   // [256][128] tile of this CTA's features, aliases the stages
    __shared__ TcItems s_items;

    int m0 = 0, mat = 0, f0 = 0;
    auto set_features = [&](int m_tile) {
        m0 = m_tile * kBM;
        mat = 0;
        f0 = m0;
        if (args.mode != TC_LOGITS) {
            mat = m0 / args.d;
            f0 = m0 % args.d;
            if (args.mode == TC_PREFILL && mat == 1) mat = 2;
        }
    };
    // Tile width of this launch: 256 rows, except that a launch with nothing to prefill (most decode steps, and the
    // logits) takes the width the host chose so that its items fill the pairs in whole rounds -- e.g. 1024 active
    // rows x 12 feature pairs are 48 items of 256 rows for 74 pairs (one round, a third of the SMs idle) but 72
    // items of 192 rows (one round of 3/4 the length).  Uniform over the grid: the scheduler's counts are final
    // before this kernel can start (see gemm_tf32x3_kernel).
    const int bn = (args.bn_decode > 0 && (args.mode != TC_STEP || !args.use_gran || args.counts[1] == 0))
                       ? args.bn_decode : kMaxBN;
    // first item of the pair = its index
    int item = (int)blockIdx.x >> 1, nt = 0, n_items = 0;
    {
        const TcItems items = tc_items_pair(args, tc_n_valid(args), bn);
        if (tid == 0) s_items = items;
        n_items = items.n_items;
        if (item < n_items) {
            int mu;
            tc_item_pair(args, items, item, &mu, &nt);
            set_features(2 * mu + (int)rank);
        }
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_a_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full_w[s], 1);                               // the leader's expect_tx arrival
            mbar_init(&full_x[s], 2 * (kTcConvThreads / 64));       // one group of four converter warps in each CTA
            mbar_init(&empty_bar[s], 1);                            // multicast commit
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, (uint32_t)args.tmem_cols);
    tc_fence_before();
    cluster_sync_all();   // the peer's barriers exist before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_ptr_smem;

    TC_STAMP(1);
    // emb_dim 4096: the two accumulators hold chains of 1024 products each, so K is walked in two passes; the second
    // adds its sums to what the first stored (same pair, same threads' order: deterministic)
    const int n_pass = max(1, args.n_pass);
    const int num_kb = args.K / kBK / n_pass;       // k-blocks per pass
    const int kb_per_acc = (num_kb + args.n_acc - 1) / args.n_acc;
    const int n_pre = (item >= n_items) ? 0 : min(nst, num_kb);
    const uint64_t w_policy = l2_policy_evict_last();
    auto load_weights = [&](int s, int kb) {
        unsigned char* st = base + (size_t)s * stage_bytes;
        if (leader) mbar_expect_tx(&full_w[s], 4 * kWBytes);   // hi + lo of both CTAs
        tma_load_2d_2sm(st, &map_a_hi, kb * kBK, m0, &full_w[s], w_policy);
        tma_load_2d_2sm(st + kWBytes, &map_a_lo, kb * kBK, m0, &full_w[s], w_policy);
    };
    if (warp == 0 && lane == 0)
        for (int i = 0; i < n_pre; ++i) load_weights(i, i);
    griddep_wait();
    GRIDDEP_TRIGGER_EARLY();
    trace_stamp(args.trace, args.trace_slot);

    const int n_valid = tc_n_valid(args);
    uint32_t it = 0, tile_iter = 0, item_iter = 0;   // k-blocks / accumulator rounds (passes) / items so far
    while (item < n_items) {
        // the leader claims the pair's NEXT item and leaves it in both CTAs; it is read after the cluster
        // barrier that ends this tile
        if (leader && tid == 0) {
            const int nxt = n_pairs + atomicAdd(&args.ctr[0], 1);
            s_next_item[item_iter & 1] = nxt;
            st_dsmem_i32(dsmem_addr(&s_next_item[item_iter & 1], 1), nxt);
        }
        const int n0 = nt * bn;
        const int n_eff = min(bn, ((n_valid - n0) + 15) & ~15);   // UMMA N of this tile
        const int n_half = n_eff >> 1;                                // rows this CTA converts
        if (tid < bn) {
            const RowIO io = row_io(args, n0 + tid, n_valid, mat, f0);
            src_tab[tid] = io.src;
            dst_tab[tid] = io.dst;
        }
        __syncthreads();

        for (int pass = 0; pass < n_pass; ++pass) {
        const int kb0 = pass * num_kb;   // first k-block of the pass
        if (warp == 0) {
            // ===================== TMA producer (this CTA's weight tile) =====================
            if (lane == 0) {
                uint32_t i = it;
                for (int kb = 0; kb < num_kb; ++kb, ++i) {
                    if (tile_iter == 0 && kb < n_pre) continue;   // issued before the wait
                    const int s = i % nst;
                    mbar_wait(&empty_bar[s], ((i / nst) & 1) ^ 1);
                    load_weights(s, kb0 + kb);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===================== MMA issuer (leader CTA only) =====================
            if (leader && lane == 0) {
                const uint32_t idesc = make_idesc_tf32(2 * kBM, n_eff);
                uint32_t i = it;
                // diagnostics (tools/ncu_jobs.py prefill_stamps): cycles this item's issuer spent waiting for
                // weights / for activations, in row 2048 + cta of the stamp buffer
                long long w_wait = 0, x_wait = 0;
                const long long t_item = args.dbg ? clock64() : 0;
                for (int kb = 0; kb < num_kb; ++kb, ++i) {
                    const int s = i % nst;
                    const uint32_t ph = (i / nst) & 1;
                    if (args.dbg) {
                        const long long t0 = clock64();
                        mbar_wait(&full_w[s], ph);
                        const long long t1 = clock64();
                        mbar_wait(&full_x[s], ph);
                        w_wait += t1 - t0;
                        x_wait += clock64() - t1;
                    } else {
                        mbar_wait(&full_w[s], ph);
                        mbar_wait(&full_x[s], ph);
                    }
                    tc_fence_after();
                    unsigned char* st = base + (size_t)s * stage_bytes;
                    const uint64_t a_hi = make_kmajor_sw128_desc(st);
                    const uint64_t a_lo = make_kmajor_sw128_desc(st + kWBytes);
                    const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * kWBytes);
                    const uint64_t b_lo = make_kmajor_sw128_desc(st + 2 * kWBytes + kPairXBytes);
                    const uint32_t acc = tmem_acc + (uint32_t)((kb / kb_per_acc) * args.acc_stride);
                    const bool first_kb = (kb % kb_per_acc) == 0;
                    if (kb == 0) TC_STAMP(2);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint64_t koff = (uint64_t)((k * kUmmaK * 4) >> 4);
                        umma_tf32_2sm(acc, a_lo + koff, b_hi + koff, idesc, (first_kb && k == 0) ? 0u : 1u);
                        umma_tf32_2sm(acc, a_hi + koff, b_lo + koff, idesc, 1);
                        umma_tf32_2sm(acc, a_hi + koff, b_hi + koff, idesc, 1);
                    }
                    umma_commit_pair(&empty_bar[s]);
                }
                umma_commit_pair(tmem_full_bar);
                if (args.dbg) {
                    long long* row = args.dbg + (size_t)(2048 + blockIdx.x) * 8;
                    row[0] = w_wait;
                    row[1] = x_wait;
                    row[2] = clock64() - t_item;   // issue loop of the item, waits included
                    row[3] = num_kb;
                    row[4] += 1;                   // items this pair has processed
                }
            }
            __syncwarp();
        } else if (warp >= 4) {
            // ===================== activation gather + tf32 split (this CTA's half of the rows) =====================
            // Two groups of four warps take the even / the odd k-blocks.  Measured with the earlier scheme (all
            // eight warps on every k-block, raw rows prefetched into registers two or four k-blocks ahead;
            // tools/ncu_jobs.py prefill_stamps): the proxy fence that publishes a stage to the tensor core also
            // waits for the thread's OUTSTANDING GLOBAL LOADS, so every k-block paid a memory round trip (700-1400
            // cycles in the fence) and the MMA issuer waited for activations 46 % of an item.  Here a thread has
            // nothing in flight when it fences (it loads, converts, stores, fences, then loads again), and a
            // group has two MMA periods (~3100 cycles) for that round trip.
            const int grp = (warp - 4) >> 2;          // 0: even k-blocks, 1: odd
            const int c = tid - 128 - grp * 128;      // 0..127 inside the group
            const int chunk = c & 7;                  // 16-byte chunk of the 128-byte k-slice
            const int rbase = c >> 3;                 // rows rbase, rbase + 16, ...
            constexpr int kRows = kPairHalf / 16;     // 8 rows per thread
            const float4* rp[kRows];
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
                const int r = rbase + 16 * i;
                const float* p = (r < n_half) ? src_tab[(int)rank * n_half + r] : nullptr;
                rp[i] = p ? reinterpret_cast<const float4*>(p) + chunk : nullptr;
            }
            long long c_wait = 0, c_store = 0, c_fence = 0, c_arrive = 0;   // diagnostics: warp 4's time per phase of an item
            const bool c_dbg = args.dbg != nullptr && warp == 4;
            uint32_t i2 = it + (uint32_t)grp;
            for (int kb = grp; kb < num_kb; kb += 2, i2 += 2) {
                float4 raw[kRows];
#pragma unroll
                for (int i = 0; i < kRows; ++i)
                    raw[i] = rp[i] ? ldg_stream(rp[i] + (kb0 + kb) * (kBK / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const int s = i2 % nst;
                const long long t0 = c_dbg ? clock64() : 0;
                if (lane == 0) mbar_wait(&empty_bar[s], ((i2 / nst) & 1) ^ 1);   // one poller per warp
                __syncwarp();
                const long long t1 = c_dbg ? clock64() : 0;
                unsigned char* xh = base + (size_t)s * stage_bytes + 2 * kWBytes;
                unsigned char* xl = xh + kPairXBytes;
#pragma unroll
                for (int i = 0; i < kRows; ++i) {
                    const int r = rbase + 16 * i;
                    if (r < n_half) {
                        float4 h, l;
                        split_tf32(raw[i], h, l);
                        const int off = r * 128 + ((chunk ^ (r & 7)) << 4);   // 128B swizzle
                        *reinterpret_cast<float4*>(xh + off) = h;
                        *reinterpret_cast<float4*>(xl + off) = l;
                    }
                }
                const long long t2 = c_dbg ? clock64() : 0;
                fence_proxy_async_smem();
                const long long t2b = c_dbg ? clock64() : 0;
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive(&full_x[s]);
                    else mbar_arrive_remote(&full_x[s], 0);
                }
                if (c_dbg) {
                    const long long t3 = clock64();
                    c_wait += t1 - t0;
                    c_store += t2 - t1;
                    c_fence += t2b - t2;
                    c_arrive += t3 - t2b;
                }
            }
            TC_STAMP(3);
            if (c_dbg && lane == 0) {
                long long* row = args.dbg + (size_t)(2304 + blockIdx.x) * 8;
                row[0] = c_wait;
                row[1] = c_store;
                row[2] = c_fence;
                row[3] = c_arrive;
            }
            // ===================== epilogue: this CTA's 128 features x all rows of the tile =====================
            mbar_wait(tmem_full_bar, tile_iter & 1);
            tc_fence_after();
            TC_STAMP(4);
            const int q = warp & 3;
            const int half = (warp - 4) >> 2;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            for (int c0 = half * 32; c0 < n_eff; c0 += 64) {
                float v[32];
                tmem_ld32(tmem_acc + lane_base + (uint32_t)c0, v);
                for (int a = 1; a < args.n_acc; ++a) {
                    float u[32];
                    tmem_ld32(tmem_acc + lane_base + (uint32_t)(a * args.acc_stride + c0), u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += u[j];
                }
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < n_eff) part[(size_t)(c0 + j) * kBM + q * 32 + lane] = v[j];
            }
            tc_fence_before();
        }
        cluster_sync_all();   // both CTAs' MMAs have read every stage; both accumulators are in `part`
        TC_STAMP(5);
        {
            const size_t plane = 0;
            const int f4 = tid & 31;
            const bool to_bf16 = args.kv_bf16 && args.mode != TC_LOGITS && mat != 1;
            for (int n = tid >> 5; n < n_eff; n += kTcThreadsV2 / 32) {
                float* p = dst_tab[n];
                if (p == nullptr) continue;
                float4 v = *reinterpret_cast<const float4*>(part + (size_t)n * kBM + f4 * 4);
                if (pass > 0) {   // (never with bf16 rows: the host plans passes for fp32 destinations only)
                    const float4 o = reinterpret_cast<const float4*>(p + plane)[f4];
                    v.x += o.x;
                    v.y += o.y;
                    v.z += o.z;
                    v.w += o.w;
                }
                if (to_bf16) {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                    uint2 u;
                    u.x = *reinterpret_cast<const uint32_t*>(&lo);
                    u.y = *reinterpret_cast<const uint32_t*>(&hi);
                    reinterpret_cast<uint2*>(p + plane)[f4] = u;
                } else {
                    reinterpret_cast<float4*>(p + plane)[f4] = v;
                }
            }
        }
        TC_STAMP(6);
        asm volatile("fence.proxy.async;" ::: "memory");
        cluster_sync_all();   // `part` and the row tables may be reused; the next item id has landed
        tc_fence_after();
        it += (uint32_t)num_kb;
        ++tile_iter;
        }   // pass
        ++item_iter;
        item = s_next_item[(item_iter - 1) & 1];
        if (item < n_items) {
            int mu;
            tc_item_pair(args, s_items, item, &mu, &nt);
            set_features(2 * mu + (int)rank);
        }
    }
    if (leader && tid == 0) {
        __threadfence();
        if (atomicAdd(&args.ctr[1], 1) == n_pairs - 1) {   // last pair out re-arms the counters
            args.ctr[0] = 0;
            args.ctr[1] = 0;
            __threadfence();
        }
    }
    GRIDDEP_TRIGGER_LATE();
    tc_fence_before();
    cluster_sync_all();   // neither CTA leaves while the other may still signal it
    TC_STAMP(7);
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_acc, (uint32_t)args.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
// Wt_hi/lo[(mat*d + f)][k] = split(W_mat[k][f]) : transpose to K-major + tf32 split (one time)
__global__ void prepack_weights_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                       const float* __restrict__ w2, float* __restrict__ hi,
                                       float* __restrict__ lo, int d) {
    __shared__ float tile[32][33];
    const float* w = blockIdx.z == 0 ? w0 : (blockIdx.z == 1 ? w1 : w2);
    const int k0 = blockIdx.y * 32, f0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        tile[i][threadIdx.x] = w[(size_t)(k0 + i) * d + f0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const float x = tile[threadIdx.x][i];   // (k = k0 + tx, f = f0 + i)
        const float h = to_tf32_rna(x);
        const size_t o = ((size_t)blockIdx.z * d + f0 + i) * d + k0 + threadIdx.x;
        hi[o] = h;
        lo[o] = to_tf32_rna(x - h);
    }
}

// plain split of a K-major matrix (emb_table for the logits)
__global__ void split_matrix_kernel(const float4* __restrict__ x, float4* __restrict__ hi,
                                    float4* __restrict__ lo, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        float4 h = make_float4(to_tf32_rna(v.x), to_tf32_rna(v.y), to_tf32_rna(v.z), to_tf32_rna(v.w));
        hi[i] = h;
        lo[i] = make_float4(to_tf32_rna(v.x - h.x), to_tf32_rna(v.y - h.y), to_tf32_rna(v.z - h.z),
                            to_tf32_rna(v.w - h.w));
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps, operand caches, launches
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [rows][K] fp32 row-major, box = 32 floats x box_rows rows, 128B swizzle, OOB rows read as zero
int make_map(CUtensorMap* map, const float* ptr, size_t rows, int K, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable");
        return MLI_ERR_UNSUPPORTED;
    }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d)", (int)r);
        set_error(buf);
        return MLI_ERR_CUDA;
    }
    return 0;
}

// split operands of a weight set (either [Wk^T;Wq^T;Wv^T] or emb_table), with their tensor maps
struct OperandEntry {
    const float *k0, *k1, *k2;  // identifying pointers
    int rows, K;
    bool registered;
    int refs;       // registrations alive (engines / layers sharing one weight set on this context)
    float *hi, *lo;
    CUtensorMap map_hi, map_lo;
};

struct TcState {
    std::vector<OperandEntry> ops;
};

std::mutex g_tc_mu;
std::vector<std::pair<mli_ctx*, TcState*>> g_states;

TcState* state_of(mli_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    for (auto& p : g_states)
        if (p.first == ctx) return p.second;
    g_states.push_back({ctx, new TcState()});
    return g_states.back().second;
}

// finds or builds the split copy.  Registered sets are built once; anything else is rebuilt on
// every call (the pointers may have been rewritten by the caller), re-using one scratch entry.
int get_operand(mli_ctx* ctx, const float* p0, const float* p1, const float* p2, int rows, int K,
                OperandEntry** out) {
    TcState* st = state_of(ctx);
    OperandEntry* scratch = nullptr;
    for (auto& e : st->ops) {
        if (e.registered && e.k0 == p0 && e.k1 == p1 && e.k2 == p2 && e.rows == rows && e.K == K) {
            *out = &e;
            return 0;
        }
        if (!e.registered && e.rows == rows && e.K == K && (p1 != nullptr) == (e.k1 != nullptr) &&
            (p2 != nullptr) == (e.k2 != nullptr))
            scratch = &e;
    }
    if (!scratch) {
        if (ctx->ws_frozen) {
            set_error("tcgen05: unregistered weights while a captured graph is alive");
            return MLI_ERR_STATE;
        }
        OperandEntry e{};
        e.rows = rows;
        e.K = K;
        e.registered = false;
        MLI_CUDA(cudaMalloc(reinterpret_cast<void**>(&e.hi), sizeof(float) * (size_t)rows * K));
        MLI_CUDA(cudaMalloc(reinterpret_cast<void**>(&e.lo), sizeof(float) * (size_t)rows * K));
        int rc;
        if ((rc = make_map(&e.map_hi, e.hi, rows, K, kBM))) return rc;
        if ((rc = make_map(&e.map_lo, e.lo, rows, K, kBM))) return rc;
        st->ops.push_back(e);
        scratch = &st->ops.back();
    }
    scratch->k0 = p0;
    scratch->k1 = p1;
    scratch->k2 = p2;
    if (p1 != nullptr) {
        const int d = K;
        dim3 grid(d / 32, d / 32, p2 != nullptr ? 3 : 2), block(32, 8);
        prepack_weights_kernel<<<grid, block, 0, ctx->stream>>>(p0, p1, p2, scratch->hi, scratch->lo, d);
    } else {
        const size_t n4 = (size_t)rows * K / 4;
        split_matrix_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(
            reinterpret_cast<const float4*>(p0), reinterpret_cast<float4*>(scratch->hi),
            reinterpret_cast<float4*>(scratch->lo), n4);
    }
    MLI_LAUNCH_CHECK();
    *out = scratch;
    return 0;
}

bool shapes_ok(int K, int feat_per_mat) { return K % 128 == 0 && feat_per_mat % kBM == 0 && K >= 128; }

size_t tc_smem_bytes(int n_stages, int bn) {
    return (size_t)n_stages * (2 * kWBytes + 2 * (size_t)bn * kBK * 4) + 1024 /*align*/ + 256 /*barriers*/ +
           2 * kMaxBN * 8 /*row tables*/;
}

// Tile / split selection.  Per-CTA operand traffic is what bounds these GEMMs at decode sizes
// (one SM ingests ~60-100 GB/s), so the activation tile is made as wide as the UMMA allows
// (weights are then read once per 256 rows) and K is split across a cluster until the launch
// covers the GPU in a single wave.
int run_gemm(mli_ctx* ctx, const OperandEntry* w, TcArgs args, int m_tiles, int expect_rows,
             int* split_out = nullptr) {
    const long long rows = args.n_rows;
    if (rows <= 0) return 0;
    // expect_rows: rows the launch will typically see (prefill: device-side count, usually small)
    long long plan_rows = expect_rows > 0 ? std::min<long long>(rows, expect_rows) : rows;
    // deferred-reduce launches (logits) are all fixed cost around a tiny main loop: narrower tiles
    // halve the TMEM -> smem -> global epilogue per CTA
    // per CTA and the number of partial planes the decoder sums (measured: decoder 5.9 -> 4.9 us)
    const int max_bn = args.defer ? 128 : kMaxBN;
    const int n_tiles_plan = (int)((plan_rows + max_bn - 1) / max_bn);
    long long per_tile = (plan_rows + n_tiles_plan - 1) / n_tiles_plan;
    int bn = (int)((per_tile + 15) / 16 * 16);
    if (bn > max_bn) bn = max_bn;
    if (rows > plan_rows) bn = max_bn;            // the device count may exceed the plan: full-width tiles
    const int n_tiles_all = (int)((rows + bn - 1) / bn);
    const int num_kb = args.K / kBK;
    int split = 1;
    // a cluster holds at most 8 CTAs; deferred-reduce launches have independent ranks
    for (int s = args.defer ? kMaxLogitSplit : 8; s >= 2; s >>= 1) {
        if (num_kb % s == 0 && num_kb / s >= 2 && (long long)m_tiles * n_tiles_plan * s <= ctx->num_sms) {
            split = s;
            break;
        }
    }
    // accuracy before occupancy: the TMEM holds 512 / bn accumulators, and a single fp32 chain should
    // not exceed 1024 products (truncating accumulation: 2048-long chains on all-positive data drift
    // 1.4e-5).  With wide tiles and K >= 2048 that takes a K split even when one wave would not need it.
    // The engine's merged step launch on CTA pairs keeps split == 1 and walks K in passes instead (second pass adds to
    // what the first stored): emb_dim 4096 then runs the same persistent bulk kernel as 1024 / 2048 (the cluster
    // split-K plan measured 0.63 of the tf32 peak on the configs[3] prefill against 0.85 for the bulk kernel)
    int n_pass = 1;
    {
        const int acc_max = std::max(1, 512 / ((bn + 31) / 32 * 32));
        const bool pass_ok = split == 1 && args.mode == TC_STEP && bn == kMaxBN && !ctx->kv_bf16 && m_tiles % 2 == 0 &&
                             (args.d / kBM) % 2 == 0 && !getenv("MLI_TC_STATIC_TILES") && !getenv("MLI_TC_NO_PAIR") &&
                             !getenv("MLI_TC_NO_KPASS");
        if (pass_ok && args.K / acc_max > 1024) {
            for (int p = 2; p <= 8; p *= 2)
                if (num_kb % p == 0 && args.K / (p * acc_max) <= 1024) {
                    n_pass = p;
                    break;
                }
        }
        while (n_pass == 1 && args.K / (split * acc_max) > 1024 && split < (args.defer ? kMaxLogitSplit : 8) &&
               num_kb % (2 * split) == 0)
            split *= 2;
    }
    args.n_pass = n_pass;
    args.kv_group = 0;
    if (args.mode == TC_STEP) {
        // weight tile of a pair: 256 features x K x (hi + lo); keep a group within ~32 MB of the L2 (configs[3] prefill: groups of 2 / 4 / 8 / 16 / 32 pairs = 1.29 / 1.27 / 1.29 / 1.34 / 1.53 s)
        const long long tile_bytes = 256LL * args.K * 8;
        const int kv_tiles = args.d / kBM;
        int g = (int)std::max<long long>(1, (32LL << 20) / tile_bytes);
        if (const char* e = getenv("MLI_TC_KV_GROUP")) g = atoi(e);
        while (g > 1 && kv_tiles % g) --g;
        if (g >= 1 && g < kv_tiles) args.kv_group = g;
    }
    // a deferred-reduce launch (logits) that ends up without a K split is a plain bulk GEMM over thousands of
    // rows: full-width tiles, persistent launch (measured at 8192 rows: 190 us as 512 CTAs of 128 x 128 tiles)
    // (only while two accumulators of 256 columns keep the chains at 1024: emb_dim <= 2048)
    if (args.defer && split == 1 && plan_rows >= 4 * kMaxBN && args.K <= 2048) bn = kMaxBN;
    int ny = n_tiles_all;
    const int cap = std::max(1, ctx->num_sms / (m_tiles * split));
    if (ny > cap) ny = std::max(cap, std::min(n_tiles_plan, n_tiles_all));

    // engine step: when an admission burst pushes the row count past one tile, a second set of clusters
    // takes the odd tiles at the same time instead of the first set walking all of them (measured:
    // 89 -> 62 us on steps with more than 512 rows, 75.6 -> 74.1 us per step on average); without an
    // overflow those clusters find nothing and leave
    if (args.mode == TC_STEP && args.use_gran) ny = std::max(ny, std::min(2, n_tiles_all));
    args.bn = bn;
    args.acc_stride = (bn + 31) / 32 * 32;
    const int k_per_cta = args.K / split / n_pass;
    int n_acc = (k_per_cta + 511) / 512;           // fp32 chains of at most 512 (see kChunks note above)
    if (n_acc > 512 / args.acc_stride) n_acc = 512 / args.acc_stride;
    if (n_acc < 1) n_acc = 1;
    if (n_acc > num_kb / split / n_pass) n_acc = num_kb / split / n_pass;
    args.n_acc = n_acc;
    int cols = 32;
    while (cols < n_acc * args.acc_stride) cols <<= 1;
    args.tmem_cols = cols;
    int nst = (int)((200 * 1024) / (2 * kWBytes + 2 * (size_t)bn * kBK * 4));
    if (nst > kMaxTcStages) nst = kMaxTcStages;
    if (nst < 2) nst = 2;
    args.n_stages = nst;
    args.kv_bf16 = ctx->kv_bf16;
    args.dbg = reinterpret_cast<long long*>(ctx->tc_dbg);
    args.trace = ctx->trace;
    args.trace_slot = (args.mode == TC_LOGITS) ? 4 : 2;
    const size_t smem = tc_smem_bytes(nst, bn);

    { int rc0 = ensure_dyn_smem(ctx, gemm_tf32x3_kernel<false>, tc_smem_bytes(2, kMaxBN)); if (rc0) return rc0; }
    { int rc0 = ensure_dyn_smem(ctx, gemm_tf32x3_kernel<true>, tc_smem_bytes(2, kMaxBN)); if (rc0) return rc0; }
    // Bulk plans (no K split, no deferred reduce; prefill-sized launches and decode steps of thousands of rows)
    // run PERSISTENT: one CTA per SM, (feature tile, activation tile) items from an atomic counter.  Measured on
    // the static grid (ncu, profiles/r2_gemm_prefill_ncu.md): the CTAs that own q features have nothing to do on
    // prefill tiles, so a third of the SMs idled (SM active 65 % of elapsed) unless the grid was many waves deep.
    args.m_tiles = m_tiles;
    args.dyn = 0;
    args.ctr = nullptr;
    if (split == 1 && (args.mode == TC_STEP || args.mode == TC_PREFILL || args.mode == TC_LOGITS) && bn == kMaxBN &&
        !getenv("MLI_TC_STATIC_TILES")) {
        void* p = nullptr;
        int rc0 = ws_get_zeroed(ctx, WS_TC_CTR, 64, &p);
        if (rc0) return rc0;
        args.ctr = reinterpret_cast<int*>(p) + 2 * (args.mode == TC_STEP ? 0 : (args.mode == TC_PREFILL ? 1 : 2));
        args.dyn = 1;
    }
    // ... and on CTA pairs (cta_group::2) when the feature tiles pair up: emb_dim a multiple of 256
    args.bn_decode = 0;
    const bool pair = args.dyn && m_tiles % 2 == 0 && (args.mode == TC_LOGITS || (args.d / kBM) % 2 == 0) &&
                      args.n_acc * args.acc_stride <= 512 && !getenv("MLI_TC_NO_PAIR");
    if (n_pass > 1 && !pair) {
        set_error("tcgen05 GEMM plan: K passes need the CTA-pair kernel");
        return MLI_ERR_STATE;
    }
    if (pair) {
        int rc0 = ensure_dyn_smem(ctx, gemm_tf32x3_pair_kernel, tc_smem_bytes(kPairStages, kPairHalf));
        if (rc0) return rc0;
        // tile width of launches without prefill granules: the multiple of 32 rows in [128, 256] that minimises
        // (rounds over the pairs) x (width) for the rows the host expects (all B rows active)
        const int pairs = ctx->num_sms / 2;
        const int units = (args.mode == TC_STEP) ? 3 * (args.d / kBM) / 2 : m_tiles / 2;
        long long best = -1;
        for (int w = kMaxBN; w >= 128; w -= 32) {
            const long long tiles = (plan_rows + w - 1) / w;
            const long long rounds = (tiles * units + pairs - 1) / pairs;
            const long long cost = rounds * w;
            if (best < 0 || cost < best) {
                best = cost;
                args.bn_decode = w;
            }
        }
        if (args.mode == TC_PREFILL || getenv("MLI_TC_FIXED_TILE")) args.bn_decode = 0;   // prefill tiles stay 256 wide
    }
    {
        int* lp = ctx->tc_last_plan[args.mode & 3];
        lp[0] = pair ? 2 : args.dyn;
        lp[1] = split;
        lp[2] = bn;
        lp[3] = args.bn_decode;
        lp[4] = n_pass;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = args.dyn ? dim3((unsigned)(pair ? (ctx->num_sms & ~1) : ctx->num_sms), 1u, 1u)
                           : dim3((unsigned)m_tiles, (unsigned)ny, (unsigned)split);
    cfg.blockDim = dim3(kTcThreadsV2);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = pair ? 2 : 1;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = args.defer ? 1u : (unsigned)split;   // deferred: ranks are independent
    ++na;
    if (split_out) *split_out = split;
    if (ctx->use_pdl) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = (unsigned)na;
    if (pair) {
        cfg.dynamicSmemBytes = tc_smem_bytes(kPairStages, kPairHalf);
        MLI_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_pair_kernel, w->map_hi, w->map_lo, args));
    } else if (args.dyn)
        MLI_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<true>, w->map_hi, w->map_lo, args));
    else
        MLI_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<false>, w->map_hi, w->map_lo, args));
    count_launch();
    return 0;
}

}  // namespace

bool tcgen05_supported(mli_ctx*) { return get_encode_fn() != nullptr; }

// registered operands are split once and trusted until tc_unregister_all()
int tc_register_weights(mli_ctx* ctx, const float* wk, const float* wq, const float* wv, const float* emb,
                        int d, int V) {
    OperandEntry* e = nullptr;
    int rc;
    if (wk && wq && wv && shapes_ok(d, d)) {
        if ((rc = get_operand(ctx, wk, wq, wv, 3 * d, d, &e))) return rc;
        e->registered = true;
        e->refs += 1;
        if ((rc = get_operand(ctx, wk, wv, nullptr, 2 * d, d, &e))) return rc;
        e->registered = true;
        e->refs += 1;
    }
    if (emb && shapes_ok(d, V)) {
        if ((rc = get_operand(ctx, emb, nullptr, nullptr, V, d, &e))) return rc;
        e->registered = true;
        e->refs += 1;
    }
    return 0;
}

// drop the registered copies identified by these pointers (an engine or layer going away)
void tc_unregister_weights(mli_ctx* ctx, const float* wk, const float* emb) {
    TcState* st = state_of(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (size_t i = 0; i < st->ops.size();) {
        OperandEntry& e = st->ops[i];
        if (e.registered && ((wk && e.k0 == wk) || (emb && e.k0 == emb)) && --e.refs > 0) {
            ++i;   // another engine / layer still uses this copy (and may hold it in a captured graph)
        } else if (e.registered && ((wk && e.k0 == wk) || (emb && e.k0 == emb))) {
            cudaFree(e.hi);
            cudaFree(e.lo);
            st->ops.erase(st->ops.begin() + i);
        } else {
            ++i;
        }
    }
}

void tc_unregister_all(mli_ctx* ctx) {
    TcState* st = state_of(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto& e : st->ops) {
        cudaFree(e.hi);
        cudaFree(e.lo);
    }
    st->ops.clear();
}

int launch_qkv_latest_paged_tc(mli_ctx* ctx, float* const* page_table, const int* lengths,
                               const float* wk, const float* wq, const float* wv, float* q_output,
                               int B, int S, int d) {
    if (!shapes_ok(d, d)) return launch_qkv_latest_paged_simt(ctx, page_table, lengths, wk, wq, wv, q_output, B, S, d);
    OperandEntry* w = nullptr;
    int rc = get_operand(ctx, wk, wq, wv, 3 * d, d, &w);
    if (rc) return rc;
    TcArgs a{};
    a.mode = TC_LATEST; a.K = d; a.d = d; a.n_rows = B; a.page_table = page_table; a.lengths = lengths;
    a.q_out = q_output; a.W = S / kPage; a.B = B;
    return run_gemm(ctx, w, a, 3 * d / kBM, 0);
}

int launch_prefill_kv_paged_tc(mli_ctx* ctx, float* const* page_table, const TileDesc* tiles,
                               const int* n_tiles, int max_tiles, const int* lengths,
                               const float* wk, const float* wv, int S, int d) {
    if (!shapes_ok(d, d) || max_tiles <= 0)
        return launch_prefill_kv_paged_simt(ctx, page_table, tiles, n_tiles, max_tiles, lengths, wk, wv, S, d);
    OperandEntry* w = nullptr;
    int rc = get_operand(ctx, wk, wv, nullptr, 2 * d, d, &w);
    if (rc) return rc;
    TcArgs a{};
    a.mode = TC_PREFILL; a.K = d; a.d = d; a.n_rows = max_tiles * kTileM; a.n_tiles = n_tiles; a.tiles = tiles;
    a.page_table = page_table; a.lengths = lengths; a.W = S / kPage;
    // the tile count is only known on the device; admissions per engine step are few, so plan the
    // split for one activation tile and let the clusters walk the rest when there is more
    return run_gemm(ctx, w, a, 2 * d / kBM, kMaxBN);
}

int launch_step_qkv_tc(mli_ctx* ctx, float* const* page_table, const int* lengths, const int* act_rows,
                       const int* counts, const TileDesc* gran, int max_gran, int use_gran,
                       const float* wk, const float* wq, const float* wv, float* q_output, int B, int S,
                       int d, const int* gran_bound) {
    if (!shapes_ok(d, d)) {
        set_error("tcgen05 step projection: emb_dim must be a multiple of 128");
        return MLI_ERR_UNSUPPORTED;
    }
    OperandEntry* w = nullptr;
    int rc = get_operand(ctx, wk, wq, wv, 3 * d, d, &w);
    if (rc) return rc;
    TcArgs a{};
    a.mode = TC_STEP; a.K = d; a.d = d; a.page_table = page_table; a.lengths = lengths;
    a.q_out = q_output; a.W = S / kPage; a.B = B; a.act = act_rows; a.counts = counts; a.gran = gran;
    a.use_gran = use_gran;
    a.gran_bound = gran_bound;
    const long long bound = (long long)((B + 15) / 16 * 16) + (use_gran ? (long long)max_gran * kGranM : 0);
    a.n_rows = (int)std::min<long long>(bound, 1 << 30);
    // a step usually has at most B active rows plus a few short prompts: plan the split for that
    // and let the clusters walk further tiles when an admission wave brings more
    if (ctx->gemm_ev_start) MLI_CUDA(cudaEventRecord(ctx->gemm_ev_start, ctx->stream));
    rc = run_gemm(ctx, w, a, 3 * d / kBM, std::max(kMaxBN, (B + 15) / 16 * 16));
    if (rc) return rc;
    if (ctx->gemm_ev_stop) MLI_CUDA(cudaEventRecord(ctx->gemm_ev_stop, ctx->stream));
    return 0;
}

int launch_logits_tc(mli_ctx* ctx, const float* attn, const float* emb, float* score, int B, int V, int d,
                     int* n_split, const int* act_rows, const int* counts) {
    *n_split = 1;
    if (!shapes_ok(d, V)) return launch_logits_simt(ctx, attn, emb, score, B, V, d);
    OperandEntry* w = nullptr;
    int rc = get_operand(ctx, emb, nullptr, nullptr, V, d, &w);
    if (rc) return rc;
    TcArgs a{};
    a.mode = TC_LOGITS; a.K = d; a.d = d; a.n_rows = B; a.score = score; a.V = V; a.B = B;
    a.dense_src = attn; a.act = act_rows; a.counts = (act_rows != nullptr) ? counts : nullptr;
    a.defer = 1; a.defer_stride = (size_t)B * V;
    return run_gemm(ctx, w, a, V / kBM, 0, n_split);
}

}  // namespace mli
