// gemm_tcgen05.cu -- the dense contractions of the path on the 5th-gen tensor cores
// (MLI_OPT_GEMM_MODE = 0): latest-token QKV projection, prefill K/V and the logits GEMM.
//
// All three are  D[f, n] = sum_k W[f, k] * X[n, k]  with
//     f  = output feature (rows of the K-major, pre-transposed weight matrix [Wk^T; Wq^T; Wv^T],
//          or rows of emb_table for the logits)           -> UMMA M = 128 (TMEM lanes)
//     n  = activation row (batch row / prompt position)   -> UMMA N = 64 or 128 (TMEM columns)
//     k  = emb_dim                                        -> 32 floats (= one 128-byte swizzle row)
//          per pipeline stage, UMMA_K = 8 for tf32
// "swap-AB": the (small, ragged) batch dimension is the MMA N, so decode batches of any size fill
// the 128-lane accumulator.  Inputs are fp32; the tensor cores take tf32, so every operand is
// split  x = hi + lo  (hi = rna_tf32(x), lo = rna_tf32(x - hi)) and each k-step issues three
// MMAs, hi*hi + lo*hi + hi*lo, into ONE fp32 accumulator in TMEM (3xTF32; relative error ~5e-7,
// against 4.9e-4 for plain tf32 -- SURVEY App. C).  Weights are split/transposed once per
// registered weight set; activations by a gather-and-split pass that reads the page rows.
//
// Kernel anatomy (one CTA per 128 x BN output tile, 256 threads):
//   warp 0  one elected lane: TMA producer (cp.async.bulk.tensor.2d, 128B swizzle, 4 tiles/stage)
//   warp 1  one elected lane: tcgen05.mma.cta_group::1.kind::tf32 issuer, tcgen05.commit -> mbarriers
//   warp 2  TMEM allocator (tcgen05.alloc / dealloc)
//   warps 4-7  epilogue: tcgen05.ld 32x32b -> registers -> coalesced stores straight into the KV
//              pages / q_output / logits (the reference's save_to_page_table scatter,
//              src/kernels/paged_attention_cublas.cu:45-67, fused into the GEMM)
#include "common.cuh"
#include "kernels.h"

#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

namespace mli {

namespace {

constexpr int kBM = 128;       // features per tile (TMEM lanes)
constexpr int kBK = 32;        // fp32 per stage row = 128 B = one swizzle row
constexpr int kUmmaK = 8;      // tf32 MMA K
constexpr int kTcThreads = 256;
// The tensor core adds each MMA into the fp32 accumulator with truncation, so one long chain over
// K drifts by ~K/16 ulp (measured 2.4e-5 relative at K = 2048 on all-positive data).  The K loop is
// therefore cut into kChunks accumulators in TMEM which the epilogue adds with ordinary fp32 adds.
constexpr int kChunks = 4;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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

enum TcMode { TC_LATEST = 0, TC_PREFILL = 1, TC_LOGITS = 2 };

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
    int V, W, B;
};

template <int BN>
struct TcSmem {
    static constexpr int kStages = (BN == 128) ? 3 : 4;
    static constexpr int kABytes = kBM * kBK * 4;   // 16 KB
    static constexpr int kBBytes = BN * kBK * 4;    // 8 / 16 KB
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kTotal = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers, tmem ptr*/ +
                                  BN * 8 /*row destinations*/;
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   TcArgs args) {
    using SM = TcSmem<BN>;
    constexpr int kStages = SM::kStages;
    extern __shared__ unsigned char tc_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // rows this launch really has (device-side for prefill); uniform early exit before any barrier
    int n_valid = args.n_rows;
    if (args.mode == TC_PREFILL) n_valid = min(n_valid, *args.n_tiles * kTileM);
    if ((int)blockIdx.y * BN >= n_valid) return;

    // which features does this CTA produce?
    // LATEST: operand rows are [Wk^T; Wq^T; Wv^T] -> mat 0 = K, 1 = q, 2 = V
    // PREFILL: operand rows are [Wk^T; Wv^T]        -> mat 0 = K, 1 -> 2 = V
    const int m0 = blockIdx.x * kBM;
    int mat = 0, f0 = m0;
    if (args.mode != TC_LOGITS) {
        mat = m0 / args.d;
        f0 = m0 % args.d;
        if (args.mode == TC_PREFILL && mat == 1) mat = 2;
    }

    unsigned char* base = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* tiles_smem = base;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + kStages * SM::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);
    float** dst = reinterpret_cast<float**>(base + kStages * SM::kStageBytes + 256);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_a_lo);
        prefetch_tmap(&map_b_hi);
        prefetch_tmap(&map_b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, 128);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, kChunks * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_ptr_smem;
    const int num_kb = args.K / kBK;
    const int kb_per_chunk = (num_kb + kChunks - 1) / kChunks;

    // Every CTA owns one 128-feature slab and walks the activation tiles nt = blockIdx.y,
    // blockIdx.y + gridDim.y, ... (the grid is capped at ~2 CTAs per SM so that a launch whose
    // device-side row count turns out to be tiny does not pay for thousands of empty CTAs).
    // Pipeline counters run on across tiles; TMEM is handed back and forth with full/empty barriers.
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int nt = blockIdx.y; nt * BN < n_valid; nt += gridDim.y) {
                const int n0 = nt * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
                    unsigned char* st = tiles_smem + (size_t)s * SM::kStageBytes;
                    mbar_expect_tx(&full_bar[s], SM::kStageBytes);
                    tma_load_2d(st, &map_a_hi, kb * kBK, m0, &full_bar[s]);
                    tma_load_2d(st + SM::kABytes, &map_a_lo, kb * kBK, m0, &full_bar[s]);
                    tma_load_2d(st + 2 * SM::kABytes, &map_b_hi, kb * kBK, n0, &full_bar[s]);
                    tma_load_2d(st + 2 * SM::kABytes + SM::kBBytes, &map_b_lo, kb * kBK, n0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(kBM, BN);
            uint32_t it = 0, tile_iter = 0;
            for (int nt = blockIdx.y; nt * BN < n_valid; nt += gridDim.y, ++tile_iter) {
                if (tile_iter > 0) {   // epilogue must have drained the accumulators of the last tile
                    mbar_wait(tmem_empty_bar, (tile_iter - 1) & 1);
                    tc_fence_after();
                }
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&full_bar[s], (it / kStages) & 1);
                    tc_fence_after();
                    unsigned char* st = tiles_smem + (size_t)s * SM::kStageBytes;
                    const uint64_t a_hi = make_kmajor_sw128_desc(st);
                    const uint64_t a_lo = make_kmajor_sw128_desc(st + SM::kABytes);
                    const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * SM::kABytes);
                    const uint64_t b_lo = make_kmajor_sw128_desc(st + 2 * SM::kABytes + SM::kBBytes);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint64_t koff = (uint64_t)((k * kUmmaK * 4) >> 4);  // 32 B per k-step
                        const uint32_t acc = tmem_acc + (uint32_t)((kb / kb_per_chunk) * BN);
                        umma_tf32(acc, a_lo + koff, b_hi + koff, idesc, ((kb % kb_per_chunk) | k) != 0);
                        umma_tf32(acc, a_hi + koff, b_lo + koff, idesc, 1);
                        umma_tf32(acc, a_hi + koff, b_hi + koff, idesc, 1);
                    }
                    umma_commit(&empty_bar[s]);   // smem stage reusable once these MMAs have read it
                }
                umma_commit(tmem_full_bar);       // accumulators of this tile complete
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int et = threadIdx.x - 128;  // 0..127
        const int ew = warp - 4;           // TMEM lane quadrant of this warp
        const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
        const int n_chunks = (num_kb + kb_per_chunk - 1) / kb_per_chunk;
        uint32_t tile_iter = 0;
        for (int nt = blockIdx.y; nt * BN < n_valid; nt += gridDim.y, ++tile_iter) {
            const int n0 = nt * BN;
            named_bar_sync(1, 128);   // previous tile's stores no longer read dst[]
            // destination of activation row n0 + i (base pointer for this CTA's feature range)
            for (int i = et; i < BN; i += 128) {
                const int n = n0 + i;
                float* p = nullptr;
                if (n < n_valid) {
                    if (args.mode == TC_LOGITS) {
                        p = args.score + (size_t)n * args.V + f0;
                    } else {
                        int r, j;
                        bool ok;
                        if (args.mode == TC_LATEST) {
                            r = n;
                            const int L = args.lengths[r];
                            j = L - 1;
                            ok = L > 0;
                        } else {
                            const TileDesc t = args.tiles[n / kTileM];
                            r = t.row;
                            j = t.j0 + (n % kTileM);
                            ok = j < args.lengths[r];
                        }
                        if (ok) {
                            if (mat == 1) {
                                p = args.q_out + (size_t)r * args.d + f0;
                            } else {
                                float* page = args.page_table[(size_t)r * args.W + j / kPage];
                                p = page_row_ptr(page, j, args.d, mat == 0 ? 1 : 2) + f0;
                            }
                        }
                    }
                }
                dst[i] = p;
            }
            named_bar_sync(1, 128);
            mbar_wait(tmem_full_bar, tile_iter & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_acc + lane_base + (uint32_t)c0, v);
                for (int ch = 1; ch < n_chunks; ++ch) {
                    float u[32];
                    tmem_ld32(tmem_acc + lane_base + (uint32_t)(ch * BN + c0), u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += u[j];
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float* p = dst[c0 + j];
                    if (p != nullptr) p[ew * 32 + lane] = v[j];   // 32 lanes -> 128 contiguous bytes
                }
            }
            tc_fence_before();
            mbar_arrive(tmem_empty_bar);   // accumulators may be overwritten by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, kChunks * BN);
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

// gather the activation rows of a launch and split them: one warp per row
__global__ void __launch_bounds__(256)
gather_split_rows_kernel(TcArgs args, const float* __restrict__ dense_src, float* __restrict__ hi,
                         float* __restrict__ lo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n_valid = args.n_rows;
    if (args.mode == TC_PREFILL) n_valid = min(n_valid, *args.n_tiles * kTileM);
    const int d4 = args.K >> 2;
    for (int n = blockIdx.x * 8 + warp; n < n_valid; n += gridDim.x * 8) {
        const float* src = nullptr;
        if (args.mode == TC_LOGITS) {
            src = dense_src + (size_t)n * args.K;
        } else if (args.mode == TC_LATEST) {
            const int L = args.lengths[n];
            if (L > 0)
                src = page_row_ptr(args.page_table[(size_t)n * args.W + (L - 1) / kPage], L - 1, args.d, 0);
        } else {
            const TileDesc t = args.tiles[n / kTileM];
            const int j = t.j0 + (n % kTileM);
            if (j < args.lengths[t.row])
                src = page_row_ptr(args.page_table[(size_t)t.row * args.W + j / kPage], j, args.d, 0);
        }
        float4* h4 = reinterpret_cast<float4*>(hi + (size_t)n * args.K);
        float4* l4 = reinterpret_cast<float4*>(lo + (size_t)n * args.K);
        for (int c = lane; c < d4; c += 32) {
            float4 v = src ? reinterpret_cast<const float4*>(src)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 h = make_float4(to_tf32_rna(v.x), to_tf32_rna(v.y), to_tf32_rna(v.z), to_tf32_rna(v.w));
            h4[c] = h;
            l4[c] = make_float4(to_tf32_rna(v.x - h.x), to_tf32_rna(v.y - h.y), to_tf32_rna(v.z - h.z),
                                to_tf32_rna(v.w - h.w));
        }
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

// activation staging (hi | lo) with cached tensor maps per (pointer, rows, K, BN)
struct StageMaps {
    float* base;
    size_t rows;
    int K, bn;
    CUtensorMap hi, lo;
};
std::vector<StageMaps> g_stage_maps;

int get_staging(mli_ctx* ctx, size_t rows, int K, int bn, float** hi, float** lo, const CUtensorMap** mh,
                const CUtensorMap** ml) {
    void* p = nullptr;
    int rc = ws_get(ctx, WS_GEMM_A, sizeof(float) * 2 * rows * K, &p);
    if (rc) return rc;
    float* base = reinterpret_cast<float*>(p);
    *hi = base;
    *lo = base + rows * K;
    std::lock_guard<std::mutex> lk(g_tc_mu);
    for (auto& m : g_stage_maps)
        if (m.base == base && m.rows == rows && m.K == K && m.bn == bn) {
            *mh = &m.hi;
            *ml = &m.lo;
            return 0;
        }
    StageMaps m{};
    m.base = base;
    m.rows = rows;
    m.K = K;
    m.bn = bn;
    if ((rc = make_map(&m.hi, *hi, rows, K, bn))) return rc;
    if ((rc = make_map(&m.lo, *lo, rows, K, bn))) return rc;
    if (g_stage_maps.size() > 64) g_stage_maps.erase(g_stage_maps.begin());
    g_stage_maps.push_back(m);
    *mh = &g_stage_maps.back().hi;
    *ml = &g_stage_maps.back().lo;
    return 0;
}

bool shapes_ok(int K, int feat_per_mat) { return K % kBK == 0 && feat_per_mat % kBM == 0 && K >= kBK; }

template <int BN>
int launch_tc(mli_ctx* ctx, const OperandEntry* w, const CUtensorMap* mbh, const CUtensorMap* mbl, dim3 grid,
              const TcArgs& args) {
    auto kern = gemm_tf32x3_kernel<BN>;
    static bool configured = false;
    if (!configured) {
        MLI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::kTotal));
        configured = true;
    }
    kern<<<grid, kTcThreads, TcSmem<BN>::kTotal, ctx->stream>>>(w->map_hi, w->map_lo, *mbh, *mbl, args);
    MLI_LAUNCH_CHECK();
    return 0;
}

int run_gemm(mli_ctx* ctx, const OperandEntry* w, TcArgs args, const float* dense_src, int m_tiles) {
    const size_t rows = (size_t)args.n_rows;
    // BN = 64 when 128-wide tiles would leave most SMs idle
    const bool bn64 = (long long)m_tiles * ((rows + 127) / 128) < ctx->num_sms;
    const int bn = bn64 ? 64 : 128;
    float *hi, *lo;
    const CUtensorMap *mh, *ml;
    int rc = get_staging(ctx, rows, args.K, bn, &hi, &lo, &mh, &ml);
    if (rc) return rc;
    int ggrid = (int)((rows + 7) / 8);
    if (ggrid > ctx->num_sms * 8) ggrid = ctx->num_sms * 8;
    gather_split_rows_kernel<<<ggrid, 256, 0, ctx->stream>>>(args, dense_src, hi, lo);
    MLI_LAUNCH_CHECK();
    // cap the grid at ~2 CTAs per SM; CTAs walk the remaining activation tiles themselves
    unsigned ny = (unsigned)((rows + bn - 1) / bn);
    const unsigned cap = (unsigned)((2 * ctx->num_sms + m_tiles - 1) / m_tiles);
    if (ny > cap) ny = cap;
    dim3 grid(m_tiles, ny);
    return bn64 ? launch_tc<64>(ctx, w, mh, ml, grid, args) : launch_tc<128>(ctx, w, mh, ml, grid, args);
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
        if ((rc = get_operand(ctx, wk, wv, nullptr, 2 * d, d, &e))) return rc;
        e->registered = true;
    }
    if (emb && shapes_ok(d, V)) {
        if ((rc = get_operand(ctx, emb, nullptr, nullptr, V, d, &e))) return rc;
        e->registered = true;
    }
    return 0;
}

// drop the registered copies identified by these pointers (an engine or layer going away)
void tc_unregister_weights(mli_ctx* ctx, const float* wk, const float* emb) {
    TcState* st = state_of(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (size_t i = 0; i < st->ops.size();) {
        OperandEntry& e = st->ops[i];
        if (e.registered && ((wk && e.k0 == wk) || (emb && e.k0 == emb))) {
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
    return run_gemm(ctx, w, a, nullptr, 3 * d / kBM);
}

int launch_prefill_kv_paged_tc(mli_ctx* ctx, float* const* page_table, const TileDesc* tiles,
                               const int* n_tiles, int max_tiles, const int* lengths,
                               const float* wk, const float* wv, int S, int d) {
    // staging holds hi|lo copies of every row the launch could cover; past 4 GiB use the SIMT path
    if (!shapes_ok(d, d) || max_tiles <= 0 || (size_t)max_tiles * kTileM * d * 8 > (size_t(4) << 30))
        return launch_prefill_kv_paged_simt(ctx, page_table, tiles, n_tiles, max_tiles, lengths, wk, wv, S, d);
    OperandEntry* w = nullptr;
    int rc = get_operand(ctx, wk, wv, nullptr, 2 * d, d, &w);
    if (rc) return rc;
    TcArgs a{};
    a.mode = TC_PREFILL; a.K = d; a.d = d; a.n_rows = max_tiles * kTileM; a.n_tiles = n_tiles; a.tiles = tiles;
    a.page_table = page_table; a.lengths = lengths; a.W = S / kPage;
    return run_gemm(ctx, w, a, nullptr, 2 * d / kBM);
}

int launch_logits_tc(mli_ctx* ctx, const float* attn, const float* emb, float* score, int B, int V, int d) {
    if (!shapes_ok(d, V)) return launch_logits_simt(ctx, attn, emb, score, B, V, d);
    OperandEntry* w = nullptr;
    int rc = get_operand(ctx, emb, nullptr, nullptr, V, d, &w);
    if (rc) return rc;
    TcArgs a{};
    a.mode = TC_LOGITS; a.K = d; a.d = d; a.n_rows = B; a.score = score; a.V = V; a.B = B;
    return run_gemm(ctx, w, a, attn, V / kBM);
}

}  // namespace mli
