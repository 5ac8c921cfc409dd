// gemm_simt.cu -- exact-order fp32 GEMMs on the CUDA cores (MLI_OPT_GEMM_MODE = 1).
//
// Every output element is ONE fp32 accumulator updated by fmaf(a[k], b[k], acc) for k ascending
// from 0 -- the same chain the reference's naive kernels execute
// (src/kernels/paged_attention.cu:64-66, :167-171; src/kernels/gemm.cu:42-44;
//  src/kernels/self_attention_inference_optimized.cu:63-79, :128-134), so K, V, q and the logits
// are bit-identical to the reference's naive CUDA path.  The tiling (64x64x16, 4x4 register
// tiles, double-buffered shared memory, persistent CTAs over a device-side tile list) only
// changes which thread owns which output, never the order inside an output.
//
// One kernel template serves five call sites through a "row map" functor that says where row m
// of an M-tile reads its activation vector and where each of up to three weight matrices'
// results go (page sub-rows, q_output, logits, or the dense transposed K cache).
#include "common.cuh"
#include "kernels.h"

namespace mli {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NTHREADS = 256;
constexpr int APAD = 4, BPAD = 4;

struct RowRec {
    const float* a;  // activation row (K floats) or nullptr when the row is inactive
    float* o[3];     // output base (n = 0) per weight matrix
};

// ---- row maps ------------------------------------------------------------------------------
struct LatestPagedMap {  // a10: x = page[L-1].inp -> k,v into the page, q -> q_output
    float* const* pt;
    const int* lengths;
    float* q_out;
    int B, W, d;
    __device__ int num_tiles() const { return (B + BM - 1) / BM; }
    __device__ void row(int tile, int m, RowRec& rr) const {
        rr.a = nullptr;
        const int r = tile * BM + m;
        if (r >= B) return;
        const int L = lengths[r];
        if (L <= 0) return;
        const int j = L - 1;
        float* base = page_row_ptr(pt[(size_t)r * W + j / kPage], j, d, 0);
        rr.a = base;
        rr.o[0] = base + d;                   // Wk -> K sub-row
        rr.o[1] = q_out + (size_t)r * d;      // Wq -> q_output
        rr.o[2] = base + 2 * (size_t)d;       // Wv -> V sub-row
    }
};

struct PrefillPagedMap {  // a9: new rows, j < L
    float* const* pt;
    const TileDesc* tiles;
    const int* n_tiles;
    const int* lengths;
    int W, d;
    __device__ int num_tiles() const { return *n_tiles; }
    __device__ void row(int tile, int m, RowRec& rr) const {
        rr.a = nullptr;
        const TileDesc t = tiles[tile];
        const int j = t.j0 + m;
        if (j >= lengths[t.row]) return;
        float* base = page_row_ptr(pt[(size_t)t.row * W + j / kPage], j, d, 0);
        rr.a = base;
        rr.o[0] = base + d;
        rr.o[1] = base + 2 * (size_t)d;
        rr.o[2] = nullptr;
    }
};

struct LogitsMap {  // a14: logits[r, :] = attn[r, :] . E^T
    const float* attn;
    float* score;
    int B, V, d;
    __device__ int num_tiles() const { return (B + BM - 1) / BM; }
    __device__ void row(int tile, int m, RowRec& rr) const {
        rr.a = nullptr;
        const int r = tile * BM + m;
        if (r >= B) return;
        rr.a = attn + (size_t)r * d;
        rr.o[0] = score + (size_t)r * V;
        rr.o[1] = rr.o[2] = nullptr;
    }
};

struct LatestDenseMap {  // a16: kt_cache is transposed [B, dn, S]
    const float* x;
    const int* lengths;
    float *kt, *v, *q_out;
    int B, S, di, dn;
    __device__ int num_tiles() const { return (B + BM - 1) / BM; }
    __device__ void row(int tile, int m, RowRec& rr) const {
        rr.a = nullptr;
        const int r = tile * BM + m;
        if (r >= B) return;
        const int L = lengths[r];
        if (L <= 0) return;
        const int j = L - 1;
        rr.a = x + ((size_t)r * S + j) * di;
        rr.o[0] = kt + (size_t)r * dn * S + j;        // element stride S
        rr.o[1] = q_out + (size_t)r * dn;
        rr.o[2] = v + ((size_t)r * S + j) * dn;
    }
};

struct PrefillDenseMap {
    const float* x;
    const TileDesc* tiles;
    const int* n_tiles;
    const int* lengths;
    float *kt, *v;
    int S, di, dn;
    __device__ int num_tiles() const { return *n_tiles; }
    __device__ void row(int tile, int m, RowRec& rr) const {
        rr.a = nullptr;
        const TileDesc t = tiles[tile];
        const int j = t.j0 + m;
        if (j >= lengths[t.row]) return;
        rr.a = x + ((size_t)t.row * S + j) * di;
        rr.o[0] = kt + (size_t)t.row * dn * S + j;
        rr.o[1] = v + ((size_t)t.row * S + j) * dn;
        rr.o[2] = nullptr;
    }
};

struct GemmParams {
    const float* w[3];  // weight matrices
    int n_mats;
    int K;              // contraction length
    int N;              // output columns per matrix
    int ldb;            // leading dimension of the weight matrices
    int ostride[3];     // element stride of the outputs along n (1, or S for the transposed K cache)
};

// B_NT = false: w is [K][N] row-major (x . W);  true: w is [N][K] row-major (x . E^T)
template <class Map, bool B_NT>
__global__ void __launch_bounds__(NTHREADS)
gemm_exact_kernel(Map map, GemmParams prm) {
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + BPAD];
    __shared__ RowRec rows[BM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int K = prm.K, N = prm.N;
    const int n_tiles_per_mat = (N + BN - 1) / BN;
    const int n_ntiles = prm.n_mats * n_tiles_per_mat;
    const long long total = (long long)map.num_tiles() * n_ntiles;
    const bool a_vec = (K & 3) == 0;
    const bool b_vec = B_NT ? ((K & 3) == 0) : ((prm.ldb & 3) == 0);
    const int nk = (K + BK - 1) / BK;

    // A loader: thread -> (row am, 4 consecutive k starting at 4*ak4)
    const int am = tid >> 2, ak4 = tid & 3;
    // B loader (NN): thread -> (k = tid/16, 4 consecutive n at 4*(tid%16))
    // B loader (NT): thread -> (n = tid/4, 4 consecutive k at 4*(tid%4))
    const int bk = B_NT ? 0 : (tid >> 4), bn4 = B_NT ? 0 : (tid & 15);
    const int bn = B_NT ? (tid >> 2) : 0, bk4 = B_NT ? (tid & 3) : 0;

    for (long long work = blockIdx.x; work < total; work += gridDim.x) {
        const int tile = (int)(work / n_ntiles);
        const int nt = (int)(work % n_ntiles);
        const int mat = nt / n_tiles_per_mat;
        const int n0 = (nt % n_tiles_per_mat) * BN;
        const float* __restrict__ w = prm.w[mat];

        __syncthreads();  // previous work item fully done with rows[] / smem tiles
        if (tid < BM) map.row(tile, tid, rows[tid]);
        __syncthreads();

        const float* arow = rows[am].a;
        float4 ra, rb;
        auto load_tiles = [&](int k0) {
            // ---- A ----
            const int ka = k0 + 4 * ak4;
            ra = make_float4(0.f, 0.f, 0.f, 0.f);
            if (arow != nullptr) {
                if (a_vec) {
                    if (ka < K) ra = *reinterpret_cast<const float4*>(arow + ka);
                } else {
                    if (ka + 0 < K) ra.x = arow[ka + 0];
                    if (ka + 1 < K) ra.y = arow[ka + 1];
                    if (ka + 2 < K) ra.z = arow[ka + 2];
                    if (ka + 3 < K) ra.w = arow[ka + 3];
                }
            }
            // ---- B ----
            rb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!B_NT) {
                const int k = k0 + bk, n = n0 + 4 * bn4;
                if (k < K) {
                    const float* src = w + (size_t)k * prm.ldb + n;
                    if (b_vec && n + 3 < N) {
                        rb = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        if (n + 0 < N) rb.x = __ldg(src + 0);
                        if (n + 1 < N) rb.y = __ldg(src + 1);
                        if (n + 2 < N) rb.z = __ldg(src + 2);
                        if (n + 3 < N) rb.w = __ldg(src + 3);
                    }
                }
            } else {
                const int n = n0 + bn, k = k0 + 4 * bk4;
                if (n < N) {
                    const float* src = w + (size_t)n * prm.ldb + k;
                    if (b_vec) {
                        if (k < K) rb = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        if (k + 0 < K) rb.x = __ldg(src + 0);
                        if (k + 1 < K) rb.y = __ldg(src + 1);
                        if (k + 2 < K) rb.z = __ldg(src + 2);
                        if (k + 3 < K) rb.w = __ldg(src + 3);
                    }
                }
            }
        };
        auto store_tiles = [&](int buf) {
            As[buf][4 * ak4 + 0][am] = ra.x;
            As[buf][4 * ak4 + 1][am] = ra.y;
            As[buf][4 * ak4 + 2][am] = ra.z;
            As[buf][4 * ak4 + 3][am] = ra.w;
            if (!B_NT) {
                *reinterpret_cast<float4*>(&Bs[buf][bk][4 * bn4]) = rb;
            } else {
                Bs[buf][4 * bk4 + 0][bn] = rb.x;
                Bs[buf][4 * bk4 + 1][bn] = rb.y;
                Bs[buf][4 * bk4 + 2][bn] = rb.z;
                Bs[buf][4 * bk4 + 3][bn] = rb.w;
            }
        };

        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

        load_tiles(0);
        store_tiles(0);
        __syncthreads();
        for (int kt = 0; kt < nk; ++kt) {
            const int buf = kt & 1;
            if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][4 * ty]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][4 * tx]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (kt + 1 < nk) store_tiles(buf ^ 1);
            __syncthreads();
        }

        // ---- store ----
        const int ostride = prm.ostride[mat];
        const int n = n0 + 4 * tx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const RowRec& rr = rows[4 * ty + i];
            if (rr.a == nullptr) continue;
            float* o = rr.o[mat];
            if (ostride == 1 && (N & 3) == 0 && n + 3 < N) {
                *reinterpret_cast<float4*>(o + n) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) o[(size_t)(n + j) * ostride] = acc[i][j];
            }
        }
    }
}

template <class Map, bool B_NT>
int launch_gemm(mli_ctx* ctx, const Map& map, const GemmParams& prm, long long work_upper_bound) {
    long long grid = (long long)ctx->num_sms * 4;
    if (work_upper_bound > 0 && grid > work_upper_bound) grid = work_upper_bound;
    if (grid < 1) grid = 1;
    gemm_exact_kernel<Map, B_NT><<<(unsigned)grid, NTHREADS, 0, ctx->stream>>>(map, prm);
    MLI_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int launch_prefill_kv_paged_simt(mli_ctx* ctx, float* const* page_table, const TileDesc* tiles,
                                 const int* n_tiles, int max_tiles, const int* lengths,
                                 const float* wk, const float* wv, int S, int d) {
    if (ctx->kv_bf16) {
        set_error("the compact KV format (bf16 K/V) is only implemented for the tcgen05 GEMM mode and emb_dim % 128 == 0");
        return MLI_ERR_UNSUPPORTED;
    }
    PrefillPagedMap map{page_table, tiles, n_tiles, lengths, S / kPage, d};
    GemmParams prm{{wk, wv, nullptr}, 2, d, d, d, {1, 1, 1}};
    return launch_gemm<PrefillPagedMap, false>(ctx, map, prm,
                                               (long long)max_tiles * 2 * ceil_div(d, BN));
}

int launch_qkv_latest_paged_simt(mli_ctx* ctx, float* const* page_table, const int* lengths,
                                 const float* wk, const float* wq, const float* wv, float* q_output,
                                 int B, int S, int d) {
    if (ctx->kv_bf16) {
        set_error("the compact KV format (bf16 K/V) is only implemented for the tcgen05 GEMM mode and emb_dim % 128 == 0");
        return MLI_ERR_UNSUPPORTED;
    }
    LatestPagedMap map{page_table, lengths, q_output, B, S / kPage, d};
    GemmParams prm{{wk, wq, wv}, 3, d, d, d, {1, 1, 1}};
    return launch_gemm<LatestPagedMap, false>(ctx, map, prm,
                                              (long long)ceil_div(B, BM) * 3 * ceil_div(d, BN));
}

int launch_logits_simt(mli_ctx* ctx, const float* attn, const float* emb, float* score, int B, int V,
                       int d) {
    LogitsMap map{attn, score, B, V, d};
    GemmParams prm{{emb, nullptr, nullptr}, 1, d, V, d, {1, 1, 1}};
    return launch_gemm<LogitsMap, true>(ctx, map, prm, (long long)ceil_div(B, BM) * ceil_div(V, BN));
}

int launch_prefill_kv_dense_simt(mli_ctx* ctx, const float* inp_embedding, const TileDesc* tiles,
                                 const int* n_tiles, int max_tiles, const int* lengths,
                                 const float* wk, const float* wv, float* kt_cache, float* v_cache,
                                 int S, int di, int dn) {
    PrefillDenseMap map{inp_embedding, tiles, n_tiles, lengths, kt_cache, v_cache, S, di, dn};
    GemmParams prm{{wk, wv, nullptr}, 2, di, dn, dn, {S, 1, 1}};
    return launch_gemm<PrefillDenseMap, false>(ctx, map, prm,
                                               (long long)max_tiles * 2 * ceil_div(dn, BN));
}

int launch_qkv_latest_dense_simt(mli_ctx* ctx, const float* inp_embedding, const int* lengths,
                                 const float* wk, const float* wq, const float* wv, float* kt_cache,
                                 float* v_cache, float* q_output, int B, int S, int di, int dn) {
    LatestDenseMap map{inp_embedding, lengths, kt_cache, v_cache, q_output, B, S, di, dn};
    GemmParams prm{{wk, wq, wv}, 3, di, dn, dn, {S, 1, 1}};
    return launch_gemm<LatestDenseMap, false>(ctx, map, prm,
                                              (long long)ceil_div(B, BM) * 3 * ceil_div(dn, BN));
}

}  // namespace mli
