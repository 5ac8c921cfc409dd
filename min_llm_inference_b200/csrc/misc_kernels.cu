// misc_kernels.cu -- encoder, decoder (argmax / length update / next embedding) and the
// device-side tile-list builder.  All memory-bound: float4 accesses, one page-pointer load per
// CTA, grids that are multiples of the SM count where the amount of work is only known on the
// device.
#include "common.cuh"
#include "kernels.h"
#include "encoder_body.cuh"

#include <cfloat>
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

namespace mli {

// ---------------------------------------------------------------------------------------------
// tile list: positions [0, L_r) of every new row, kTileM at a time.  One CTA, block scan.
// ---------------------------------------------------------------------------------------------
__global__ void build_new_row_tiles_kernel(const int* __restrict__ new_idx,
                                           const int* __restrict__ lengths, int n_new_host,
                                           const int* __restrict__ n_new_dev,
                                           TileDesc* __restrict__ tiles, int* __restrict__ n_tiles,
                                           int max_tiles) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    griddep_wait();
    griddep_launch_dependents();
    const int n_new = n_new_dev ? *n_new_dev : n_new_host;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_new; base += blockDim.x) {
        const int i = base + tid;
        int r = -1, n = 0;
        if (i < n_new) {
            r = new_idx[i];
            n = (lengths[r] + kTileM - 1) / kTileM;
        }
        int v = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < nwarps) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int first = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + v - n;
        for (int c = 0; c < n; ++c) {
            if (first + c < max_tiles) {
                tiles[first + c].row = r;
                tiles[first + c].j0 = c * kTileM;
            }
        }
        __syncthreads();
        if (tid == 0) carry_s = carry + warp_tot[nwarps - 1];
        __syncthreads();
    }
    if (tid == 0) *n_tiles = min(carry_s, max_tiles);
}

int launch_build_new_row_tiles(mli_ctx* ctx, const int* new_idx, const int* lengths, int n_new_host,
                               const int* n_new_dev, TileDesc* tiles, int* n_tiles, int max_tiles) {
    return launch_kernel(ctx, build_new_row_tiles_kernel, dim3(1), dim3(1024), 0, new_idx, lengths,
                         n_new_host, n_new_dev, tiles, n_tiles, max_tiles);
}

// ---------------------------------------------------------------------------------------------
// paged encoder (src/kernels/encoder.cu:102-147): page[r][j].inp = E[tok_j] + P[j]
// persistent CTAs over the tile list; a warp handles one position at a time.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
paged_encoder_tiles_kernel(const float* __restrict__ emb, const float* __restrict__ pos,
                           const int* __restrict__ inp, const int* __restrict__ row_req,
                           const int* __restrict__ req_tok, float* const* __restrict__ page_table,
                           const TileDesc* __restrict__ tiles, const int* __restrict__ n_tiles,
                           const int* __restrict__ lengths, int S, int d, int tile_m,
                           unsigned long long* trace, int kv_bf16) {
    griddep_wait();
    // the encoder is empty on two steps of three and short otherwise: releasing the projection GEMM at
    // once lets its prologue (TMEM allocation, the first weight tiles) run under it (measured: encoder
    // 3.5 -> 3.1 us per step; for the long kernels a late release is better, see common.cuh)
    griddep_launch_dependents();
    trace_stamp(trace, 1);
    encode_tiles_body(emb, pos, inp, row_req, req_tok, page_table, tiles, *n_tiles, lengths, S, d, tile_m,
                      kv_bf16, (int)blockIdx.x, (int)gridDim.x, (int)(threadIdx.x >> 5), (int)(blockDim.x >> 5),
                      (int)(threadIdx.x & 31));
}

int launch_paged_encoder_tiles(mli_ctx* ctx, const float* emb, const float* pos, const int* inp,
                               const int* row_req, const int* req_tok, float* const* page_table,
                               const TileDesc* tiles, const int* n_tiles, int max_tiles,
                               const int* lengths, int S, int d, int tile_m) {
    int grid = ctx->num_sms * 4;
    if (grid > max_tiles) grid = max_tiles;
    if (grid < 1) grid = 1;
    return launch_kernel(ctx, paged_encoder_tiles_kernel, dim3(grid), dim3(256), 0, emb, pos, inp, row_req,
                         req_tok, page_table, tiles, n_tiles, lengths, S, d, tile_m, ctx->trace, ctx->kv_bf16);
}

// dense encoder (src/kernels/encoder.cu:56-92); element-wise so any emb_dim works
__global__ void dense_encoder_kernel(const float* __restrict__ emb, const float* __restrict__ pos,
                                     const int* __restrict__ inp, float* __restrict__ out,
                                     const int* __restrict__ lengths,
                                     const int* __restrict__ new_idx, int S, int d) {
    const int r = new_idx[blockIdx.y];
    const int L = lengths[r];
    for (int j = blockIdx.x; j < L; j += gridDim.x) {
        const int tok = inp[(size_t)r * S + j];
        const float* e = emb + (size_t)tok * d;
        const float* p = pos + (size_t)j * d;
        float* x = out + ((size_t)r * S + j) * d;
        for (int c = threadIdx.x; c < d; c += blockDim.x) x[c] = e[c] + p[c];
    }
}

int launch_dense_encoder(mli_ctx* ctx, const float* emb, const float* pos, const int* inp,
                         float* inp_embedding, const int* lengths, const int* new_idx, int S, int d,
                         int n_new) {
    if (n_new <= 0) return 0;
    dim3 grid(S < 64 ? S : 64, n_new);
    dense_encoder_kernel<<<grid, 128, 0, ctx->stream>>>(emb, pos, inp, inp_embedding, lengths,
                                                        new_idx, S, d);
    MLI_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// decoder (src/kernels/decoder.cu:128-205 paged, :25-91 dense).  One 256-thread CTA per row.
// The argmax reproduces the reference's tie-break exactly: thread t scans
// indices t, t+256, ... keeping the first strict maximum, then a shared-memory tree in which the
// lower thread wins ties.  lengths = L+1, or 0 when token == EOF or L+1 >= S; otherwise the next
// input embedding E[token] + P[L] is written for position L.
// ---------------------------------------------------------------------------------------------
template <bool PAGED>
__global__ void __launch_bounds__(256)
decoder_kernel(const float* __restrict__ score, int n_split, size_t split_stride,
               float* __restrict__ score_out, int* __restrict__ decoder_result,
               int* __restrict__ lengths, float* const* __restrict__ page_table,
               float* __restrict__ inp_embedding, const float* __restrict__ pos,
               const float* __restrict__ emb, int V, int S, int d, int n_dec, int i_dec,
               unsigned long long* trace, int kv_bf16) {
    const int r = blockIdx.x;
    const int tid = threadIdx.x;
    griddep_wait();
    griddep_launch_dependents();
    trace_stamp(trace, 5);
    const int L = lengths[r];
    if (L == 0) {
        if (tid == 0) decoder_result[(size_t)r * n_dec + i_dec] = MLI_EMPTY_ROW_TOKEN_ID;
        return;
    }
    __shared__ float mv[256];
    __shared__ int mi[256];
    const float* s = score + (size_t)r * V;
    float lm = -FLT_MAX;
    int li = -1;
    // thread t scans t, t+256, ... in ascending order (the reference's order, so the first strict
    // maximum wins); four indices are handled per round so that the loads of all partial planes
    // are in flight together, then the planes are added in rank order
    for (int i0 = tid; i0 < V; i0 += 1024) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (i0 + 256 * u < V) ? s[i0 + 256 * u] : 0.f;
        for (int z0 = 1; z0 < n_split; z0 += 8) {
            float t[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 8; ++w)
                    t[u][w] = (z0 + w < n_split && i0 + 256 * u < V)
                                  ? s[(size_t)(z0 + w) * split_stride + i0 + 256 * u] : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 8; ++w)
                    if (z0 + w < n_split) v[u] += t[u][w];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 256 * u;
            if (i < V) {
                if (score_out != nullptr) score_out[(size_t)r * V + i] = v[u];
                if (v[u] > lm) { lm = v[u]; li = i; }
            }
        }
    }
    mv[tid] = lm;
    mi[tid] = li;
    __syncthreads();
    for (int gap = 128; gap > 0; gap >>= 1) {
        if (tid < gap) {
            if (mv[tid + gap] > mv[tid]) { mv[tid] = mv[tid + gap]; mi[tid] = mi[tid + gap]; }
        }
        __syncthreads();
    }
    const int tok = mi[0];
    const bool stop = (tok == MLI_EOF_TOKEN_ID) || (L + 1 >= S);
    if (tid == 0) {
        decoder_result[(size_t)r * n_dec + i_dec] = tok;
        lengths[r] = stop ? 0 : L + 1;
    }
    if (stop) return;
    float* x;
    if (PAGED) {
        x = page_row_ptr(page_table[(size_t)r * (S / kPage) + L / kPage], L, d, 0, kv_bf16);
    } else {
        x = inp_embedding + ((size_t)r * S + L) * d;
    }
    const float* e = emb + (size_t)tok * d;
    const float* p = pos + (size_t)L * d;
    if ((d & 3) == 0) {
        const int d4 = d >> 2;
        for (int c = tid; c < d4; c += 256) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(e) + c);
            const float4 b = __ldg(reinterpret_cast<const float4*>(p) + c);
            reinterpret_cast<float4*>(x)[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        }
    } else {
        for (int c = tid; c < d; c += 256) x[c] = e[c] + p[c];
    }
}

int launch_paged_decoder(mli_ctx* ctx, const float* score, int n_split, float* score_out,
                         int* decoder_result, int* lengths,
                         float* const* page_table, const float* pos, const float* emb, int B, int V,
                         int S, int d, int n_dec, int i_dec) {
    return launch_kernel(ctx, decoder_kernel<true>, dim3(B), dim3(256), 0, score, n_split, (size_t)B * V,
                         score_out, decoder_result, lengths,
                         page_table, static_cast<float*>(nullptr), pos, emb, V, S, d, n_dec, i_dec, ctx->trace,
                         ctx->kv_bf16);
}

int launch_dense_decoder(mli_ctx* ctx, const float* score, int* decoder_result, int* lengths,
                         float* inp_embedding, const float* pos, const float* emb, int B, int V, int S,
                         int d) {
    return launch_kernel(ctx, decoder_kernel<false>, dim3(B), dim3(256), 0, score, 1, (size_t)0,
                         static_cast<float*>(nullptr), decoder_result, lengths,
                         static_cast<float* const*>(nullptr), inp_embedding, pos, emb, V, S, d, 1, 0,
                         static_cast<unsigned long long*>(nullptr), 0);
}

// ---------------------------------------------------------------------------------------------
// The reference's three UNFUSED attention stages, kept as stand-alone launches so that code (and the
// reference's own kernel tests, tests/paged_attention_kernels_test.cpp:115-169) written against
//   launch_qkt_paged_attention               src/kernels/paged_attention.cu:208-280
//   launch_softmax_in_place_with_lengths     src/kernels/self_attention_inference_optimized.cu:191-242
//   launch_softmax_v_paged_attention         src/kernels/paged_attention.cu:287-345
// still has something to call.  The product path never uses them (it runs the fused kernel); they keep
// the reference's summation ORDER -- one k-ascending FMA chain per score, one j-ascending chain per
// output column -- so scores and P.V are bit-identical to the reference's kernels.
// ---------------------------------------------------------------------------------------------
constexpr int kQktKc = 128;   // k columns staged per pass

// grid (W, B), 128 threads: the CTA stages a [16 positions][128 columns] slab of K with coalesced
// loads; thread p < 16 owns position p of the page and runs the reference's sequential chain
__global__ void __launch_bounds__(128)
qkt_unfused_kernel(const float* __restrict__ q, float* const* __restrict__ page_table,
                   const int* __restrict__ lengths, float* __restrict__ qkt, int S, int d) {
    __shared__ float ks[kPage][kQktKc + 1];
    __shared__ float qs[kQktKc];
    const int r = blockIdx.y, pg = blockIdx.x, tid = threadIdx.x;
    const int L = lengths[r];
    const int j0 = pg * kPage;
    if (j0 >= L) return;
    const int np = min(kPage, L - j0);
    const float* page = page_table[(size_t)r * (S / kPage) + pg];
    float acc = 0.f;
    for (int c0 = 0; c0 < d; c0 += kQktKc) {
        const int nc = min(kQktKc, d - c0);
        for (int i = tid; i < nc; i += 128) qs[i] = q[(size_t)r * d + c0 + i];
        for (int i = tid; i < np * kQktKc; i += 128) {
            const int p = i / kQktKc, c = i % kQktKc;
            if (c < nc) ks[p][c] = page[(size_t)p * 3 * d + d + c0 + c];
        }
        __syncthreads();
        if (tid < np)
            for (int c = 0; c < nc; ++c) acc = fmaf(qs[c], ks[tid][c], acc);
        __syncthreads();
    }
    if (tid < np) qkt[(size_t)r * S + j0 + tid] = acc / sqrtf((float)d);
}

// One warp per row, in the arithmetic order of the reference's kernel (self_attention_inference_optimized.cu:191-242)
// so that the probabilities are BIT-identical to it: lane t walks the groups of four scores t, t + 32, ... keeping a
// running (max, sum) pair -- the sum is rescaled by expf(old max - new max) once per group, then the group's terms are
// added in index order; the lanes' pairs are merged with the toolkit's own warp reduction (cg::reduce, the same
// primitive and therefore the same tree as the reference's build); p_j = expf(x_j - max) * (1.f / sum).
__global__ void __launch_bounds__(256)
softmax_lengths_unfused_kernel(float* __restrict__ qkt, const int* __restrict__ lengths, int B, int S) {
    namespace cg = cooperative_groups;
    const cg::thread_block_tile<32> warp = cg::tiled_partition<32>(cg::this_thread_block());
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= B) return;
    const int L = lengths[r];
    float* row = qkt + (size_t)r * S;
    float m = -FLT_MAX, s = 0.f;
    for (int g = lane; 4 * g < L; g += 32) {
        const int n = min(4, L - 4 * g);
        const float4 x4 = reinterpret_cast<const float4*>(row)[g];
        const float x[4] = {x4.x, x4.y, x4.z, x4.w};
        const float before = m;
        for (int j = 0; j < n; ++j) m = fmaxf(m, x[j]);
        s *= expf(before - m);
        for (int j = 0; j < n; ++j) s += expf(x[j] - m);
    }
    const float m_all = cg::reduce(warp, m, cg::greater<float>());
    s *= expf(m - m_all);
    const float inv = 1.f / cg::reduce(warp, s, cg::plus<float>());
    for (int g = lane; g < S / 4; g += 32) {
        const float4 x4 = reinterpret_cast<const float4*>(row)[g];
        const float x[4] = {x4.x, x4.y, x4.z, x4.w};
        float p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = (4 * g + j < L) ? expf(x[j] - m_all) * inv : 0.f;
        reinterpret_cast<float4*>(row)[g] = make_float4(p[0], p[1], p[2], p[3]);
    }
}

// grid (ceil(d / 256), B), 256 threads: thread = one output column, positions ascending
__global__ void __launch_bounds__(256)
softmax_v_unfused_kernel(const float* __restrict__ p, float* const* __restrict__ page_table,
                         float* __restrict__ out, const int* __restrict__ lengths, int S, int d) {
    __shared__ float ps[256];
    __shared__ const float* pages[16];
    const int r = blockIdx.y, col = blockIdx.x * 256 + threadIdx.x;
    const int L = lengths[r];
    float acc = 0.f;
    for (int j0 = 0; j0 < L; j0 += 256) {
        const int n = min(256, L - j0);
        if (threadIdx.x < n) ps[threadIdx.x] = p[(size_t)r * S + j0 + threadIdx.x];
        if (threadIdx.x < 16 && j0 + threadIdx.x * kPage < L)
            pages[threadIdx.x] = page_table[(size_t)r * (S / kPage) + j0 / kPage + threadIdx.x];
        __syncthreads();
        if (col < d)
            for (int j = 0; j < n; ++j)
                acc = fmaf(ps[j], pages[j / kPage][(size_t)(j % kPage) * 3 * d + 2 * d + col], acc);
        __syncthreads();
    }
    if (col < d) out[(size_t)r * d + col] = acc;
}

int launch_qkt_unfused(mli_ctx* ctx, const float* q, float* const* page_table, const int* lengths, float* qkt,
                       int B, int S, int d) {
    qkt_unfused_kernel<<<dim3(S / kPage, B), 128, 0, ctx->stream>>>(q, page_table, lengths, qkt, S, d);
    MLI_LAUNCH_CHECK();
    return 0;
}

int launch_softmax_lengths_unfused(mli_ctx* ctx, float* qkt, const int* lengths, int B, int S) {
    softmax_lengths_unfused_kernel<<<(B + 7) / 8, 256, 0, ctx->stream>>>(qkt, lengths, B, S);
    MLI_LAUNCH_CHECK();
    return 0;
}

int launch_softmax_v_unfused(mli_ctx* ctx, const float* p, float* const* page_table, float* out,
                             const int* lengths, int B, int S, int d) {
    softmax_v_unfused_kernel<<<dim3((d + 255) / 256, B), 256, 0, ctx->stream>>>(p, page_table, out, lengths, S, d);
    MLI_LAUNCH_CHECK();
    return 0;
}

}  // namespace mli
