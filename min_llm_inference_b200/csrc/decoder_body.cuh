// Body of the decoder (src/kernels/decoder.cu:128-205 paged, :25-91 dense) for one batch row, executed by
// one 256-thread CTA.  Kept apart from its kernel (misc_kernels.cu) because the engine also runs it inside
// the launch that does the next step's scheduling (engine.cu: decoder_sched_kernel).
//
// The argmax reproduces the reference's tie-break exactly: thread t scans indices
// t, t+256, ... keeping the first strict maximum, then a shared-memory tree in which the lower thread
// wins ties.  lengths = L+1, or 0 when token == EOF or L+1 >= S; otherwise the next input embedding
// E[token] + P[L] is written for position L.
#pragma once
#include "common.cuh"
#include "kernels.h"

#include <cfloat>

namespace mli {

template <bool PAGED>
__device__ __forceinline__ void decoder_row(const float* __restrict__ score, int n_split, size_t split_stride,
                                            float* __restrict__ score_out, int* __restrict__ decoder_result,
                                            int* __restrict__ lengths, float* const* __restrict__ page_table,
                                            float* __restrict__ inp_embedding, const float* __restrict__ pos,
                                            const float* __restrict__ emb, int V, int S, int d, int n_dec,
                                            int i_dec, int kv_bf16, int r) {
    const int tid = threadIdx.x;
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

}  // namespace mli
