// Body of the paged encoder (src/kernels/encoder.cu:102-147): page[r][j].inp = E[tok_j] + P[j] over a
// list of (row, first position) tiles.  Kept apart from its kernel (misc_kernels.cu) so that it can be
// called with any block shape: an experiment ran it inside the scheduler's launch (CTA 0 scheduling, the
// other CTAs waiting on a release flag, then encoding) to save the dependent-launch hop between the two
// -- measured slower (77.9 vs 73.0 us per engine step: the flag hand-off and the parked CTAs cost more
// than the hop) and dropped.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace mli {

// vb / n_vb: this (virtual) block and their number; w / nw: this warp and the warps per block.
// A warp handles two positions at a time (m and m + tile_m / 2): both token ids first, then all the
// embedding loads of both positions, then the stores -- the kernel is a chain of dependent (mostly
// L2-cold) loads, so what matters is how many are in flight together.
__device__ __forceinline__ void encode_tiles_body(const float* emb, const float* pos, const int* inp,
                                                  const int* row_req, const int* req_tok,
                                                  float* const* page_table, const TileDesc* tiles, int nt,
                                                  const int* lengths, int S, int d, int tile_m, int kv_bf16,
                                                  int vb, int n_vb, int w, int nw, int lane) {
    const int W = S / kPage, d4 = d >> 2;
    const int half = (tile_m + 1) / 2;
    for (int t = vb; t < nt; t += n_vb) {
        const TileDesc td = tiles[t];
        const int L = lengths[td.row];
        // engine mode: tokens come straight from the device request table (no inp[B,S] copy)
        const int* toks = row_req ? req_tok + (size_t)row_req[td.row] * S : inp + (size_t)td.row * S;
        for (int m0 = w; m0 < half; m0 += nw) {
            int jj[2], tk[2];
            float4* xx[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int m = m0 + half * u;
                jj[u] = td.j0 + m;
                const bool on = m < tile_m && jj[u] < L;
                tk[u] = on ? toks[jj[u]] : -1;
                xx[u] = on ? reinterpret_cast<float4*>(
                                 page_row_ptr(page_table[(size_t)td.row * W + jj[u] / kPage], jj[u], d, 0, kv_bf16))
                           : nullptr;
            }
            for (int c0 = lane; c0 < d4; c0 += 128) {
                float4 a[2][4], b[2][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (tk[u] < 0) continue;
                    const float4* e = reinterpret_cast<const float4*>(emb + (size_t)tk[u] * d);
                    const float4* p = reinterpret_cast<const float4*>(pos + (size_t)jj[u] * d);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int c = c0 + 32 * v;
                        if (c < d4) { a[u][v] = __ldg(e + c); b[u][v] = __ldg(p + c); }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (tk[u] < 0) continue;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int c = c0 + 32 * v;
                        if (c < d4)
                            xx[u][c] = make_float4(a[u][v].x + b[u][v].x, a[u][v].y + b[u][v].y,
                                                   a[u][v].z + b[u][v].z, a[u][v].w + b[u][v].w);
                    }
                }
            }
        }
    }
}

}  // namespace mli
