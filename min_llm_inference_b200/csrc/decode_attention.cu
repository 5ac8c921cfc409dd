// decode_attention.cu -- fused, length-aware, split-KV paged decode attention for sm_100a.
//
// Replaces the reference's three-kernel chain
//   qkt_paged_attention            (src/kernels/paged_attention.cu:208-263)
//   softmax_in_place_with_lengths  (src/kernels/self_attention_inference_optimized.cu:191-242)
//   softmax_v_paged_attention      (src/kernels/paged_attention.cu:287-326)
// and its [B,S] score round-trips with one pass over K and V.
//
// Design (HBM-bound; SURVEY 8d: ATTN_BYTES = sum_r 8*d*L_r + 8*d + 8*ceil(L_r/16) + 4):
//   * ONE launch: persistent CTAs (a multiple of the SM count) each take slices of the flattened
//     work (in units of pipeline stages) derived inside the kernel from the device lengths;
//   * one producer warp streams K|V rows (they are adjacent inside a page: [inp|K|V] per
//     position, so one 8*d-byte bulk copy per position) into a shared-memory ring with
//     cp.async.bulk + mbarrier complete_tx (TMA, non-tensor form -- page tables hold raw pointers,
//     so a tensor map cannot follow them), marked L2::evict_first; the page pointers of the next
//     segment are fetched while the current one streams;
//   * eight consumer warps: each thread owns float4 columns of d; q.K partial dots are reduced by
//     warp shuffle + a small shared array; online softmax (running max / sum) in fp32;
//     P.V accumulates in registers; K and V are each read exactly once from HBM;
//   * rows cut by a slice boundary leave (m, l, acc[d]) partials that the CTA completing the row
//     merges; empty rows are zero-filled; a three-launch variant (prep, main, combine) also
//     materialises the reference's [B,S] probabilities for the tests.
// Scale is dot / sqrtf(d) (a division, paged_attention.cu:261) and the exponent is expf, as in
// the reference.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cfloat>

namespace mli {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kAttnThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kMaxStages = 8;

// ---------------------------------------------------------------------------------------------
// prep: list the work items.  One CTA; rows are scanned in blocks of blockDim.x.
//   row_first[r]   first item of row r (exclusive prefix of chunk counts), row_first[B] = total
//   item_row/item_chunk[i]
// ---------------------------------------------------------------------------------------------
__global__ void attn_prep_kernel(const int* __restrict__ lengths, int B, int chunk_pos,
                                 int* __restrict__ row_first, int* __restrict__ item_row,
                                 int* __restrict__ item_chunk) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < B; base += blockDim.x) {
        int r = base + tid;
        int L = (r < B) ? lengths[r] : 0;
        int n = (L + chunk_pos - 1) / chunk_pos;
        // inclusive warp scan
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
            warp_tot[lane] = w;  // inclusive totals of warps
        }
        __syncthreads();
        int carry = carry_s;
        int first = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + v - n;
        if (r < B) {
            row_first[r] = first;
            for (int c = 0; c < n; ++c) {
                item_row[first + c] = r;
                item_chunk[first + c] = c;
            }
        }
        __syncthreads();
        if (tid == 0) carry_s = carry + warp_tot[nwarps - 1];
        __syncthreads();
    }
    if (tid == 0) row_first[B] = carry_s;
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
// FUSED = true (the production path, ONE launch per decode step): every CTA scans the lengths into
// a shared-memory prefix over pipeline STAGES (G positions of one row) and takes equal, contiguous
// slices of the flattened stage space (stream-K style: the streaming time of a CTA is proportional
// to its number of stages, so slices are balanced whatever the mix of row lengths).  A slice may
// start or end inside a row; such partial rows (at most two per slice: its head and its tail) are
// merged by whichever CTA finishes the row's last segment (per-row arrival counter in global
// memory, self-resetting).  B <= kMaxFusedRows.
// FUSED = false: (row, chunk) items from attn_prep_kernel, merge by attn_combine_kernel -- used when
// the [B,S] probabilities are requested or B is too large for the shared-memory prefix.
constexpr int kMaxFusedRows = 8192;
constexpr int kMaxPend = 32;
// Slice geometry.  A launch whose fair share per CTA reaches ctx->attn_min_dyn positions (default
// 4096, MLI_OPT_ATTN_MIN_DYN) hands the last quarter of the work out dynamically.  Dynamic slices pay for
// themselves only on long launches: rows that cross the small dynamic slices are cut into a few dozen
// partial rows, and the CTAs that finish last merge them serially (measured at B=128, d=4096,
// S=2048: a 50-65 us tail on a 650 us launch, 6.5 vs 6.8 TB/s all-static; at the configs[3] shape,
// 9.4 ms per launch, the same tail is noise and the dynamic quarter is worth 3 %).
#ifndef MLI_ATTN_STATIC_PCT
#define MLI_ATTN_STATIC_PCT 75
#endif
#ifndef MLI_ATTN_DYN_PARTS
#define MLI_ATTN_DYN_PARTS 3
#endif
constexpr int kAttnCtrlInts = 32 + kMaxStages + 3 * kMaxPend;

struct AttnSeg {
    int r;        // batch row
    int p0, p1;   // positions [p0, p1) of the row
    int nseg;     // segments the row is cut into (1 = this one produces the final output)
    int pidx;     // partial slot of this segment
};

// four consecutive columns of a K or V row in the ring: fp32, or bf16 (compact page format) widened
template <bool KVB>
__device__ __forceinline__ float4 ld_row4(const float* row, int col) {
    if constexpr (!KVB) {
        return reinterpret_cast<const float4*>(row)[col];
    } else {
        const uint2 u = reinterpret_cast<const uint2*>(row)[col];
        return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                           __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    }
}

// KVB = true: compact page format (MLI_OPT_KV_FORMAT = 1): K and V rows are bf16, half the bytes
// Two CTAs per SM is what plan_attention() sizes the grid and the rings for, so the register file
// must allow it (<= 96 registers at 288 threads).  Without the bound the NC = 4 instantiations
// (emb_dim 4096) took 131 registers, one CTA per SM was resident and the grid ran as two waves
// (measured at B=128, d=4096, S=2048: 5.0 TB/s, CTA start times spread over 0..610 us; 6.5 TB/s with it).
template <int NC, int G, bool FUSED, bool KVB = false>
__global__ void __launch_bounds__(kAttnThreads, 2)
decode_attention_kernel(const float* __restrict__ q, float* const* __restrict__ page_table,
                        const int* __restrict__ lengths, const int* __restrict__ row_first_g,
                        const int* __restrict__ item_row, const int* __restrict__ item_chunk,
                        float* __restrict__ out, float* __restrict__ part_acc,
                        float* __restrict__ part_ml, float* __restrict__ scores_out,
                        int* __restrict__ row_done, int B, int S, int d, int chunk_pages, int nstage,
                        int min_dyn, int lengths_final, long long* __restrict__ dbg,
                        unsigned long long* trace) {
    // optional phase stamps (tools/attn_timing.py): [cta][16]; slots 0-6 clock64 of consumer thread 0,
    // slot 7 = segments << 32 | stages this CTA processed, slots 8-10 %globaltimer at CTA start / end
    // of the last segment / CTA end, slot 11 = segments merged << 32 | rows merged
#define ATTN_STAMP(slot) do { if (dbg != nullptr && threadIdx.x == 0) dbg[(size_t)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
#define ATTN_GT(slot) do { if (dbg != nullptr && threadIdx.x == 0) dbg[(size_t)blockIdx.x * 16 + (slot)] = (long long)globaltimer_ns(); } while (0)
    ATTN_STAMP(0);
    ATTN_GT(8);
    if (dbg != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        dbg[(size_t)blockIdx.x * 16 + 12] = (long long)smid;
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = S / kPage;
    const int d4 = d >> 2;
    const int row_floats = KVB ? d : 2 * d;  // K row followed by V row (as 4-byte words)
    const int v_off = KVB ? d / 2 : d;       // V row inside a ring row
    const size_t pos_floats = page_pos_floats(d, KVB ? 1 : 0);
    const int stage_floats = G * row_floats;
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)nstage * stage_floats);
    uint64_t* empty_bar = full_bar + kMaxStages;
    float* red = reinterpret_cast<float*>(empty_bar + kMaxStages);  // [2][kConsumerWarps][G]
    int* scan_tmp = reinterpret_cast<int*>(red + 2 * kConsumerWarps * G);   // [2][16]: warp totals of the prefix scan
    int* stage_meta = scan_tmp + 32;                                         // [kMaxStages] slice opened by a stage
    int* pend_r = stage_meta + kMaxStages;                                   // [kMaxPend] partial rows to merge
    int* pend_nseg = pend_r + kMaxPend;
    int* pend_flag = pend_nseg + kMaxPend;
    int* stage_first = pend_flag + kMaxPend;                                   // FUSED: [B + 1]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int chunk_pos = chunk_pages * kPage;

    // Engine launches (lengths_final): the lengths were written two or more kernels up the dependency
    // chain, i.e. they are complete and visible before this kernel can start at all, so the first
    // round of the prefix scan fetches its length here, under the barrier set-up and the dependency
    // wait.  Stage calls keep every load behind the wait.
    int len_early = 0;
    if (FUSED && lengths_final && tid < B) len_early = lengths[tid];
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    griddep_wait();
    GRIDDEP_TRIGGER_EARLY();
    trace_stamp(trace, 3);
    ATTN_STAMP(1);

    int n_items = 0;      // legacy: number of (row, chunk) items
    // fused: the work is flattened in units of pipeline STAGES (G positions of one row; the last
    // stage of a row may be partial).  Measured: a CTA's streaming time is proportional to its number
    // of stages, not to its bytes (2750 cycles per stage, full or not), so slices hold equal numbers
    // of stages.  The space [0, P) is cut into gridDim.x STATIC slices of qs stages (CTA b starts
    // with slice b: a fair share, or ~3/4 of it when shares are large) followed by DYNAMIC slices of
    // qd stages that CTAs claim from a global counter as they run dry.
    int qs = 1, qd = 1, dyn0 = 0, n_slices = 0, P = 0;
    if constexpr (FUSED) {
        // exclusive prefix of the rows' stage counts, kAttnThreads rows at a time: one barrier per
        // round (the warp totals are double-buffered, the running total lives in a register of every
        // thread) and one at the end
        int carry = 0;
        for (int base = 0, round = 0; base < B; base += kAttnThreads, ++round) {
            const int r = base + tid;
            const int L = (r < B) ? ((lengths_final && round == 0) ? len_early : lengths[r]) : 0;
            const int n = (L + G - 1) / G;
            int v = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            int* tot = scan_tmp + 16 * (round & 1);
            if (lane == 31) tot[warp] = v;
            __syncthreads();
            int before = carry, all = 0;
#pragma unroll
            for (int w = 0; w < kAttnThreads / 32; ++w) {
                const int t = tot[w];
                all += t;
                if (w < warp) before += t;
            }
            if (r < B) stage_first[r] = before + v - n;
            carry += all;
        }
        P = carry;
        if (tid == 0) stage_first[B] = P;
        __syncthreads();
        const int grid = (int)gridDim.x;
        const int fair = (P + grid - 1) / grid;
        // small problems (a fair share of a few dozen positions) are split statically: dynamic
        // slices would be a single pipeline stage each and their claims / merges cost more than
        // the tail they remove
        qs = (fair * G >= min_dyn) ? max(1, fair * MLI_ATTN_STATIC_PCT / 100) : max(1, fair);
        dyn0 = (int)min((long long)P, (long long)grid * qs);
        const int dyn = P - dyn0;
        qd = max(1, (dyn + MLI_ATTN_DYN_PARTS * grid - 1) / (MLI_ATTN_DYN_PARTS * grid));
        n_slices = grid + (dyn + qd - 1) / qd;
    } else {
        n_items = row_first_g[B];
    }
    // largest r in [0, B) with arr[r] <= x (arr non-decreasing, arr[0] <= x).  Warp-cooperative: every
    // lane probes one candidate per round (32-ary search: two rounds up to 1024 rows instead of ten
    // dependent shared-memory loads; 72.5 -> 71.8 us per engine step); all 32 lanes of the warp must
    // call it together.
    auto find_row = [&](const int* arr, int x) -> int {
        int lo = 0, n = B;
        while (n > 1) {
            const int step = (n + 31) >> 5;
            const int idx = lo + lane * step;
            const bool le = (idx < lo + n) && (arr[idx] <= x);
            const unsigned m = __ballot_sync(0xffffffffu, le);
            const int k = 31 - __clz(m);   // lane 0 always qualifies
            const int end = lo + n;
            lo += k * step;
            n = min(step, end - lo);
        }
        return lo;
    };
    auto slice_start = [&](int sl) -> int {
        return sl < (int)gridDim.x ? min(P, sl * qs) : min(P, dyn0 + (sl - (int)gridDim.x) * qd);
    };
    auto slice_of = [&](int pos) -> int {
        return pos < dyn0 ? pos / qs : (int)gridDim.x + (pos - dyn0) / qd;
    };
    int slice = 0, g0 = 0, g1 = 0;   // fused: current slice and its range of stages

    // work iterator, identical in the producer and the consumers.  `cur` is a global stage index
    // (fused) or an item index (legacy).
    auto next_seg = [&](int& cur, AttnSeg& sg) -> bool {
        if constexpr (FUSED) {
            if (cur >= g1) return false;
            const int lo = find_row(stage_first, cur);   // (empty rows share a start: skipped)
            const int start = stage_first[lo], n_st = stage_first[lo + 1] - start;   // in stages
            const int st1 = min(g1 - start, n_st);
            sg.r = lo;
            sg.p0 = (cur - start) * G;
            sg.p1 = min(st1 * G, lengths[lo]);   // only the row's last stage can be partial
            sg.nseg = slice_of(start + n_st - 1) - slice_of(start) + 1;
            sg.pidx = 2 * slice + (cur == g0 ? 0 : 1);
            cur = start + st1;
            return true;
        } else {
            if (cur >= n_items) return false;
            const int r = item_row[cur], c = item_chunk[cur];
            const int L = lengths[r];
            sg.r = r;
            sg.p0 = c * chunk_pos;
            sg.p1 = min(L, sg.p0 + chunk_pos);
            sg.nseg = (L + chunk_pos - 1) / chunk_pos;
            sg.pidx = cur;
            cur += gridDim.x;
            return true;
        }
    };
    int cur = (int)blockIdx.x;   // legacy: first item
    AttnSeg sg;
    ATTN_STAMP(2);

    if (warp == kConsumerWarps) {
        // ===================== producer warp =====================
        uint32_t it = 0;  // running stage counter across segments
        // K|V rows are read once per step and the cache of a step (~100 MB on the bench job) is as
        // large as the L2: mark them evict-first so that they do not flush what IS re-read every
        // step (the split weights, the scheduler's state, the work lists)
        const uint64_t kv_policy = l2_policy_evict_first();
        slice = (int)blockIdx.x;
        for (;;) {
            bool open_slice = false;
            if constexpr (FUSED) {
                if (slice >= n_slices) break;
                g0 = slice_start(slice);
                g1 = slice_start(slice + 1);
                cur = g0;
                open_slice = true;
            }
            // The iterator runs one segment ahead: the page pointers of the next segment are loaded
            // while the current one streams.  A bandwidth-bound ring cannot win back an issue gap, and
            // with rows of a few dozen positions every segment boundary used to open one (measured:
            // ~2000 cycles per segment).
            AttnSeg sg_nx;
            bool have = next_seg(cur, sg);
            auto first_pages = [&](const AttnSeg& g) -> const float* {
                const int pg = g.p0 / kPage + lane;
                return (pg * kPage < g.p1) ? page_table[(size_t)g.r * W + pg] : nullptr;
            };
            const float* pages_cur = have ? first_pages(sg) : nullptr;
            while (have) {
                const bool have_nx = next_seg(cur, sg_nx);
                const float* pages_nx = have_nx ? first_pages(sg_nx) : nullptr;
                const int r = sg.r, p1 = sg.p1;
                // page pointers are held 32 pages at a time: lane i holds page (pgb + i)
                int pgb = sg.p0 / kPage;
                const float* my_page = pages_cur;
                for (int pos = sg.p0; pos < p1; pos += G, ++it) {
                    const int stage = it % nstage;
                    const uint32_t parity = (it / nstage) & 1u;
                    const int nvalid = min(G, p1 - pos);
                    if ((pos + nvalid - 1) / kPage >= pgb + 32) {
                        pgb = pos / kPage;
                        const int pg = pgb + lane;
                        my_page = (pg * kPage < p1) ? page_table[(size_t)r * W + pg] : nullptr;
                    }
                    if (lane == 0) {
                        mbar_wait(&empty_bar[stage], parity ^ 1u);
                        // the first stage of a slice tells the consumers which slice it opens
                        if (open_slice) stage_meta[stage] = slice;
                        mbar_expect_tx(&full_bar[stage], (uint32_t)nvalid * row_floats * 4u);
                    }
                    open_slice = false;
                    __syncwarp();
                    // lane g copies position pos+g (K|V rows are contiguous: 8*d bytes)
                    const int j = pos + (lane < G ? lane : 0);
                    const int pg = min(j / kPage - pgb, 31);
                    const float* page = reinterpret_cast<const float*>(
                        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(my_page), pg));
                    if (lane < nvalid) {
                        const float* src = page + (size_t)(j & (kPage - 1)) * pos_floats + d;   // K starts d floats in
                        bulk_g2s_hint(ring + (size_t)stage * stage_floats + (size_t)lane * row_floats, src,
                                      (uint32_t)row_floats * 4u, &full_bar[stage], kv_policy);
                    }
                }
                sg = sg_nx;
                pages_cur = pages_nx;
                have = have_nx;
            }
            if constexpr (!FUSED) break;
            if (n_slices == (int)gridDim.x) break;   // no dynamic slices in this launch
            // next slice: claim a dynamic one
            int nxt = 0;
            if (lane == 0) nxt = (int)gridDim.x + atomicAdd(&row_done[B], 1);
            slice = __shfl_sync(0xffffffffu, nxt, 0);
        }
        if constexpr (FUSED) {
            // end-of-work marker: a data-less stage whose meta is -1
            if (lane == 0) {
                const int stage = it % nstage;
                mbar_wait(&empty_bar[stage], ((it / nstage) & 1u) ^ 1u);
                stage_meta[stage] = -1;
                mbar_arrive(&full_bar[stage]);
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    const float sqrt_d = sqrtf((float)d);
    if constexpr (FUSED) {
        // empty rows produce zeros (the reference stores result = 0, paged_attention.cu:289,:323)
        for (int r = blockIdx.x; r < B; r += gridDim.x) {
            if (stage_first[r + 1] != stage_first[r]) continue;
            for (int col = tid; col < d4; col += kConsumerThreads)
                reinterpret_cast<float4*>(out + (size_t)r * d)[col] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    uint32_t it = 0;
    int n_seg_dbg = 0;
    // Partial rows are merged by whoever completes a row's last segment, in slice order.  The
    // fence / atomic / merge latency is kept off the K|V pipeline: rows are queued and flushed after
    // the CTA has run out of slices (or when the queue is full).
    int n_pend = 0;
    auto flush_pending = [&]() {
        if constexpr (FUSED) {
            if (n_pend == 0) return;
            // release: the CTA barrier orders every consumer thread's partial stores before the
            // gpu-scope acq_rel atomic of the arriving thread (cumulativity); acquire: the same
            // atomic, then the barrier, then L1-bypassing loads (__ldcg) by all threads
            named_bar_sync(1, kConsumerThreads);
            if (tid < n_pend)
                pend_flag[tid] = (atom_add_acq_rel_gpu(&row_done[pend_r[tid]], 1) == pend_nseg[tid] - 1) ? 1 : 0;
            named_bar_sync(1, kConsumerThreads);
            for (int pi = 0; pi < n_pend; ++pi) {
                if (!pend_flag[pi]) continue;
                const int r = pend_r[pi], nseg = pend_nseg[pi];
                if (dbg != nullptr && tid == 0) dbg[(size_t)blockIdx.x * 16 + 11] += ((long long)nseg << 32) | 1;
                const int start = stage_first[r];
                const int b_first = slice_of(start);
                // segment k of the row lives in slice b_first + k: its head slot, except that the
                // row's first segment is its slice's tail slot unless the row opens that slice
                auto slot_of = [&](int k) -> size_t {
                    return (size_t)2 * (b_first + k) + ((k == 0 && start != slice_start(b_first)) ? 1 : 0);
                };
                if (NC < 4 && nseg <= 4) {
                    // common case: every load of the merge is issued before anything is consumed
                    float2 ml[4];
                    float4 pv[4][NC];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const size_t sl = slot_of(min(u, nseg - 1));
                        ml[u] = __ldcg(reinterpret_cast<const float2*>(part_ml + 2 * sl));
#pragma unroll
                        for (int i = 0; i < NC; ++i) {
                            const int col = tid + i * kConsumerThreads;
                            pv[u][i] = (col < d4) ? __ldcg(reinterpret_cast<const float4*>(part_acc + sl * d) + col)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    float M = ml[0].x;
#pragma unroll
                    for (int u = 1; u < 4; ++u)
                        if (u < nseg) M = fmaxf(M, ml[u].x);
                    float Lsum = 0.f;
                    float4 a[NC];
#pragma unroll
                    for (int i = 0; i < NC; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (u < nseg) {
                            const float w = expf(ml[u].x - M);
                            Lsum += ml[u].y * w;
#pragma unroll
                            for (int i = 0; i < NC; ++i) {
                                a[i].x = fmaf(w, pv[u][i].x, a[i].x); a[i].y = fmaf(w, pv[u][i].y, a[i].y);
                                a[i].z = fmaf(w, pv[u][i].z, a[i].z); a[i].w = fmaf(w, pv[u][i].w, a[i].w);
                            }
                        }
                    }
                    const float norm = 1.f / Lsum;
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const int col = tid + i * kConsumerThreads;
                        if (col < d4)
                            reinterpret_cast<float4*>(out + (size_t)r * d)[col] =
                                make_float4(a[i].x * norm, a[i].y * norm, a[i].z * norm, a[i].w * norm);
                    }
                    if (tid == 0) row_done[r] = 0;   // ready for the next launch
                    continue;
                }
                // (m, l) of up to 32 segments at a time, one per lane: all loads in flight together
                float M = -INFINITY;
                for (int k0 = 0; k0 < nseg; k0 += 32) {
                    const int k = k0 + lane;
                    float m = (k < nseg) ? __ldcg(part_ml + 2 * slot_of(k)) : -INFINITY;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                    M = fmaxf(M, m);
                }
                float Lsum = 0.f;
                float4 a[NC];
#pragma unroll
                for (int i = 0; i < NC; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k0 = 0; k0 < nseg; k0 += 32) {
                    const int k = k0 + lane;
                    float wl = 0.f, lw = 0.f;
                    if (k < nseg) {
                        const float2 ml = __ldcg(reinterpret_cast<const float2*>(part_ml + 2 * slot_of(k)));
                        wl = expf(ml.x - M);
                        lw = ml.y * wl;
                    }
                    Lsum += warp_sum(lw);
                    const int kn = min(32, nseg - k0);
                    constexpr int MU = NC >= 4 ? 2 : 4;   // segments in flight per round (register budget)
                    for (int kk = 0; kk < kn; kk += MU) {
                        float4 pv[MU][NC];
                        float w[MU];
#pragma unroll
                        for (int u = 0; u < MU; ++u) {
                            w[u] = __shfl_sync(0xffffffffu, wl, min(kk + u, 31));
                            const size_t sl = slot_of(min(k0 + kk + u, nseg - 1));
#pragma unroll
                            for (int i = 0; i < NC; ++i) {
                                const int col = tid + i * kConsumerThreads;
                                pv[u][i] = (col < d4 && kk + u < kn)
                                               ? __ldcg(reinterpret_cast<const float4*>(part_acc + sl * d) + col)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < MU; ++u) {
                            if (kk + u < kn) {
#pragma unroll
                                for (int i = 0; i < NC; ++i) {
                                    a[i].x = fmaf(w[u], pv[u][i].x, a[i].x); a[i].y = fmaf(w[u], pv[u][i].y, a[i].y);
                                    a[i].z = fmaf(w[u], pv[u][i].z, a[i].z); a[i].w = fmaf(w[u], pv[u][i].w, a[i].w);
                                }
                            }
                        }
                    }
                }
                const float norm = 1.f / Lsum;
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const int col = tid + i * kConsumerThreads;
                    if (col < d4)
                        reinterpret_cast<float4*>(out + (size_t)r * d)[col] =
                            make_float4(a[i].x * norm, a[i].y * norm, a[i].z * norm, a[i].w * norm);
                }

                if (tid == 0) row_done[r] = 0;   // ready for the next launch
            }
            named_bar_sync(1, kConsumerThreads);
            n_pend = 0;
        }
    };
    ATTN_STAMP(3);
    for (;;) {
    if constexpr (FUSED) {
        // the next stage of the ring opens a slice (or ends the work): learn which
        const int stage = it % nstage;
        mbar_wait(&full_bar[stage], (it / nstage) & 1u);
        slice = stage_meta[stage];
        if (slice < 0) break;
        g0 = slice_start(slice);
        g1 = slice_start(slice + 1);
        cur = g0;
        if (n_pend + 2 > kMaxPend) flush_pending();
    }
    while (next_seg(cur, sg)) {
        const int r = sg.r, p0 = sg.p0, p1 = sg.p1;

        float4 qv[NC];
        float4 acc[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int col = tid + i * kConsumerThreads;
            qv[i] = (col < d4) ? reinterpret_cast<const float4*>(q + (size_t)r * d)[col]
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float m_run = -INFINITY, l_run = 0.f;

        for (int pos = p0; pos < p1; pos += G, ++it) {
            const int stage = it % nstage;
            const uint32_t parity = (it / nstage) & 1u;
            const int nvalid = min(G, p1 - pos);
            const float* sbase = ring + (size_t)stage * stage_floats;
            float* red_buf = red + (size_t)(it & 1u) * kConsumerWarps * G;
            mbar_wait(&full_bar[stage], parity);
            if (it == 0) ATTN_STAMP(4);

            // ---- phase A: partial q.K over this thread's columns, warp reduce ----
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float s = 0.f;
                if (g < nvalid) {
                    const float* krow = sbase + (size_t)g * row_floats;
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const int col = tid + i * kConsumerThreads;
                        if (col < d4) {
                            const float4 k = ld_row4<KVB>(krow, col);
                            s = fmaf(qv[i].x, k.x, s);
                            s = fmaf(qv[i].y, k.y, s);
                            s = fmaf(qv[i].z, k.z, s);
                            s = fmaf(qv[i].w, k.w, s);
                        }
                    }
                }
                s = warp_sum(s);
                if (lane == 0) red_buf[warp * G + g] = s;
            }
            named_bar_sync(1, kConsumerThreads);

            // ---- scores: each warp reduces the 8 x G warp partials with shuffles.  Lane l ends up
            // with the score of position g = l % G, so the division, the running max and the
            // exponentials are computed once per lane instead of G times per thread ----
            float tot = 0.f;
#pragma unroll
            for (int i = 0; i < (kConsumerWarps * G + 31) / 32; ++i) {
                const int idx = lane + 32 * i;
                if (idx < kConsumerWarps * G) tot += red_buf[idx];   // idx = w * G + g, 32 % G == 0
            }
#pragma unroll
            for (int o = G; o < 32 && o < kConsumerWarps * G; o <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            const int myg = lane % G;
            tot = __shfl_sync(0xffffffffu, tot, myg);   // lanes >= 8 * G took no part in the reduction
            const float sc = tot / sqrt_d;
            float m_stage = (myg < nvalid) ? sc : -INFINITY;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) m_stage = fmaxf(m_stage, __shfl_xor_sync(0xffffffffu, m_stage, o));
            const float m_new = fmaxf(m_run, m_stage);
            if (scores_out != nullptr) {
                if (tid < G && tid < nvalid) scores_out[(size_t)r * S + pos + tid] = sc;
            }
            const float corr = expf(m_run - m_new);  // exp(-inf) = 0 on the first stage
            const float p_mine = (myg < nvalid) ? expf(sc - m_new) : 0.f;
            float p_sum = p_mine;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) p_sum += __shfl_xor_sync(0xffffffffu, p_sum, o);
            l_run = l_run * corr + p_sum;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                acc[i].x *= corr; acc[i].y *= corr; acc[i].z *= corr; acc[i].w *= corr;
            }
            // ---- phase B: P.V ----
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float p = __shfl_sync(0xffffffffu, p_mine, g);
                if (g < nvalid) {
                    const float* vrow = sbase + (size_t)g * row_floats + v_off;
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const int col = tid + i * kConsumerThreads;
                        if (col < d4) {
                            const float4 v = ld_row4<KVB>(vrow, col);
                            acc[i].x = fmaf(p, v.x, acc[i].x);
                            acc[i].y = fmaf(p, v.y, acc[i].y);
                            acc[i].z = fmaf(p, v.z, acc[i].z);
                            acc[i].w = fmaf(p, v.w, acc[i].w);
                        }
                    }
                }
            }
            m_run = m_new;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
        }

        // ---- segment epilogue ----
        ++n_seg_dbg;
        ATTN_STAMP(5);
        ATTN_GT(9);
        const int nseg = sg.nseg;
        const size_t pidx = (size_t)sg.pidx;
        if (nseg == 1) {
            if (!FUSED && tid == 0) {
                part_ml[2 * pidx] = m_run;
                part_ml[2 * pidx + 1] = l_run;
            }
            const float norm = 1.f / l_run;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int col = tid + i * kConsumerThreads;
                if (col < d4)
                    reinterpret_cast<float4*>(out + (size_t)r * d)[col] = make_float4(
                        acc[i].x * norm, acc[i].y * norm, acc[i].z * norm, acc[i].w * norm);
            }
        } else {
            if (tid == 0) {
                part_ml[2 * pidx] = m_run;
                part_ml[2 * pidx + 1] = l_run;
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int col = tid + i * kConsumerThreads;
                if (col < d4) reinterpret_cast<float4*>(part_acc + pidx * d)[col] = acc[i];
            }
            if constexpr (FUSED) {
                // queued (at most a head and a tail row per slice)
                if (tid == 0) {
                    pend_r[n_pend] = r;
                    pend_nseg[n_pend] = nseg;
                }
                ++n_pend;
            }
        }
    }
    if constexpr (!FUSED) break;
    }   // slices
    GRIDDEP_TRIGGER_LATE();
    flush_pending();
    if constexpr (FUSED) {
        // the last CTA to finish re-arms the slice counter for the next launch
        if (tid == 0 && n_slices != (int)gridDim.x) {
            __threadfence();
            if (atomicAdd(&row_done[B + 1], 1) == (int)gridDim.x - 1) {
                row_done[B] = 0;
                row_done[B + 1] = 0;
            }
        }
    }
    ATTN_STAMP(6);
    ATTN_GT(10);
    if (dbg != nullptr && threadIdx.x == 0) {
        dbg[(size_t)blockIdx.x * 16 + 7] = ((long long)n_seg_dbg << 32) | it;
    }
#undef ATTN_STAMP
#undef ATTN_GT
}

// ---------------------------------------------------------------------------------------------
// combine: one CTA per row.  nchunks == 0 -> zeros (the reference stores result = 0 for empty
// rows, paged_attention.cu:289,:323); nchunks == 1 -> already final; else merge partials.
// Optionally turns the raw scores in scores_out into the reference's probabilities
// expf(s - max) * (1.f / sum), zeros past L (self_attention_inference_optimized.cu:226-241).
// ---------------------------------------------------------------------------------------------
__global__ void attn_combine_kernel(const int* __restrict__ lengths,
                                    const int* __restrict__ row_first,
                                    const float* __restrict__ part_acc,
                                    const float* __restrict__ part_ml, float* __restrict__ out,
                                    float* __restrict__ scores_out, int S, int d) {
    const int r = blockIdx.x;
    const int first = row_first[r];
    const int n = row_first[r + 1] - first;
    const int d4 = d >> 2;
    const int L = lengths[r];
    float M = -INFINITY, Lsum = 0.f;
    if (n > 0) {
        for (int i = 0; i < n; ++i) M = fmaxf(M, part_ml[2 * (size_t)(first + i)]);
        for (int i = 0; i < n; ++i)
            Lsum += part_ml[2 * (size_t)(first + i) + 1] * expf(part_ml[2 * (size_t)(first + i)] - M);
    }
    if (n == 0) {
        for (int col = threadIdx.x; col < d4; col += blockDim.x)
            reinterpret_cast<float4*>(out + (size_t)r * d)[col] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (n > 1) {
        const float norm = 1.f / Lsum;
        for (int col = threadIdx.x; col < d4; col += blockDim.x) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < n; ++i) {
                const float w = expf(part_ml[2 * (size_t)(first + i)] - M);
                const float4 p = reinterpret_cast<const float4*>(part_acc + (size_t)(first + i) * d)[col];
                a.x = fmaf(w, p.x, a.x); a.y = fmaf(w, p.y, a.y);
                a.z = fmaf(w, p.z, a.z); a.w = fmaf(w, p.w, a.w);
            }
            reinterpret_cast<float4*>(out + (size_t)r * d)[col] =
                make_float4(a.x * norm, a.y * norm, a.z * norm, a.w * norm);
        }
    }
    if (scores_out != nullptr) {
        const float norm = (n > 0) ? 1.f / Lsum : 0.f;
        float* row = scores_out + (size_t)r * S;
        for (int j = threadIdx.x; j < S; j += blockDim.x)
            row[j] = (j < L) ? expf(row[j] - M) * norm : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
struct AttnPlan {
    int NC, G, nstage, chunk_pages, grid;
    size_t smem;
    int max_items;
};

static int plan_attention(mli_ctx* ctx, int B, int S, int d, bool fused, AttnPlan* p) {
    const int W = S / kPage;
    if (d > 4096) {
        set_error("decode attention: emb_dim > 4096 is not supported by this build");
        return MLI_ERR_UNSUPPORTED;
    }
    p->NC = (d + 1023) / 1024;
    if (p->NC == 3) p->NC = 4;
    const int kvb = ctx->kv_bf16;
    const int row_words = kvb ? d : 2 * d;          // 4-byte words of one K|V row in the ring
    int G = 1;
    while (G < 16 && G * row_words <= 4096) G <<= 1;  // a stage of at most 32 KB
    // (halving the stage for short contexts was measured: 16 KB stages run at 9 B/cycle per CTA against
    // 12 for 32 KB ones -- the consumers' per-stage barrier / shuffle / exp chain does not shrink)
    p->G = G;
    const size_t stage_bytes = (size_t)G * row_words * 4;
    int ctas = ctx->attn_ctas_per_sm > 0 ? ctx->attn_ctas_per_sm : 2;
    // the fused kernel also keeps the [B+1] position prefix in shared memory
    const size_t prefix_bytes = fused ? sizeof(int) * ((size_t)B + 1) : 0;
    const size_t budget = ((ctas >= 2) ? 108 * 1024 : 216 * 1024) - prefix_bytes - 1024;
    int nstage = (int)(budget / stage_bytes);
    if (nstage < 2) nstage = 2;
    if (nstage > kMaxStages) nstage = kMaxStages;
    p->nstage = nstage;
    p->smem = nstage * stage_bytes + 2 * kMaxStages * sizeof(uint64_t) +
              2 * kConsumerWarps * G * sizeof(float) + kAttnCtrlInts * sizeof(int) + 128;
    int ch = ctx->attn_chunk_pages;
    if (ch <= 0) {
        // aim for a few items per persistent CTA without making items tiny
        long long pages = (long long)B * W;
        ch = (int)((pages + (long long)ctx->num_sms * 16 - 1) / ((long long)ctx->num_sms * 16));
        if (ch < 2) ch = 2;
        if (ch > 16) ch = 16;
    }
    if (ch > 32) ch = 32;
    if (ch > W) ch = W;
    p->chunk_pages = ch;
    p->max_items = B * ceil_div(W, ch);
    p->grid = ctx->num_sms * ctas;
    if (p->grid > p->max_items) p->grid = p->max_items;
    if (p->grid < 1) p->grid = 1;
    return 0;
}

size_t attention_meta_bytes(int B, int max_items) {
    return sizeof(int) * ((size_t)B + 1 + 2 * (size_t)max_items + 8);
}

template <int NC, int G, bool FUSED, bool KVB = false>
static int launch_main(const AttnPlan& p, mli_ctx* ctx, const float* q, float* const* page_table,
                       const int* lengths, const int* row_first, const int* item_row,
                       const int* item_chunk, float* out, float* part_acc, float* part_ml,
                       float* scores_out, int* row_done, int B, int S, int d) {
    auto kern = decode_attention_kernel<NC, G, FUSED, KVB>;
    const size_t smem = p.smem + (FUSED ? sizeof(int) * ((size_t)B + 1) : 0);
    { int rc0 = ensure_dyn_smem(ctx, kern, smem); if (rc0) return rc0; }
    if (ctx->attn_ev_start) MLI_CUDA(cudaEventRecord(ctx->attn_ev_start, ctx->stream));
    int rc = launch_kernel(ctx, kern, dim3(p.grid), dim3(kAttnThreads), smem, q, page_table, lengths,
                           row_first, item_row, item_chunk, out, part_acc, part_ml, scores_out, row_done,
                           B, S, d, p.chunk_pages, p.nstage, ctx->attn_min_dyn, ctx->attn_lengths_final,
                           reinterpret_cast<long long*>(ctx->tc_dbg), ctx->trace);
    if (rc) return rc;
    if (ctx->attn_ev_stop) MLI_CUDA(cudaEventRecord(ctx->attn_ev_stop, ctx->stream));
    return 0;
}

int launch_decode_attention_paged(mli_ctx* ctx, const float* q, float* const* page_table,
                                  const int* lengths, float* out, float* softmax_out, int B, int S,
                                  int d) {
    AttnPlan p;
    const bool fused = (softmax_out == nullptr) && B <= kMaxFusedRows;
    int rc = plan_attention(ctx, B, S, d, fused, &p);
    if (rc) return rc;
    void* meta = nullptr;
    void* part = nullptr;
    void* cnt = nullptr;
    rc = ws_get(ctx, WS_ATTN_META, attention_meta_bytes(B, p.max_items), &meta);
    if (rc) return rc;
    // partial slots: one per (row, chunk) item (legacy) or two per CTA (fused: slice head / tail)
    // fused: <= 4 * grid + grid slices, for either kernel's grid
    const size_t n_part = std::max((size_t)p.max_items, (size_t)2 * 5 * std::max(p.grid, 2 * ctx->num_sms));
    rc = ws_get(ctx, WS_ATTN_PART, sizeof(float) * (n_part * (d + 2) + 8), &part);
    if (rc) return rc;
    rc = ws_get_zeroed(ctx, WS_ATTN_CNT, sizeof(int) * ((size_t)B + 2), &cnt);   // + slice / finished-CTA counters
    if (rc) return rc;
    int* row_first = reinterpret_cast<int*>(meta);
    int* item_row = row_first + B + 1;
    int* item_chunk = item_row + p.max_items;
    int* row_done = reinterpret_cast<int*>(cnt);
    float* part_ml = reinterpret_cast<float*>(part);
    float* part_acc = part_ml + 2 * n_part;
    // keep part_acc 16-byte aligned
    part_acc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(part_acc) + 15) & ~(uintptr_t)15);

    // Which consumer design (MLI_OPT_ATTN_KERNEL, 0 = auto).  Measured on B200 (tools/attn_sweep.py,
    // profiles/): the warp-per-position kernel is at or above the column-split one on every long
    // launch -- fp32 pages 7.0-7.3 TB/s either way, bf16 pages 6.5-7.1 vs 5.1-6.1 TB/s, emb_dim 512
    // 6.5 vs 5.5 TB/s -- but its one CTA per SM pays every row boundary with an SM-wide merge, so on
    // launches of a few dozen positions per CTA (the configs[1] step: 28.5 vs 25.0 us) the two
    // interleaved CTAs per SM of the column-split kernel win.  Lengths live on the device, so the
    // choice goes by what the launch can hold: B * S positions, at least 1024 per SM -> warp-per-position.
    const bool wp_auto = (long long)B * S >= 1024LL * ctx->num_sms;
    if (fused && attention_wp_usable(ctx, B, d) && (ctx->attn_kernel == 2 || (ctx->attn_kernel == 0 && wp_auto)))
        return launch_decode_attention_wp(ctx, q, page_table, lengths, out, part_acc, part_ml, row_done, B,
                                          S, d, ctx->attn_min_dyn);
    if (ctx->attn_kernel == 2 && fused) {
        set_error("decode attention: the warp-per-position kernel has no instantiation for this emb_dim");
        return MLI_ERR_UNSUPPORTED;
    }

    if (!fused) {
        attn_prep_kernel<<<1, 1024, 0, ctx->stream>>>(lengths, B, p.chunk_pages * kPage, row_first,
                                                       item_row, item_chunk);
        MLI_LAUNCH_CHECK();
    }

    if (ctx->kv_bf16) {
        if (!fused) {
            set_error("compact KV format: the [B,S] probabilities / more than 8192 rows are not supported");
            return MLI_ERR_UNSUPPORTED;
        }
#define MLI_ATTN_BF16(NC_, G_)                                                                       \
    if (p.NC == NC_ && p.G == G_)                                                                    \
        return launch_main<NC_, G_, true, true>(p, ctx, q, page_table, lengths, row_first, item_row, \
                                                item_chunk, out, part_acc, part_ml, softmax_out,     \
                                                row_done, B, S, d);
        MLI_ATTN_BF16(1, 16)
        MLI_ATTN_BF16(1, 8)
        MLI_ATTN_BF16(2, 4)
        MLI_ATTN_BF16(4, 2)
#undef MLI_ATTN_BF16
        set_error("compact KV format: no attention instantiation for this emb_dim");
        return MLI_ERR_UNSUPPORTED;
    }

#define MLI_ATTN_CASE(NC_, G_)                                                                       \
    if (p.NC == NC_ && p.G == G_)                                                                    \
        rc = fused ? launch_main<NC_, G_, true>(p, ctx, q, page_table, lengths, row_first, item_row, \
                                                item_chunk, out, part_acc, part_ml, softmax_out,     \
                                                row_done, B, S, d)                                   \
                   : launch_main<NC_, G_, false>(p, ctx, q, page_table, lengths, row_first, item_row, \
                                                 item_chunk, out, part_acc, part_ml, softmax_out,    \
                                                 row_done, B, S, d);                                 \
    else
    MLI_ATTN_CASE(1, 16)
    MLI_ATTN_CASE(1, 8)
    MLI_ATTN_CASE(1, 4)
    MLI_ATTN_CASE(1, 2)
    MLI_ATTN_CASE(1, 1)
    MLI_ATTN_CASE(2, 2)
    MLI_ATTN_CASE(2, 1)
    MLI_ATTN_CASE(4, 1) {
        set_error("decode attention: no kernel instantiation for this emb_dim");
        return MLI_ERR_UNSUPPORTED;
    }
#undef MLI_ATTN_CASE
    if (rc) return rc;

    if (!fused) {
        attn_combine_kernel<<<B, 256, 0, ctx->stream>>>(lengths, row_first, part_acc, part_ml, out,
                                                        softmax_out, S, d);
        MLI_LAUNCH_CHECK();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// dense (non-paged) decode attention for BASELINE config C1: kt_cache is TRANSPOSED [B,d,S]
// (self_attention_inference_optimized.cu:150-279).  One CTA per row; scores are accumulated
// c-ascending per position and P.V j-ascending per column, i.e. in the reference's own order
// (qkt :174-176, softmax_v :268-270), so only expf / the softmax sums differ from it.  C1 is
// L2-resident and launch-bound, so this kernel is written for fidelity, not bandwidth.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_attention_kernel(const float* __restrict__ q, const float* __restrict__ kt,
                       const float* __restrict__ v, const int* __restrict__ lengths,
                       float* __restrict__ out, float* __restrict__ softmax_out, int S, int d) {
    extern __shared__ float dsm[];
    float* p = dsm;          // [S]
    float* qs = dsm + S;     // [d]
    __shared__ float red[32];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = lengths[r];
    for (int c = tid; c < d; c += 256) qs[c] = q[(size_t)r * d + c];
    __syncthreads();
    const float sqrt_d = sqrtf((float)d);
    const float* ktr = kt + (size_t)r * d * S;
    float lmax = -FLT_MAX;
    for (int j = tid; j < L; j += 256) {
        float s = 0.f;
        for (int c = 0; c < d; ++c) s = fmaf(qs[c], ktr[(size_t)c * S + j], s);
        s = s / sqrt_d;
        p[j] = s;
        lmax = fmaxf(lmax, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float gmax = -FLT_MAX;
    for (int w = 0; w < 8; ++w) gmax = fmaxf(gmax, red[w]);
    __syncthreads();
    float lsum = 0.f;
    for (int j = tid; j < L; j += 256) lsum += expf(p[j] - gmax);
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    float gsum = 0.f;
    for (int w = 0; w < 8; ++w) gsum += red[w];
    const float norm = 1.f / gsum;
    for (int j = tid; j < L; j += 256) p[j] = expf(p[j] - gmax) * norm;
    __syncthreads();
    const float* vr = v + (size_t)r * S * d;
    for (int c = tid; c < d; c += 256) {
        float a = 0.f;
        for (int j = 0; j < L; ++j) a = fmaf(p[j], vr[(size_t)j * d + c], a);
        out[(size_t)r * d + c] = a;
    }
    if (softmax_out != nullptr)
        for (int j = tid; j < S; j += 256) softmax_out[(size_t)r * S + j] = (j < L) ? p[j] : 0.f;
}

int launch_decode_attention_dense(mli_ctx* ctx, const float* q, const float* kt_cache,
                                  const float* v_cache, const int* lengths, float* out,
                                  float* softmax_out, int B, int S, int d) {
    const size_t smem = sizeof(float) * ((size_t)S + d);
    if (smem > 200 * 1024) {
        set_error("dense attention: n_sequence + emb_dim too large for shared memory");
        return MLI_ERR_UNSUPPORTED;
    }
    { int rc0 = ensure_dyn_smem(ctx, dense_attention_kernel, smem); if (rc0) return rc0; }
    dense_attention_kernel<<<B, 256, smem, ctx->stream>>>(q, kt_cache, v_cache, lengths, out,
                                                          softmax_out, S, d);
    MLI_LAUNCH_CHECK();
    return 0;
}

// algorithmic bytes of one decode-attention launch, from host-side lengths
double attention_algorithmic_bytes(const int* lengths_host, int B, int d, int kv_bf16) {
    const double kv = kv_bf16 ? 4.0 : 8.0;   // K and V bytes per position and column
    double total = 0;
    for (int r = 0; r < B; ++r) {
        int L = lengths_host[r];
        if (L > 0) total += kv * d * L + 8.0 * d + 8.0 * ((L + kPage - 1) / kPage) + 4.0;
    }
    return total;
}

}  // namespace mli
