// Fused paged decode attention, "warp per position" consumers (one launch per decode step).
//
// Same job, work split and cross-CTA merge protocol as decode_attention_kernel<FUSED = true>
// (decode_attention.cu; reference: launch_qkt_paged_attention + launch_softmax_in_place_with_lengths +
// launch_softmax_v_paged_attention, src/kernels/paged_attention.cu:208-345), different consumer side:
//
//   * decode_attention_kernel splits the COLUMNS of one position over its 8 consumer warps, so every
//     position costs a shuffle reduction, a shared-memory exchange and a CTA barrier; its consumers
//     run in lock step, one pipeline stage at a time.  That chain is ~500 + 300 * (d / 1024) cycles
//     per position (tools/attn_timing.py): hidden behind HBM for fp32 pages, but the limit for the
//     compact bf16 page format (half the bytes per position) and for emb_dim < 1024.
//   * here ONE CTA per SM runs 16 warps (a producer and 15 / 14 / 12 consumers), and a position
//     belongs to one warp (emb_dim <= 1024) or to CW = emb_dim / 1024 warps that split its columns
//     (a named barrier of those CW warps per position).  Every warp (group) keeps its
//     own online-softmax state (m, l, acc) over the positions it was dealt, so there is no CTA-wide
//     synchronisation inside a row segment at all: positions of a pipeline stage are processed
//     concurrently by different warps, and warps drift apart freely.  The per-group states are merged
//     through shared memory once per segment.
//
//   * a warp visits only the ring stages that hold its positions (see wp_wait_issued), slice ids
//     travel in a small box of their own, and at a row boundary the next row's q is fetched while
//     the groups' states are merged.
// Chosen by launch_decode_attention_paged for launches that can hold >= 1024 positions per SM
// (measured against the column-split kernel in profiles/r1_attn_sweep.jsonl).
//
// Scale is dot / sqrtf(d) and the exponent is expf, as in the reference.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cfloat>

namespace mli {

namespace {

// 16 warps per CTA (one CTA per SM): four warps per scheduler leaves 128 registers per thread, enough
// for a lane's 32 q columns and 32 accumulators without spills (a 17th warp would cap it at 96).
// Warp 15 is the producer; the consumers are the first 15 warps (one warp per position), 14 (two
// warps per position) or 12 (four).
constexpr int kWpThreads = 512;
constexpr int kWpProducerWarp = 15;
__host__ __device__ constexpr int wp_consumer_warps(int cw) { return cw == 1 ? 15 : (cw == 2 ? 14 : 12); }
constexpr int kWpMaxStages = 16;
constexpr int kWpMaxPend = 32;
constexpr int kWpCtrlInts = 40 + kWpMaxStages + 4 * kWpMaxPend;
// rows per entry of the coarse stage prefix (launches with more than kWpFineRows rows, see wp_table_ints)
constexpr int kWpBlk = 32;
constexpr int kWpFineRows = 2048;
// shared-memory ints of the per-row tables: up to kWpFineRows rows keep the exact stage prefix and the
// lengths of every row ([B + 1] + [B]); beyond that the ring would lose half its depth to them (at 8192
// rows the two tables are 64 KB: 4 stages instead of 8 at emb_dim 1024, measured 5.8 instead of 7.0 TB/s),
// so only the prefix of every 32nd row is kept and a segment's row is located with one extra (L2-resident)
// load of 32 lengths
__host__ __device__ inline size_t wp_table_ints(int B) {
    return B <= kWpFineRows ? 2 * (size_t)B + 1 : (size_t)(B + kWpBlk - 1) / kWpBlk + 1;
}

struct WpSeg {
    int r;        // batch row
    int p0, p1;   // positions [p0, p1) of the row
    int nseg;     // segments the row is cut into (1 = this one produces the final output)
    int pidx;     // partial slot of this segment
    int start;    // first stage of the row in the flattened stage space
};

// four consecutive columns (float4 index col) of a K or V row in the ring
template <bool KVB>
__device__ __forceinline__ float4 wp_ld4(const unsigned char* row, int col) {
    if constexpr (!KVB) {
        return reinterpret_cast<const float4*>(row)[col];
    } else {
        const uint2 u = reinterpret_cast<const uint2*>(row)[col];
        return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                           __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    }
}

// A consumer warp only visits the ring stages that hold one of its positions, so it may come to a
// slot whose PREVIOUS use it never waited for.  A parity wait cannot tell "this use has landed" from
// "the previous use has not landed yet" (bulk copies complete out of order), so the warp first waits
// until the producer has ISSUED the stage: the producer issues a use only after every owner of the
// slot's previous use has released it, i.e. after that use had landed.  Bounded like mbar_wait().
__device__ __forceinline__ void wp_wait_issued(const volatile uint32_t* issued, uint32_t stage_no) {
    if (*issued > stage_no) return;
    const long long t0 = clock64();
    while (*issued <= stage_no) {
        if (clock64() - t0 > 4000000000LL) {
            printf("mli: attention ring wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// NCW: float4 columns per lane; CW: warps that share one position (its columns are split over them);
// a row has d = 128 * NCW * CW columns.  KVB: compact page format (K, V rows bf16).
// G (positions per pipeline stage, <= 32) and nstage are run-time.
template <int NCW, int CW, bool KVB>
__global__ void __launch_bounds__(kWpThreads, 1)
decode_attention_wp_kernel(const float* __restrict__ q, float* const* __restrict__ page_table,
                           const int* __restrict__ lengths, float* __restrict__ out,
                           float* __restrict__ part_acc, float* __restrict__ part_ml,
                           int* __restrict__ row_done, int B, int S, int d, int G, int nstage,
                           int min_dyn, long long* __restrict__ dbg, unsigned long long* trace) {
    // optional phase stamps, same slots as decode_attention_kernel (tools/attn_timing.py)
#define WP_STAMP(slot) do { if (dbg != nullptr && threadIdx.x == 0) dbg[(size_t)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
#define WP_GT(slot) do { if (dbg != nullptr && threadIdx.x == 0) dbg[(size_t)blockIdx.x * 16 + (slot)] = (long long)globaltimer_ns(); } while (0)
    WP_STAMP(0);
    WP_GT(8);
    constexpr int kWpConsumerWarps = wp_consumer_warps(CW);
    constexpr int kWpConsumerThreads = kWpConsumerWarps * 32;
    constexpr int PW = kWpConsumerWarps / CW;   // position groups
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = S / kPage;
    const int d4 = d >> 2;
    const int rowb = KVB ? 4 * d : 8 * d;     // bytes of one K|V row
    const int voff = KVB ? 2 * d : 4 * d;     // V row inside it
    const size_t pos_floats = page_pos_floats(d, KVB ? 1 : 0);
    const int stage_bytes = G * rowb;
    // G and nstage are powers of two (launcher): stage / parity / position-in-stage are shifts and masks
    const uint32_t stage_mask = (uint32_t)nstage - 1u;
    const int stage_lg = 31 - __clz(nstage);
    const int g_lg = 31 - __clz(G);
    unsigned char* ring = smem_raw;
    float* scratch = reinterpret_cast<float*>(ring + (size_t)nstage * stage_bytes);   // [PW][d] group accumulators
    float* sml = scratch + (size_t)PW * d;                                            // [PW][2] group (m, l)
    float* xch = sml + 2 * PW;                                                        // [PW][2][CW] partial scores
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(xch + 2 * PW * CW);
    uint64_t* empty_bar = full_bar + kWpMaxStages;
    uint64_t* sl_full = empty_bar + kWpMaxStages;    // [2] slice hand-off, producer -> consumers
    uint64_t* sl_empty = sl_full + 2;                // [2]
    int* scan_tmp = reinterpret_cast<int*>(sl_empty + 2);               // [40]: warp totals, carry
    int* slice_box = scan_tmp + 40;                                     // [2] slice ids in flight (-1 = no more work)
    volatile uint32_t* issued = reinterpret_cast<volatile uint32_t*>(slice_box + 2);   // stages the producer has issued
    int* pend_r = slice_box + kWpMaxStages;                             // [kWpMaxPend] partial rows to merge
    int* pend_nseg = pend_r + kWpMaxPend;
    int* pend_flag = pend_nseg + kWpMaxPend;
    int* pend_start = pend_flag + kWpMaxPend;
    const bool coarse = B > kWpFineRows;
    const int NB = (B + kWpBlk - 1) / kWpBlk;                           // coarse: blocks of 32 rows
    int* stage_first = pend_start + kWpMaxPend;                         // fine [B + 1] / coarse [NB + 1]
    int* len_s = stage_first + B + 1;                                   // fine only: [B] the rows' lengths

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], (uint32_t)(G * CW));   // one arrival per (position, column part)
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sl_full[s], 1);
            mbar_init(&sl_empty[s], kWpConsumerWarps);
        }
        mbar_fence_init();
        *issued = 0u;
    }
    __syncthreads();
    griddep_wait();
    GRIDDEP_TRIGGER_EARLY();
    trace_stamp(trace, 3);
    WP_STAMP(1);

    // ---- prefix of the rows' stage counts (a stage = G positions of one row): one barrier per round
    // (double-buffered warp totals, running total in a register of every thread) and one at the end ----
    int P = 0;
    for (int base = 0, round = 0; base < B; base += kWpThreads, ++round) {
        const int r = base + tid;
        const int L = (r < B) ? lengths[r] : 0;
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
        int before = P, all = 0;
#pragma unroll
        for (int w = 0; w < kWpThreads / 32; ++w) {
            const int t = tot[w];
            all += t;
            if (w < warp) before += t;
        }
        if (r < B) {
            if (!coarse) {
                stage_first[r] = before + v - n;
                len_s[r] = L;
            } else if ((r & (kWpBlk - 1)) == 0) {
                stage_first[r / kWpBlk] = before + v - n;
            }
        }
        P += all;
    }
    if (tid == 0) stage_first[coarse ? NB : B] = P;
    __syncthreads();
    // slices of the flattened stage space: [0, grid) static, then dynamic ones (claimed with an
    // atomic) when the launch is long enough to pay for their merges -- see decode_attention.cu
    const int grid = (int)gridDim.x;
    const int fair = (P + grid - 1) / grid;
    const int qs = (fair * G >= min_dyn) ? max(1, fair * 3 / 4) : max(1, fair);
    const int dyn0 = (int)min((long long)P, (long long)grid * qs);
    const int dyn = P - dyn0;
    const int qd = max(1, (dyn + 3 * grid - 1) / (3 * grid));
    const int n_slices = grid + (dyn + qd - 1) / qd;
    auto slice_start = [&](int sl) -> int {
        return sl < grid ? min(P, sl * qs) : min(P, dyn0 + (sl - grid) * qd);
    };
    auto slice_of = [&](int pos) -> int { return pos < dyn0 ? pos / qs : grid + (pos - dyn0) / qd; };
    int slice = 0, g0 = 0, g1 = 0;

    // work iterator, identical in the producer and the consumers; `cur` is a global stage index
    auto next_seg = [&](int& cur, WpSeg& sg) -> bool {
        if (cur >= g1) return false;
        // largest r with stage_first[r] <= cur (empty rows share a start: skipped), by a warp-cooperative
        // 32-ary search: every lane probes one candidate per round (all lanes call next_seg together)
        int lo = 0;
        for (int n = coarse ? NB : B; n > 1;) {
            const int step = (n + 31) >> 5;
            const int idx = lo + lane * step;
            const bool le = (idx < lo + n) && (stage_first[idx] <= cur);
            const int k = 31 - __clz(__ballot_sync(0xffffffffu, le));   // lane 0 always qualifies
            const int end = lo + n;
            lo += k * step;
            n = min(step, end - lo);
        }
        int start, n_st, len;
        if (!coarse) {
            start = stage_first[lo];
            n_st = stage_first[lo + 1] - start;
            len = len_s[lo];
        } else {
            // lo = the block of 32 rows that holds stage `cur`; lane i takes row 32 * lo + i
            const int row = lo * kWpBlk + lane;
            const int Lr = (row < B) ? __ldg(lengths + row) : 0;
            const int n = (Lr + G - 1) / G;
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int first = stage_first[lo] + incl - n;
            const unsigned hit = __ballot_sync(0xffffffffu, n > 0 && first <= cur && cur < first + n);
            const int k = __ffs(hit) - 1;   // exactly one lane: cur < g1 <= P lies inside some row
            start = __shfl_sync(0xffffffffu, first, k);
            n_st = __shfl_sync(0xffffffffu, n, k);
            len = __shfl_sync(0xffffffffu, Lr, k);
            lo = lo * kWpBlk + k;
        }
        const int st1 = min(g1 - start, n_st);
        sg.r = lo;
        sg.start = start;
        sg.p0 = (cur - start) * G;
        sg.p1 = min(st1 * G, len);   // only the row's last stage can be partial
        sg.nseg = slice_of(start + n_st - 1) - slice_of(start) + 1;
        sg.pidx = 2 * slice + (cur == g0 ? 0 : 1);
        cur = start + st1;
        return true;
    };
    int cur = 0;
    WpSeg sg;
    WP_STAMP(2);

    if (warp == kWpProducerWarp) {
        // ===================== producer warp =====================
        uint32_t it = 0;
        uint32_t n_open = 0;   // slices handed to the consumers so far
        const uint64_t kv_policy = l2_policy_evict_first();   // K|V is read once per step
        // the consumers learn which slice comes next through a two-entry box of its own (a ring stage
        // is only visited by the warps that own a position in it, so it cannot carry the news)
        auto publish = [&](int sl) {
            if (lane == 0) {
                const int b = n_open & 1u;
                mbar_wait(&sl_empty[b], ((n_open >> 1) & 1u) ^ 1u);
                slice_box[b] = sl;
                mbar_arrive(&sl_full[b]);
            }
            ++n_open;
        };
        slice = (int)blockIdx.x;
        for (;;) {
            if (slice >= n_slices) break;
            g0 = slice_start(slice);
            g1 = slice_start(slice + 1);
            cur = g0;
            publish(slice);
            // one segment ahead: the page pointers of the next segment load while this one streams
            WpSeg sg_nx;
            bool have = next_seg(cur, sg);
            auto first_pages = [&](const WpSeg& g) -> const float* {
                const int pg = g.p0 / kPage + lane;
                return (pg * kPage < g.p1) ? page_table[(size_t)g.r * W + pg] : nullptr;
            };
            const float* pages_cur = have ? first_pages(sg) : nullptr;
            while (have) {
                const bool have_nx = next_seg(cur, sg_nx);
                const float* pages_nx = have_nx ? first_pages(sg_nx) : nullptr;
                const int r = sg.r, p1 = sg.p1;
                int pgb = sg.p0 / kPage;   // lane i holds the pointer of page pgb + i
                const float* my_page = pages_cur;
                for (int pos = sg.p0; pos < p1; pos += G, ++it) {
                    const int stage = (int)(it & stage_mask);
                    const uint32_t parity = (it >> stage_lg) & 1u;
                    const int nvalid = min(G, p1 - pos);
                    if ((pos + nvalid - 1) / kPage >= pgb + 32) {
                        pgb = pos / kPage;
                        const int pg = pgb + lane;
                        my_page = (pg * kPage < p1) ? page_table[(size_t)r * W + pg] : nullptr;
                    }
                    if (lane == 0) {
                        mbar_wait(&empty_bar[stage], parity ^ 1u);
                        mbar_expect_tx(&full_bar[stage], (uint32_t)nvalid * (uint32_t)rowb);
                        *issued = it + 1u;
                    }
                    __syncwarp();
                    const int j = pos + (lane < G ? lane : 0);
                    const int pg = min(j / kPage - pgb, 31);
                    const float* page = reinterpret_cast<const float*>(
                        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(my_page), pg));
                    if (lane < nvalid) {
                        const float* src = page + (size_t)(j & (kPage - 1)) * pos_floats + d;   // K starts d floats in
                        bulk_g2s_hint(ring + (size_t)stage * stage_bytes + (size_t)lane * rowb, src,
                                      (uint32_t)rowb, &full_bar[stage], kv_policy);
                    }
                }
                sg = sg_nx;
                pages_cur = pages_nx;
                have = have_nx;
            }
            if (n_slices == grid) break;   // no dynamic slices in this launch
            int nxt = 0;
            if (lane == 0) nxt = grid + atomicAdd(&row_done[B], 1);
            slice = __shfl_sync(0xffffffffu, nxt, 0);
        }
        publish(-1);   // no more work
        return;
    }

    if (warp >= kWpConsumerWarps) return;   // spare warps (CW > 1)
    // ===================== consumer warps =====================
    const int cgrp = warp % CW;          // which part of the columns
    const int pgrp = warp / CW;          // which positions
    const int cbase = cgrp * (NCW * 32) + lane;   // float4 column of this lane, + 32 * i
    const float sqrt_d = sqrtf((float)d);
    // empty rows produce zeros (the reference stores result = 0, paged_attention.cu:289,:323)
    for (int r = blockIdx.x; r < B; r += gridDim.x) {
        if (coarse ? (__ldg(lengths + r) > 0) : (stage_first[r + 1] != stage_first[r])) continue;
        for (int col = tid; col < d4; col += kWpConsumerThreads)
            reinterpret_cast<float4*>(out + (size_t)r * d)[col] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t it = 0;
    int n_seg_dbg = 0;
    uint32_t xpar = 0;   // CW > 1: parity of the score exchange buffer (per position of this group)

    // Partial rows are merged by whoever completes a row's last segment, in slice order; rows are
    // queued and flushed after the CTA has run out of slices (or when the queue is full).
    int n_pend = 0;
    auto flush_pending = [&]() {
        if (n_pend == 0) return;
        // release: the barrier orders every consumer thread's partial stores before the gpu-scope
        // acq_rel atomic of the arriving thread (cumulativity); acquire: the same atomic, then the
        // barrier, then L1-bypassing loads (__ldcg) by all threads
        named_bar_sync(1, kWpConsumerThreads);
        if (tid < n_pend)
            pend_flag[tid] = (atom_add_acq_rel_gpu(&row_done[pend_r[tid]], 1) == pend_nseg[tid] - 1) ? 1 : 0;
        named_bar_sync(1, kWpConsumerThreads);
        for (int pi = 0; pi < n_pend; ++pi) {
            if (!pend_flag[pi]) continue;
            const int r = pend_r[pi], nseg = pend_nseg[pi];
            if (dbg != nullptr && tid == 0) dbg[(size_t)blockIdx.x * 16 + 11] += ((long long)nseg << 32) | 1;
            const int start = pend_start[pi];
            const int b_first = slice_of(start);
            // segment k of the row lives in slice b_first + k: its head slot, except that the row's
            // first segment is its slice's tail slot unless the row opens that slice
            auto slot_of = [&](int k) -> size_t {
                return (size_t)2 * (b_first + k) + ((k == 0 && start != slice_start(b_first)) ? 1 : 0);
            };
            // (m, l) of up to 32 segments at a time, one per lane
            float M = -INFINITY;
            for (int k0 = 0; k0 < nseg; k0 += 32) {
                const int k = k0 + lane;
                float m = (k < nseg) ? __ldcg(part_ml + 2 * slot_of(k)) : -INFINITY;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                M = fmaxf(M, m);
            }
            for (int c0 = 0; c0 < d4; c0 += kWpConsumerThreads) {   // uniform trip count: shuffles inside
                const int col = c0 + tid;
                float Lsum = 0.f;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
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
                    for (int kk = 0; kk < kn; kk += 4) {
                        float4 pv[4];
                        float w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            w[u] = __shfl_sync(0xffffffffu, wl, min(kk + u, 31));
                            const size_t sl = slot_of(min(k0 + kk + u, nseg - 1));
                            pv[u] = (col < d4 && kk + u < kn)
                                        ? __ldcg(reinterpret_cast<const float4*>(part_acc + sl * d) + col)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (kk + u < kn) {
                                a.x = fmaf(w[u], pv[u].x, a.x); a.y = fmaf(w[u], pv[u].y, a.y);
                                a.z = fmaf(w[u], pv[u].z, a.z); a.w = fmaf(w[u], pv[u].w, a.w);
                            }
                        }
                    }
                }
                const float norm = 1.f / Lsum;
                if (col < d4)
                    reinterpret_cast<float4*>(out + (size_t)r * d)[col] =
                        make_float4(a.x * norm, a.y * norm, a.z * norm, a.w * norm);
            }
            if (tid == 0) row_done[r] = 0;   // ready for the next launch
        }
        named_bar_sync(1, kWpConsumerThreads);
        n_pend = 0;
    };

    WP_STAMP(3);
    uint32_t n_open = 0;
    for (;;) {
        // which slice comes next (or none)
        {
            const int b = n_open & 1u;
            mbar_wait(&sl_full[b], (n_open >> 1) & 1u);
            slice = slice_box[b];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sl_empty[b]);
            ++n_open;
            if (slice < 0) break;
            g0 = slice_start(slice);
            g1 = slice_start(slice + 1);
            cur = g0;
            if (n_pend + 2 > kWpMaxPend) flush_pending();
        }
        // The q row of a segment is fetched while the previous segment is being merged, so a row
        // boundary costs one blocking barrier (deposit -> merge) and no global-memory round trip.
        float4 qv[NCW];
        bool have = next_seg(cur, sg);
        if (have) {
#pragma unroll
            for (int i = 0; i < NCW; ++i)
                qv[i] = reinterpret_cast<const float4*>(q + (size_t)sg.r * d)[cbase + 32 * i];
        }
        while (have) {
            const int r = sg.r, p0 = sg.p0, p1 = sg.p1, row_start = sg.start;
            float4 acc[NCW];
#pragma unroll
            for (int i = 0; i < NCW; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            float m_run = -INFINITY, l_run = 0.f;
            // what comes after this segment: looked up now, off the critical path of the row boundary
            WpSeg sg_nx;
            const bool have_nx = next_seg(cur, sg_nx);

            // positions are dealt round-robin to the PW groups, counted from the segment start; a warp
            // only visits the ring stages that hold one of its positions
            const int n_pos = p1 - p0;
            for (int rel = pgrp; rel < n_pos; rel += PW) {
                const uint32_t st = it + (uint32_t)(rel >> g_lg);
                const int g = rel & (G - 1);
                const int stage = (int)(st & stage_mask);
                const unsigned char* krow = ring + (size_t)stage * stage_bytes + (size_t)g * rowb;
                // (all 32 lanes wait: one poller per warp + __syncwarp was measured -- it does lower the power draw,
                // SM clock 1642 -> 1702 MHz under the cap, but the launch ran at 6.43 instead of 7.12 TB/s)
                wp_wait_issued(issued, st);
                mbar_wait(&full_bar[stage], (st >> stage_lg) & 1u);
                if (st == 0) WP_STAMP(4);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int i = 0; i < NCW; ++i) {
                    const float4 k = wp_ld4<KVB>(krow, cbase + 32 * i);
                    s0 = fmaf(qv[i].x, k.x, s0);
                    s1 = fmaf(qv[i].y, k.y, s1);
                    s2 = fmaf(qv[i].z, k.z, s2);
                    s3 = fmaf(qv[i].w, k.w, s3);
                }
                float s = warp_sum((s0 + s1) + (s2 + s3));
                if constexpr (CW > 1) {
                    // the CW warps of the group add their column parts in a fixed order; the
                    // exchange buffer is double-buffered, so one barrier per position is enough
                    float* x = xch + ((size_t)pgrp * 2 + (xpar & 1u)) * CW;
                    if (lane == 0) x[cgrp] = s;
                    named_bar_sync(2 + pgrp, 32 * CW);
                    s = x[0];
#pragma unroll
                    for (int c = 1; c < CW; ++c) s += x[c];
                    ++xpar;
                }
                const float sc = s / sqrt_d;
                if (sc > m_run) {   // warp-uniform
                    const float corr = expf(m_run - sc);   // exp(-inf) = 0 on the first position
                    l_run *= corr;
#pragma unroll
                    for (int i = 0; i < NCW; ++i) {
                        acc[i].x *= corr; acc[i].y *= corr; acc[i].z *= corr; acc[i].w *= corr;
                    }
                    m_run = sc;
                }
                const float p = expf(sc - m_run);
                l_run += p;
                const unsigned char* vrow = krow + voff;
#pragma unroll
                for (int i = 0; i < NCW; ++i) {
                    const float4 v = wp_ld4<KVB>(vrow, cbase + 32 * i);
                    acc[i].x = fmaf(p, v.x, acc[i].x);
                    acc[i].y = fmaf(p, v.y, acc[i].y);
                    acc[i].z = fmaf(p, v.z, acc[i].z);
                    acc[i].w = fmaf(p, v.w, acc[i].w);
                }
                __syncwarp();
                // release: one arrival per position; whoever holds the last position of a partly
                // filled stage also arrives for the positions that are not there
                if (lane == 0) {
                    const int nvalid = min(G, n_pos - (rel - g));
                    mbar_arrive_cnt(&empty_bar[stage], (g == nvalid - 1) ? (uint32_t)(1 + G - nvalid) : 1u);
                }
            }
            it += (uint32_t)((n_pos + G - 1) >> g_lg);

            // ---- segment epilogue: merge the PW group states, then final row or partial row ----
            ++n_seg_dbg;
            WP_STAMP(5);
            WP_GT(9);
            // everybody is done merging the previous segment (long ago, as a rule): scratch is free
            named_bar_sync(1, kWpConsumerThreads);
            {
                float* sc = scratch + (size_t)pgrp * d;
#pragma unroll
                for (int i = 0; i < NCW; ++i) reinterpret_cast<float4*>(sc)[cbase + 32 * i] = acc[i];
                if (cgrp == 0 && lane == 0) {
                    sml[2 * pgrp] = m_run;
                    sml[2 * pgrp + 1] = l_run;
                }
            }
            named_bar_sync(1, kWpConsumerThreads);
            // next segment's q: in flight during the merge
            if (have_nx) {
#pragma unroll
                for (int i = 0; i < NCW; ++i)
                    qv[i] = reinterpret_cast<const float4*>(q + (size_t)sg_nx.r * d)[cbase + 32 * i];
            }
            {
                // lane p weighs group p (one expf per lane instead of PW per thread)
                const float2 ml = (lane < PW) ? reinterpret_cast<const float2*>(sml)[lane]
                                              : make_float2(-INFINITY, 0.f);
                float M = ml.x;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
                const float wl = expf(ml.x - M);   // groups that got no position: exp(-inf) = 0
                float w[PW];
                float Lsum = 0.f;
#pragma unroll
                for (int p = 0; p < PW; ++p) {
                    w[p] = __shfl_sync(0xffffffffu, wl, p);
                    Lsum += __shfl_sync(0xffffffffu, ml.y, p) * w[p];   // fixed order: same in every thread
                }
                const int nseg = sg.nseg;
                const size_t pidx = (size_t)sg.pidx;
                const float norm = (nseg == 1) ? 1.f / Lsum : 1.f;
                float4* dst = (nseg == 1) ? reinterpret_cast<float4*>(out + (size_t)r * d)
                                          : reinterpret_cast<float4*>(part_acc + pidx * d);
                for (int col = tid; col < d4; col += kWpConsumerThreads) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < PW; ++p) {
                        const float4 v = reinterpret_cast<const float4*>(scratch + (size_t)p * d)[col];
                        a.x = fmaf(w[p], v.x, a.x); a.y = fmaf(w[p], v.y, a.y);
                        a.z = fmaf(w[p], v.z, a.z); a.w = fmaf(w[p], v.w, a.w);
                    }
                    dst[col] = make_float4(a.x * norm, a.y * norm, a.z * norm, a.w * norm);
                }
                if (nseg > 1) {
                    if (tid == 0) {
                        part_ml[2 * pidx] = M;
                        part_ml[2 * pidx + 1] = Lsum;
                        pend_r[n_pend] = r;
                        pend_nseg[n_pend] = nseg;
                        pend_start[n_pend] = row_start;
                    }
                    ++n_pend;
                }
            }
            sg = sg_nx;
            have = have_nx;
        }
    }   // slices
    GRIDDEP_TRIGGER_LATE();
    flush_pending();
    // the last CTA to finish re-arms the slice counter for the next launch
    if (tid == 0 && n_slices != grid) {
        __threadfence();
        if (atomicAdd(&row_done[B + 1], 1) == grid - 1) {
            row_done[B] = 0;
            row_done[B + 1] = 0;
        }
    }
    WP_STAMP(6);
    WP_GT(10);
    if (dbg != nullptr && threadIdx.x == 0) dbg[(size_t)blockIdx.x * 16 + 7] = ((long long)n_seg_dbg << 32) | it;
#undef WP_STAMP
#undef WP_GT
}

size_t wp_smem_bytes(int B, int d, int CW, int G, int nstage, bool kvb) {
    const int PW = wp_consumer_warps(CW) / CW;
    const size_t rowb = kvb ? 4 * (size_t)d : 8 * (size_t)d;
    return (size_t)nstage * G * rowb + sizeof(float) * ((size_t)PW * d + 2 * PW + 2 * PW * CW) +
           (2 * kWpMaxStages + 4) * sizeof(uint64_t) + sizeof(int) * (kWpCtrlInts + wp_table_ints(B)) + 128;
}

template <int NCW, int CW, bool KVB>
int wp_launch(mli_ctx* ctx, const float* q, float* const* page_table, const int* lengths, float* out,
              float* part_acc, float* part_ml, int* row_done, int B, int S, int d, int G, int nstage,
              int grid, int min_dyn) {
    auto kern = decode_attention_wp_kernel<NCW, CW, KVB>;
    const size_t smem = wp_smem_bytes(B, d, CW, G, nstage, KVB);
    { int rc0 = ensure_dyn_smem(ctx, kern, smem); if (rc0) return rc0; }
    if (ctx->attn_ev_start) MLI_CUDA(cudaEventRecord(ctx->attn_ev_start, ctx->stream));
    int rc = launch_kernel(ctx, kern, dim3(grid), dim3(kWpThreads), smem, q, page_table, lengths, out,
                           part_acc, part_ml, row_done, B, S, d, G, nstage, min_dyn,
                           reinterpret_cast<long long*>(ctx->tc_dbg), ctx->trace);
    if (rc) return rc;
    if (ctx->attn_ev_stop) MLI_CUDA(cudaEventRecord(ctx->attn_ev_stop, ctx->stream));
    return 0;
}

}  // namespace

// emb_dim values the warp-per-position kernel is instantiated for
bool attention_wp_supported(int d) {
    return d == 128 || d == 256 || d == 512 || d == 1024 || d == 2048 || d == 4096;
}

int attention_wp_grid(mli_ctx* ctx) { return ctx->num_sms; }

namespace {
// ring geometry for a launch: positions per stage and stages (both powers of two); false = the
// shared memory left after the per-row prefix does not hold two stages
bool wp_plan(int B, int d, bool kvb, int* G_out, int* nstage_out) {
    const int CW = d <= 1024 ? 1 : d / 1024;
    const int PW = wp_consumer_warps(CW) / CW;
    const size_t rowb = kvb ? 4 * (size_t)d : 8 * (size_t)d;
    // pipeline stages of ~16 KB: fine-grained enough that a row's last, partly filled stage wastes
    // little of the ring, large enough that the per-stage barrier traffic is noise
    int G = (int)std::max<size_t>(1, (16 * 1024) / rowb);
    if (G > 32) G = 32;
    const size_t fixed = sizeof(float) * ((size_t)PW * d + 2 * PW + 2 * PW * CW) +
                         (2 * kWpMaxStages + 4) * sizeof(uint64_t) + sizeof(int) * (kWpCtrlInts + wp_table_ints(B)) + 128;
    if (fixed >= 226 * 1024) return false;
    const size_t budget = 226 * 1024 - fixed;
    int nstage = kWpMaxStages;   // a power of two: the kernel masks instead of dividing
    while (nstage >= 2 && (size_t)nstage * G * rowb > budget) nstage >>= 1;
    *G_out = G;
    *nstage_out = nstage;
    return nstage >= 2;
}
}  // namespace

// does the kernel cover this launch (instantiated emb_dim, room for the ring)?
bool attention_wp_usable(mli_ctx* ctx, int B, int d) {
    int G, nstage;
    return attention_wp_supported(d) && wp_plan(B, d, ctx->kv_bf16 != 0, &G, &nstage);
}

// part_acc / part_ml: two partial slots per slice (<= 5 * grid slices); row_done: [B + 2] zeroed counters
int launch_decode_attention_wp(mli_ctx* ctx, const float* q, float* const* page_table, const int* lengths,
                               float* out, float* part_acc, float* part_ml, int* row_done, int B, int S,
                               int d, int min_dyn) {
    const bool kvb = ctx->kv_bf16 != 0;
    int G = 1, nstage = 0;
    if (!attention_wp_supported(d) || !wp_plan(B, d, kvb, &G, &nstage)) {
        set_error("decode attention (warp per position): shape not covered (emb_dim / shared memory)");
        return MLI_ERR_UNSUPPORTED;
    }
    const int grid = attention_wp_grid(ctx);
#define MLI_WP_CASE(D_, NCW_, CW_)                                                                          \
    if (d == D_)                                                                                            \
        return kvb ? wp_launch<NCW_, CW_, true>(ctx, q, page_table, lengths, out, part_acc, part_ml,        \
                                                row_done, B, S, d, G, nstage, grid, min_dyn)                \
                   : wp_launch<NCW_, CW_, false>(ctx, q, page_table, lengths, out, part_acc, part_ml,       \
                                                 row_done, B, S, d, G, nstage, grid, min_dyn);
    MLI_WP_CASE(128, 1, 1)
    MLI_WP_CASE(256, 2, 1)
    MLI_WP_CASE(512, 4, 1)
    MLI_WP_CASE(1024, 8, 1)
    MLI_WP_CASE(2048, 8, 2)
    MLI_WP_CASE(4096, 8, 4)
#undef MLI_WP_CASE
    set_error("decode attention (warp per position): no instantiation for this emb_dim");
    return MLI_ERR_UNSUPPORTED;
}

}  // namespace mli
