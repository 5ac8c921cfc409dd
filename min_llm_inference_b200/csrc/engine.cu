// engine.cu -- on-device continuous-batching engine (mli_engine_*).
//
// Replaces the host loop start_paged_attention_inference_engine (src/inferencer.cpp:43-85) and the
// host scheduler it drives:
//   process_decoder_result                    src/item_storage.cpp:97-139
//   allocate_or_free_memory_blocks_if_needed  src/paged_item_storage.cpp:14-60
//   paged insert_new_items                    src/paged_item_storage.cpp:62-122
//   MemoryBlockManager / PagedAttentionsManager  src/paged_item_storage.cpp:125-203
// The request table (prompt + generated tokens), the request deque, the page free list, the page
// table, the used-row list and every admission / retirement / growth / pre-emption decision live
// in HBM and are advanced by ONE scheduler kernel per iteration; the model kernels that follow read
// their work (lengths, new rows, tile list) from device memory.  A whole iteration is captured in a
// CUDA graph; the host launches graphs a few steps ahead and polls a mapped, pinned "done" word --
// the reference's per-step cudaMemcpy round trips (1 D2H + up to 4 H2D) are gone.
//
// Decision parity with the reference (what decides tokens): admission scans rows in index order and
// takes the queue head while free >= max(4, ceil((len+R)/16)) (Q6); finished rows are retired in
// row order; growth walks the used list in admission order, one page per row, and pre-empts the
// list TAIL back to the FRONT of the queue when the pool is empty (Q3); compat_stale_lengths = 1
// additionally replays quirk Q1 (every in-flight row's device length snaps back to its admission
// length whenever any row is unoccupied).  What is NOT replayed is the physical order of the free
// list (pages are returned in table order), which only changes which page a row gets, never a token.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

namespace mli {

struct SchedVars {
    int q_head, q_count, q_cap;
    int f_head, f_count;
    int n_used, n_fin, n_new;
    int iter, done, error, n_req;   // n_req = requests the scheduler has taken over from the inbox so far
    long long steps, generated, preemptions, admitted;
    int max_used, min_free;         // peak resident rows / fewest free pages seen (reported, never read back)
    int poll;                       // set by the host once mli_engine_poll_finished is in use (never written here)
};

struct SchedArgs {
    SchedVars* v;
    int* req_tok;     // [max_req][S]
    int* req_cnt;     // [max_req]
    int* req_plen;    // [max_req] prompt length at submission (for max_new_tokens)
    int* queue;       // [q_cap] ring of request ids
    int* fin_ids;     // [max_req]
    int* row_req;     // [B]
    int* lengths;     // [B] device lengths (what the kernels read)
    int* len_shadow;  // [B] the reference's lengths_host (admission length or 0)
    int* used;        // [B] used-row list in admission order
    int* npages;      // [B]
    float** page_table;  // [B][W]
    float** free_ring;   // [n_blocks]
    int* dec;         // [B][R]
    int* new_idx;     // [B]
    // work lists for the model kernels of this step (emitted at the end of the scheduler)
    int* act_rows;    // [B] rows with length > 0, in row order
    TileDesc* gran;   // [max_gran] 16-position granules covering [0, L) of every new row
    int* counts;      // [0] = number of active rows, [1] = number of granules
    int max_gran;
    volatile int* done_host;  // mapped pinned
    volatile int* n_avail;    // device: requests whose table rows are complete (written by the ingest stream);
                              // ids [v->n_req, *n_avail) are waiting to be queued
    unsigned long long* trace;  // optional step timeline
    int B, S, W, R, n_blocks, compat;
    int max_new;      // > 0: a request is finished once it has generated this many tokens (opt-in)
    int max_prefill;  // > 0: admission throttle, prompt positions admitted per step (opt-in)
    int chunk;        // > 0: chunked prefill (opt-in): prompt positions prefilled per step, a multiple
                      // of 16; an admitted row stays inactive (length 0) until its last chunk is scheduled
    int* pf_pos;      // [B] chunked prefill: positions of the row's prompt already scheduled, -1 = not prefilling
};

constexpr int kSchedThreads = 1024;
constexpr int kGran = kPage;   // positions per prefill granule

// exclusive block scan; every thread must call it.  total = sum over the block.  s_warp is
// [2][32]: consecutive calls alternate between the halves, so a call needs two barriers, not three
// (the half a call writes was last read two barriers ago).
__device__ int block_scan_excl(int v, int* total, int* s_warp2, int& phase) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    int* s_warp = s_warp2 + 32 * (phase & 1);
    ++phase;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    // every warp scans the (at most 32) warp totals itself: no second hand-off through shared memory
    int w = (lane < nwarps) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
    }
    const int before = (warp > 0) ? __shfl_sync(0xffffffffu, w, warp - 1) : 0;
    *total = __shfl_sync(0xffffffffu, w, nwarps - 1);
    return x - v + before;
}

// shared-memory footprint of the scheduler: six int arrays of B entries and the free-page window
constexpr int kMaxFreeWindow = 2048;   // pages one growth phase can take from shared memory
__host__ __device__ inline int sched_window(int B) { return B < kMaxFreeWindow ? B : kMaxFreeWindow; }
size_t sched_smem_bytes(int B) {
    return (size_t)B * (6 * sizeof(int)) + (size_t)sched_window(B) * sizeof(float*) + 64;
}

// One CTA advances the whole continuous-batching state by one iteration.  Per-row state (request
// of the row, pages of the row, the used-row list) is mirrored in shared memory for the duration of
// the kernel, and the pages the growth phase hands out come from a shared-memory window of the free
// ring, so the inherently ordered part of the reference's algorithm (grow in admission order,
// pre-empt the list tail when the pool is dry) runs without dependent global-memory round trips.
__global__ void __launch_bounds__(kSchedThreads) sched_step_kernel(SchedArgs a) {
    extern __shared__ __align__(16) unsigned char sched_smem[];
    __shared__ int s_warp[64];
    __shared__ int s_carry[4];
    int scan_phase = 0;
    __shared__ long long s_pre;
    const int tid = threadIdx.x, T = blockDim.x;
    const int B = a.B, S = a.S, W = a.W, R = a.R;
    const int n_fq = sched_window(B);
    float** fq = reinterpret_cast<float**>(sched_smem);          // [n_fq] window of the free ring
    int* s_req = reinterpret_cast<int*>(fq + n_fq);               // [B] row -> request (-1 = none)
    int* s_np = s_req + B;                                        // [B] pages of the row
    int* s_used = s_np + B;                                       // [B] used-row list
    int* s_flag = s_used + B;                                     // [B] bit0 finished, bit1 unoccupied, bits 2.. token count; later: occupied
    int* s_list = s_flag + B;                                     // [B] need list / free rows
    int* s_len = s_list + B;                                      // [B] device lengths as the model kernels will see them
    // optional phase stamps (tools/sched_timing.py): trace[2] != 0 -> 16 %globaltimer slots per step after
    // the step table
    unsigned long long* ph = nullptr;
    if (tid == 0 && a.trace != nullptr && a.trace[2] != 0 && a.trace[0] < a.trace[1])
        ph = a.trace + 8 + 8 * a.trace[1] + 16 * a.trace[0];
#define SCHED_PH(k) do { if (ph != nullptr) ph[k] = globaltimer_ns(); } while (0)
    SCHED_PH(0);
    // every thread reads the counters itself (one broadcast line) together with its rows' state:
    // one global round trip, no shared-memory hand-off.  The counters, the row -> request map, the
    // page counts and the used list are written by this kernel only (the previous iteration's
    // instance completed long ago), so they are fetched BEFORE the dependency wait; what the decoder
    // produces (lengths, tokens) is read after it.
    SchedVars sv = *a.v;
    for (int r = tid; r < B; r += T) {
        s_req[r] = a.row_req[r];
        s_np[r] = a.npages[r];
        s_flag[r] = 0;
    }
    for (int i = tid; i < sv.n_used; i += T) s_used[i] = a.used[i];
    // likewise the head of the free ring (window entry j = ring[f_head + j]; pages freed later in
    // this step are forwarded into the window by the code that frees them) and the head of the
    // queue with the request lengths: what the admission phase would otherwise fetch through two
    // dependent round trips at the very end of the kernel
    const int F0 = sv.f_count;
    for (int j = tid; j < min(F0, n_fq); j += T) fq[j] = a.free_ring[(sv.f_head + j) % a.n_blocks];
    int pq_id = -1, pq_len = 0;
    if (tid < sv.q_count) {
        pq_id = a.queue[(sv.q_head + tid) % sv.q_cap];
        pq_len = a.req_cnt[pq_id];
    }
    // requests appended by mli_engine_submit / mli_engine_enqueue (the ingest stream publishes the count
    // after their table rows are complete): ids [sv.n_req, n_avail) are queued in phase 4.  The count is
    // not produced by the predecessor kernel (any value read is a valid, monotonic snapshot), so it is
    // fetched with the other early loads
    const int n_avail = *a.n_avail;
    SCHED_PH(1);
    griddep_wait();
    SCHED_PH(2);
    GRIDDEP_TRIGGER_EARLY();
    trace_stamp(a.trace, 0);
    if (sv.done && n_avail == sv.n_req) {
        if (tid == 0) {
            a.v->n_new = 0;
            a.counts[0] = 0;
            a.counts[1] = 0;
            *a.done_host = 1;   // (a late arrival may have cleared it; idle again)
        }
        return;
    }
    const bool first = (sv.iter == 0);
    for (int r = tid; r < B; r += T) s_len[r] = a.lengths[r];
    __syncthreads();
    SCHED_PH(3);
    int n_used = sv.n_used;
    int F = sv.f_count, fh = sv.f_head, qh = sv.q_head, qc = sv.q_count;
    const int q_cap = sv.q_cap, nb = a.n_blocks;

    int n_fin_now = sv.n_fin;   // finished requests after this step
    if (!first) {
        // ================= phase 1: process_decoder_result (item_storage.cpp:97-139) =================
        int local_gen = 0, local_err = 0;
        for (int r = tid; r < B; r += T) {
            bool empty = false, finished = false;
            const int id = s_req[r];
            int c = (id >= 0) ? a.req_cnt[id] : 0;
            // opt-in cap on generated tokens: the request finishes when count - prompt length reaches it
            const int c_cap = (id >= 0 && a.max_new > 0) ? a.req_plen[id] + a.max_new : 0x7fffffff;
            for (int j = 0; j < R; ++j) {
                const int t = a.dec[(size_t)r * R + j];
                if (t == MLI_EMPTY_ROW_TOKEN_ID) {
                    // a row whose prompt is still being prefilled in chunks has no token yet and is NOT free
                    if (a.chunk > 0 && id >= 0 && a.pf_pos[r] >= 0) break;
                    empty = true;
                } else if (id < 0) {
                    local_err = 1;  // token for a row that is not processing
                    empty = true;
                } else {
                    if (c < S) a.req_tok[(size_t)id * S + c] = t;
                    c += 1;
                    ++local_gen;
                    if (c >= S || t == MLI_EOF_TOKEN_ID || c >= c_cap) finished = true;
                }
                if (finished || empty) break;
            }
            if (id >= 0) {
                c = (finished && c > S) ? S : c;
                a.req_cnt[id] = c;
            }
            s_flag[r] = (finished ? 1 : 0) | ((finished || empty) ? 2 : 0) | (c << 2);
        }
        if (local_gen) atomicAdd(reinterpret_cast<unsigned long long*>(&a.v->generated),
                                 (unsigned long long)local_gen);
        if (local_err) a.v->error = 1;
        __syncthreads();
        // finished requests are appended in row order (:118-131)
        int n_fin = sv.n_fin;
        for (int base = 0; base < B; base += T) {
            const int r = base + tid;
            const int f = (r < B) ? (s_flag[r] & 1) : 0;
            int tot;
            const int pos = block_scan_excl(f, &tot, s_warp, scan_phase);
            if (f) {
                a.fin_ids[n_fin + pos] = s_req[r];
                s_req[r] = -1;
            }
            n_fin += tot;
        }
        n_fin_now = n_fin;   // published at the very end of the kernel, after the lists it counts

        SCHED_PH(4);
        // ================= phase 2: free rows in finished_indices (paged_item_storage.cpp:20-32) =====
        {
            int kept = 0, freed = 0;
            for (int base = 0; base < n_used; base += T) {
                const int i = base + tid;
                int row = -1, rel = 0, keep = 0, np = 0;
                if (i < n_used) {
                    row = s_used[i];
                    rel = (s_flag[row] >> 1) & 1;
                    keep = !rel;
                    np = rel ? s_np[row] : 0;
                }
                int totp, totk;
                const int ppos = block_scan_excl(np, &totp, s_warp, scan_phase);
                const int kpos = block_scan_excl(keep, &totk, s_warp, scan_phase);
                // every read of s_used[] in this chunk is done (the scans contain barriers)
                if (rel) {
                    const int wj = F + freed + ppos;   // window index (the ring head does not move here)
                    for (int t = 0; t < np; ++t) {
                        float* pg = a.page_table[(size_t)row * W + t];
                        a.free_ring[(fh + wj + t) % nb] = pg;
                        if (wj + t < n_fq) fq[wj + t] = pg;
                    }
                    s_np[row] = 0;
                }
                if (keep) s_used[kept + kpos] = row;
                freed += totp;
                kept += totk;
                __syncthreads();
            }
            F += freed;
            n_used = kept;
        }
        __syncthreads();   // ring writes of phase 2 are read back below

        SCHED_PH(5);
        // ================= phase 3: grow / pre-empt (paged_item_storage.cpp:36-59) ===================
        {
            int m = 0;
            for (int base = 0; base < n_used; base += T) {
                const int i = base + tid;
                int need = 0;
                if (i < n_used) {
                    const int row = s_used[i];
                    need = ((s_flag[row] >> 2) + R > s_np[row] * kPage) ? 1 : 0;   // count from phase 1
                }
                int tot;
                const int pos = block_scan_excl(need, &tot, s_warp, scan_phase);
                if (need) s_list[m + pos] = i;
                m += tot;
            }
            // window of the free ring the growth loop may consume: entry j = ring[fh + j]
            const int fh0 = fh;   // entry j of the window is ring[fh0 + j]
            // (its first min(F, n_fq) entries are in shared memory already: fetched before the
            // dependency wait, or forwarded by phase 2)
            __syncthreads();
            if (F >= m) {
                // every row that needs a page finds one: no pre-emption can happen and the pages are
                // handed out by a scan (row k of the need list takes the k-th free page it is due)
                int taken = 0;
                for (int base = 0; base < m; base += T) {
                    const int k = base + tid;
                    int row = -1, np = 0, take = 0;
                    if (k < m) {
                        row = s_used[s_list[k]];
                        np = s_np[row];
                        take = (np < W) ? 1 : 0;   // allocate_memory_block (:196-203): new page at index size-1
                    }
                    int tot;
                    const int idx = taken + block_scan_excl(take, &tot, s_warp, scan_phase);
                    if (take) {
                        a.page_table[(size_t)row * W + np] =
                            (idx < n_fq) ? fq[idx] : a.free_ring[(fh0 + idx) % nb];
                        s_np[row] = np + 1;
                    }
                    taken += tot;
                }
                F -= taken;
                fh = (fh + taken) % nb;
                __syncthreads();
            } else {
            if (tid == 0) {
                int n = n_used, taken = 0;   // taken = window entries consumed so far
                long long pre = 0;
                auto preempt = [&](int row) {
                    // move_to_new: front of the queue, keeping generated tokens (item_storage.cpp:75-79)
                    qh = (qh - 1 + q_cap) % q_cap;
                    a.queue[qh] = s_req[row];   // (the queue count is adjusted by every thread below)
                    s_req[row] = -1;
                    if (a.chunk > 0) a.pf_pos[row] = -1;   // a pre-empted prompt starts over when it is re-admitted
                    const int np = s_np[row];
                    for (int t = 0; t < np; ++t) {
                        float* pg = a.page_table[(size_t)row * W + t];
                        a.free_ring[(fh + F + t) % nb] = pg;
                        if (taken + F + t < n_fq) fq[taken + F + t] = pg;   // only reachable when F <= m
                    }
                    F += np;
                    s_np[row] = 0;
                    ++pre;
                };
                for (int k = 0; k < m; ++k) {
                    const int i = s_list[k];
                    if (i >= n) break;  // already pre-empted as a tail
                    const int row = s_used[i];
                    for (;;) {
                        if (F > 0) {
                            // allocate_memory_block (:196-203): new page at table index size-1
                            const int np = s_np[row];
                            if (np < W) {
                                // beyond the window (more than kMaxFreeWindow growths in one step) the
                                // page comes from the ring itself: everything written to it so far
                                // was written by this thread
                                a.page_table[(size_t)row * W + np] =
                                    (taken < n_fq) ? fq[taken] : a.free_ring[(fh0 + taken) % nb];
                                ++taken;
                                fh = (fh + 1) % nb;
                                --F;
                                s_np[row] = np + 1;
                            }
                            break;
                        } else if (i == n - 1) {
                            preempt(row);
                            --n;
                            break;
                        } else {
                            preempt(s_used[n - 1]);
                            --n;
                        }
                    }
                }
                s_carry[0] = n;
                s_carry[1] = F;
                s_carry[2] = fh;
                s_carry[3] = qh;
                s_pre = pre;
            }
            __syncthreads();
            qc += n_used - s_carry[0];   // every row dropped from the used list went to the queue
            n_used = s_carry[0];
            F = s_carry[1];
            fh = s_carry[2];
            qh = s_carry[3];
            sv.preemptions += s_pre;
            __syncthreads();
            }
        }
    }

    SCHED_PH(6);
    // ================= phase 4: insert_new_items (paged_item_storage.cpp:62-122) =====================
    int k_adm = 0;
    {
        // newly arrived requests join the tail of the queue in id order (ItemStorage::add_new_item,
        // item_storage.cpp:190-192)
        const int n_arrived = n_avail - sv.n_req;
        for (int k = tid; k < n_arrived; k += T) a.queue[(qh + qc + k) % q_cap] = sv.n_req + k;
        qc += n_arrived;
        for (int r = tid; r < B; r += T) s_flag[r] = 0;
        __syncthreads();
        for (int i = tid; i < n_used; i += T) s_flag[s_used[i]] = 1;
        __syncthreads();
        // unoccupied rows in index order
        int n_free_rows = 0;
        for (int base = 0; base < B; base += T) {
            const int r = base + tid;
            const int fr = (r < B && !s_flag[r]) ? 1 : 0;
            int tot;
            const int pos = block_scan_excl(fr, &tot, s_warp, scan_phase);
            if (fr) s_list[n_free_rows + pos] = r;
            n_free_rows += tot;
        }
        // an admission takes at least MLI_DEFAULT_INIT_NUM_BLOCKS pages (:86-88): with fewer free pages
        // nobody gets in and the candidate scans (four block barriers) are skipped
        const int n_cand = (F >= MLI_DEFAULT_INIT_NUM_BLOCKS) ? min(n_free_rows, qc) : 0;
        const int w_used = (fh - sv.f_head + nb) % nb;   // window entries the growth phase consumed
        // Candidate j (the j-th unoccupied row) takes queue item j.  The reference tests the queue head
        // against the pages that are free at that moment (:84-88) and stops admitting at the first
        // head that does not fit (every later row sees the same head and the same pool), so the
        // admitted candidates are a PREFIX: j is admitted iff every candidate up to j passes
        //   free pages left after the earlier admissions  >=  need_j = max(4, ceil((len_j + R) / 16)).
        // A row stores (and the pool gives up) only take_j = min(need_j, W) pages: the reference pops
        // need_j and cannot represent more than W in the table (the oracle returns the surplus).
        // Opt-in throttle (max_prefill > 0): a step admits prompts while their positions add up to at
        // most max_prefill; the first candidate of a step always passes.
        if (tid == 0) {
            s_carry[0] = 0;            // pages consumed by the admitted prefix
            s_carry[1] = 0x7fffffff;   // first candidate that fails
        }
        __syncthreads();
        int take_before = 0, len_before = 0;
        for (int base = 0; base < n_cand; base += T) {
            const int j = base + tid;
            int id = -1, len = 0, need = 0, take = 0;
            if (j < n_cand) {
                if (base == 0 && qh == sv.q_head && tid < sv.q_count) {
                    id = pq_id;   // nothing was pushed to the front of the queue in this step
                    len = pq_len;
                } else {
                    id = a.queue[(qh + j) % q_cap];
                    len = a.req_cnt[id];
                }
                need = max((len + R + kPage - 1) / kPage, MLI_DEFAULT_INIT_NUM_BLOCKS);
                take = min(need, W);
            }
            int tott, totl = 0;
            const int before = take_before + block_scan_excl(take, &tott, s_warp, scan_phase);
            // (the second scan only when the throttle is on: every block scan is two barriers)
            const int lbefore = (a.max_prefill > 0)
                                    ? len_before + block_scan_excl(len, &totl, s_warp, scan_phase) : 0;
            const bool ok = before + need <= F &&
                            (a.max_prefill <= 0 || j == 0 || lbefore + len <= a.max_prefill);
            if (j < n_cand && !ok) atomicMin(&s_carry[1], j);
            __syncthreads();
            const int first_fail = s_carry[1];
            if (j < n_cand && j < first_fail) {
                const int row = s_list[j];
                for (int t = 0; t < take; ++t) {
                    const int wi = w_used + before + t;   // window entry = ring[f_head at entry + wi]
                    a.page_table[(size_t)row * W + t] =
                        (wi < n_fq) ? fq[wi] : a.free_ring[(fh + before + t) % nb];
                }
                s_np[row] = take;
                if (a.chunk > 0) {
                    // chunked prefill: the row becomes active (length = prompt length) in the step that
                    // schedules its last chunk (below); until then the model kernels see an empty row
                    s_len[row] = 0;
                    a.lengths[row] = 0;
                    a.pf_pos[row] = 0;
                } else {
                    s_len[row] = len;
                    a.lengths[row] = len;
                }
                a.len_shadow[row] = len;
                s_req[row] = id;
                s_used[n_used + j] = row;
                a.new_idx[j] = row;
                atomicMax(&s_carry[0], before + take);
            }
            k_adm = min(first_fail, min(n_cand, base + T));
            take_before += tott;
            len_before += totl;
            if (first_fail != 0x7fffffff) break;   // uniform: read from shared memory after the barrier
        }
        __syncthreads();
        const int pages_taken = s_carry[0];
        // unoccupied rows that got nothing: length 0 (:109-112)
        for (int j = k_adm + tid; j < n_free_rows; j += T) {
            const int row = s_list[j];
            s_len[row] = 0;
            a.lengths[row] = 0;
            a.len_shadow[row] = 0;
        }
        __syncthreads();
        // quirk Q1 (:113-118): any unoccupied row => the whole stale host array is copied back
        if (a.compat && n_free_rows > 0)
            for (int r = tid; r < B; r += T) {
                const int l0 = a.len_shadow[r];
                s_len[r] = l0;
                a.lengths[r] = l0;
            }
        fh = (fh + pages_taken) % nb;
        F -= pages_taken;
        qh = (qh + k_adm) % q_cap;
        qc -= k_adm;
        n_used += k_adm;
        __syncthreads();
    }

    SCHED_PH(7);
    GRIDDEP_TRIGGER_LATE();
    // ---- write the mirrors back ----
    for (int r = tid; r < B; r += T) {
        a.row_req[r] = s_req[r];
        a.npages[r] = s_np[r];
    }
    for (int i = tid; i < n_used; i += T) a.used[i] = s_used[i];

    // ================= chunked prefill (opt-in): this step's share of the prompts still being prefilled =====
    // Rows in admission order (the used list) take 16-position granules of their remaining prompt until the
    // step's budget is spent -- a prefix fill decided by one scan.  A row whose last granule is scheduled here
    // becomes active in this very step: the merged projection computes its granules and its latest-token q/K/V
    // in one launch, the attention that follows sees all of it.
    int n_gran_chunk = 0, n_done_chunk = 0;
    if (a.chunk > 0) {
        const int budget = a.chunk / kGran;
        int g_before = 0;
        for (int base = 0; base < n_used; base += T) {
            const int i = base + tid;
            int row = -1, p = -1, L = 0, g_rem = 0;
            if (i < n_used) {
                row = s_used[i];
                p = a.pf_pos[row];
                if (p >= 0) {
                    L = a.len_shadow[row];
                    g_rem = (L - p + kGran - 1) / kGran;
                }
            }
            int tot;
            const int before = g_before + block_scan_excl(g_rem, &tot, s_warp, scan_phase);
            const int take = max(0, min(g_rem, budget - before));
            const int done = (p >= 0 && take == g_rem) ? 1 : 0;
            int totd;
            const int dpos = n_done_chunk + block_scan_excl(done, &totd, s_warp, scan_phase);
            if (take > 0 || done) {
                for (int c = 0; c < take; ++c)
                    if (before + c < a.max_gran) {
                        a.gran[before + c].row = row;
                        a.gran[before + c].j0 = p + c * kGran;
                    }
                if (done) {
                    s_len[row] = L;
                    a.lengths[row] = L;
                    a.pf_pos[row] = -1;
                    a.new_idx[dpos] = row;   // (exact-order mode prefills a whole row when it becomes active)
                } else {
                    a.pf_pos[row] = p + take * kGran;
                }
            }
            g_before += tot;
            n_done_chunk += totd;
        }
        n_gran_chunk = min(g_before, budget);
        __syncthreads();
    }

    // ---- work lists of this step: active rows, prefill granules of the new rows ----
    int n_act = 0;
    for (int base = 0; base < B; base += T) {
        const int r = base + tid;
        const int on = (r < B && s_len[r] > 0) ? 1 : 0;
        int tot;
        const int pos = block_scan_excl(on, &tot, s_warp, scan_phase);
        if (on) a.act_rows[n_act + pos] = r;
        n_act += tot;
    }
    int n_gran = 0;
    for (int base = 0; base < k_adm && a.chunk <= 0; base += T) {
        const int j = base + tid;
        int row = -1, n = 0;
        if (j < k_adm) {
            row = s_list[j];   // free row j took queue item j
            n = (s_len[row] + kGran - 1) / kGran;
        }
        int tot;
        const int pos = block_scan_excl(n, &tot, s_warp, scan_phase);
        for (int c = 0; c < n; ++c) {
            if (n_gran + pos + c < a.max_gran) {
                a.gran[n_gran + pos + c].row = row;
                a.gran[n_gran + pos + c].j0 = c * kGran;
            }
        }
        n_gran += tot;
    }

    SCHED_PH(8);
    if (tid == 0) {
        SchedVars* v = a.v;
        v->f_head = fh;
        v->f_count = F;
        v->q_head = qh;
        v->q_count = qc;
        v->n_used = n_used;
        v->n_new = (a.chunk > 0) ? n_done_chunk : k_adm;
        v->admitted = sv.admitted + k_adm;
        v->preemptions = sv.preemptions;
        v->iter = sv.iter + 1;
        if (n_fin_now != sv.n_fin) {
            // mli_engine_poll_finished reads this count from another stream while the engine runs: the
            // token lists and ids it covers (written by other threads, many barriers ago) come first.  The
            // fence (0.35 us on the critical path of a 75 us step) only once polling is in use
            if (sv.poll) __threadfence();
            v->n_fin = n_fin_now;
        }
        v->n_req = n_avail;
        v->max_used = max(sv.max_used, n_used);
        v->min_free = min(sv.min_free, F);
        if (a.chunk > 0) n_gran = n_gran_chunk;
        a.counts[0] = n_act;
        a.counts[1] = min(n_gran, a.max_gran);
        if (a.trace != nullptr && a.trace[0] >= 1 && a.trace[0] - 1 < a.trace[1]) {
            a.trace[8 + 8 * (a.trace[0] - 1) + 6] = (unsigned long long)n_act;    // for tools/step_timeline.py
            a.trace[8 + 8 * (a.trace[0] - 1) + 7] = (unsigned long long)n_gran;
        }
        // is_done (item_storage.cpp:186-188): nothing processing and nothing queued
        if (n_used + qc == 0) {
            v->done = 1;
            *a.done_host = 1;
            __threadfence_system();
        } else if (n_used == 0 && F == nb) {
            // Nothing is resident, the whole pool is free and the head of the queue still was not admitted:
            // it never will be (a pre-empted request that outgrew the pool).  The reference spins here for
            // ever (paged_item_storage.cpp:84-113 keeps returning 0 new items); the engine ends the job
            v->error = max(v->error, 4);
            v->done = 1;
            *a.done_host = 1;
            __threadfence_system();
        } else {
            v->steps = sv.steps + 1;
            if (sv.done) {   // requests arrived after the engine had gone idle
                v->done = 0;
                *a.done_host = 0;
            }
        }
    }
    SCHED_PH(9);
#undef SCHED_PH
}

__global__ void engine_reset_kernel(SchedArgs a, float* pool, size_t page_floats, int max_req) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = gridDim.x * blockDim.x;
    for (int r = i; r < a.B; r += n) {
        a.row_req[r] = -1;
        a.lengths[r] = 0;
        a.len_shadow[r] = 0;
        a.npages[r] = 0;
        a.new_idx[r] = 0;
        a.pf_pos[r] = -1;
        for (int j = 0; j < a.R; ++j) a.dec[(size_t)r * a.R + j] = MLI_EMPTY_ROW_TOKEN_ID;
    }
    for (int b = i; b < a.n_blocks; b += n) a.free_ring[b] = pool + (size_t)b * page_floats;
    for (int q = i; q < max_req; q += n) {
        a.req_cnt[q] = 0;
        a.req_plen[q] = 0;
    }
    if (i == 0) {
        SchedVars* v = a.v;
        v->q_head = 0; v->q_count = 0; v->q_cap = max_req + 1;
        v->f_head = 0; v->f_count = a.n_blocks;
        v->n_used = 0; v->n_fin = 0; v->n_new = 0;
        v->iter = 0; v->done = 0; v->error = 0; v->n_req = 0;
        v->steps = 0; v->generated = 0; v->preemptions = 0; v->admitted = 0;
        v->max_used = 0; v->min_free = a.n_blocks; v->poll = 0;
        a.counts[0] = 0;
        a.counts[1] = 0;
        *a.n_avail = 0;
    }
}

// scatter (offsets, tokens) of requests [first, first + n) into the request table.  Lengths that the
// engine cannot represent are flagged (v->error) and clamped so nothing is written out of bounds:
//   2 = prompt length outside [1, n_sequence - 1]
//   3 = a prompt needs more KV pages than the pool has (it could never be admitted)
__global__ void engine_append_kernel(SchedArgs a, const int* __restrict__ offs,
                                     const int* __restrict__ toks, int first, int n_req) {
    const int k = blockIdx.x;
    if (k >= n_req) return;
    const int q = first + k;
    const int o = offs[k];
    int n = offs[k + 1] - o;
    int err = 0;
    if (n < 1 || n + 1 > a.S) {
        err = 2;
        n = n < 1 ? 0 : a.S - 1;
    }
    const int need = max((n + a.R + kPage - 1) / kPage, MLI_DEFAULT_INIT_NUM_BLOCKS);
    if (!err && need > a.n_blocks) err = 3;
    for (int j = threadIdx.x; j < n; j += blockDim.x) a.req_tok[(size_t)q * a.S + j] = toks[o + j];
    if (threadIdx.x == 0) {
        a.req_cnt[q] = err ? 0 : n;   // a flagged request is never queued into a row (the job fails)
        a.req_plen[q] = n;
        if (err) atomicMax(&a.v->error, err);
    }
}

// runs after the append kernel on the same stream: the rows are complete, the scheduler may queue them
__global__ void engine_publish_kernel(SchedArgs a, int n_total) {
    __threadfence();
    *a.n_avail = n_total;
}

// finished requests [first, first + n) in finish order -> ids and exclusive token offsets (one CTA)
__global__ void engine_fin_offsets_kernel(const int* __restrict__ fin_ids, const int* __restrict__ req_cnt,
                                          int first, int n, int* __restrict__ out_ids,
                                          int* __restrict__ out_offs) {
    __shared__ int s_warp[64];
    int phase = 0, carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int k = base + threadIdx.x;
        int id = -1, c = 0;
        if (k < n) {
            id = fin_ids[first + k];
            c = req_cnt[id];
        }
        int tot;
        const int off = carry + block_scan_excl(c, &tot, s_warp, phase);
        if (k < n) {
            out_ids[k] = id;
            out_offs[k] = off;
        }
        carry += tot;
    }
    if (threadIdx.x == 0) out_offs[n] = carry;
}

// token lists of those requests, packed (one CTA per request)
__global__ void engine_fin_gather_kernel(const int* __restrict__ req_tok, int S, const int* __restrict__ ids,
                                         const int* __restrict__ offs, int* __restrict__ out_toks) {
    const int k = blockIdx.x;
    const int id = ids[k], o = offs[k], c = offs[k + 1] - o;
    for (int j = threadIdx.x; j < c; j += blockDim.x) out_toks[o + j] = req_tok[(size_t)id * S + j];
}

}  // namespace mli

using namespace mli;

struct mli_engine {
    mli_ctx* ctx = nullptr;
    mli_engine_cfg cfg{};
    const float *emb = nullptr, *pos = nullptr, *wk = nullptr, *wq = nullptr, *wv = nullptr;
    SchedArgs a{};
    float* pool = nullptr;
    bool own_pool = false;
    float *q_out = nullptr, *attn_out = nullptr, *score = nullptr;
    TileDesc* tiles = nullptr;
    int* n_tiles = nullptr;
    int max_tiles = 0;
    int* done_host = nullptr;  // mapped pinned done word
    int* stage_buf = nullptr;  // device staging for host prompts (mli_engine_submit / _enqueue)
    size_t stage_ints = 0;
    std::vector<void*> allocs;
    cudaGraphExec_t graph_exec = nullptr;   // one step (bounded runs)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graphn_exec = nullptr;  // kStepsPerGraph steps (runs to completion)
    cudaGraph_t graphn = nullptr;
    bool froze_ws = false;         // this engine holds one count of ctx->ws_frozen
    bool registered_weights = false;
    int n_req = 0;                 // requests submitted + enqueued since the last reset
    int n_polled = 0;              // finished requests already handed out by mli_engine_poll_finished
    bool poll_armed = false;       // the device has been told that mli_engine_poll_finished is in use
    bool pending_wake = false;     // requests were enqueued since the last run started: the done word may be
                                   // stale (the device can set it before it has seen them)
    std::mutex mu;                 // guards n_req, stage_buf and the ingest stream (mli_engine_enqueue may
                                   // be called from another host thread while mli_engine_run executes)
    mli_engine_stats stats{};
    cudaEvent_t ev_submit = nullptr, ev_end = nullptr;  // job timing: start of submit .. end of run
    bool submit_timed = false;
    cudaEvent_t ring_ev[4] = {};
    int launches_per_step = 0;     // kernels in the captured step graph
    int kv_bf16 = 0;               // page format the engine was created with
    cudaEvent_t prof_ev[32] = {};  // profile mode: one event pair per step of a batch (attention)
    cudaEvent_t prof_gev[32] = {}; // the same for the merged GEMM
    int* prof_lengths = nullptr;   // pinned [16][B + 2] (lengths, then the scheduler's two counts), profile mode
    // the engine runs on its own non-blocking stream: the caller's stream may be the legacy
    // default stream, which cannot be captured into a graph
    cudaStream_t stream = nullptr;
    cudaEvent_t join_ev = nullptr;
    // streaming side channels: requests come in on `ingest`, finished token lists go out
    // on `io`, both concurrent with the step graphs on `stream`
    cudaStream_t ingest = nullptr, io = nullptr;
    cudaEvent_t reset_ev = nullptr;   // recorded after every reset; the ingest stream waits for it
    int* cnt_host = nullptr;          // pinned landing word of finished_count()
    int* res_host = nullptr;          // mapped pinned staging for finished lists: ids | offsets | tokens
    int* res_dev = nullptr;           // device alias of res_host
    size_t res_ints = 0;
};

namespace {
// while alive, every launch helper (they take the stream from the context) targets the engine's
// stream, ordered after whatever the caller already enqueued on the context's stream
struct StreamScope {
    mli_ctx* ctx;
    cudaStream_t saved;
    StreamScope(mli_engine* e) : ctx(e->ctx), saved(e->ctx->stream) {
        cudaEventRecord(e->join_ev, saved);
        cudaStreamWaitEvent(e->stream, e->join_ev, 0);
        ctx->stream = e->stream;
    }
    ~StreamScope() { ctx->stream = saved; }
};
}  // namespace

namespace {

// one thread per row up to 1024; fewer warps make the (many) block barriers of the scheduler cheaper
int sched_threads(int n_batch) {
    int t = ((n_batch + 31) / 32) * 32;
    if (t < 128) t = 128;
    if (t > kSchedThreads) t = kSchedThreads;
    return t;
}

template <typename T>
int dev_alloc(mli_engine* e, T** out, size_t n) {
    void* p = nullptr;
    MLI_CUDA(cudaMalloc(&p, sizeof(T) * (n > 0 ? n : 1)));
    e->allocs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return 0;
}

// the model part of one engine iteration: n_forward_rounds x (encoder -> attention -> decoder)
// (inference_model.cpp:52-82).  ev0/ev1, when given, bracket the first round's fused attention.
int enqueue_model(mli_engine* e, cudaEvent_t ev0, cudaEvent_t ev1, cudaEvent_t gev0 = nullptr,
                  cudaEvent_t gev1 = nullptr) {
    mli_ctx* ctx = e->ctx;
    const mli_engine_cfg& c = e->cfg;
    const int B = c.n_batch, S = c.n_sequence, d = c.emb_dim, V = c.n_vocab;
    int rc;
    // everything after the scheduler may be chained with programmatic dependent launch
    struct PdlScope {
        mli_ctx* c;
        explicit PdlScope(mli_ctx* c_) : c(c_) { c->use_pdl = c->opt_pdl != 0; }
        ~PdlScope() { c->use_pdl = false; }
    } pdl(ctx);
    const bool tc = (ctx->gemm_mode == 0 && ctx->tc_available && d % 128 == 0);
    if (tc) {
        // tensor-core mode: the scheduler already listed the active rows and the 16-position prefill
        // granules of the new rows; encoder -> ONE merged projection (latest K,q,V + prefill K,V)
        // (chunked prefill: a row that is still prefilling has length 0; its granules are bounded by the
        // admission length the scheduler keeps in len_shadow)
        if ((rc = launch_paged_encoder_tiles(ctx, e->emb, e->pos, nullptr, e->a.row_req, e->a.req_tok,
                                             e->a.page_table, e->a.gran, e->a.counts + 1, e->a.max_gran,
                                             e->a.chunk > 0 ? e->a.len_shadow : e->a.lengths, S, d, kPage)))
            return rc;
    } else {
        if ((rc = launch_build_new_row_tiles(ctx, e->a.new_idx, e->a.lengths, 0, &e->a.v->n_new, e->tiles,
                                             e->n_tiles, e->max_tiles)))
            return rc;
        if ((rc = launch_paged_encoder_tiles(ctx, e->emb, e->pos, nullptr, e->a.row_req, e->a.req_tok,
                                             e->a.page_table, e->tiles, e->n_tiles, e->max_tiles,
                                             e->a.lengths, S, d)))
            return rc;
        if ((rc = launch_prefill_kv_paged_simt(ctx, e->a.page_table, e->tiles, e->n_tiles, e->max_tiles,
                                               e->a.lengths, e->wk, e->wv, S, d)))
            return rc;
    }
    for (int round = 0; round < c.n_forward_rounds; ++round) {
        if (tc) {
            if (round == 0 && gev0) {
                ctx->gemm_ev_start = gev0;
                ctx->gemm_ev_stop = gev1;
            }
            rc = launch_step_qkv_tc(ctx, e->a.page_table, e->a.lengths, e->a.act_rows, e->a.counts,
                                    e->a.gran, e->a.max_gran, round == 0 ? 1 : 0, e->wk, e->wq, e->wv,
                                    e->q_out, B, S, d, e->a.chunk > 0 ? e->a.len_shadow : nullptr);
            ctx->gemm_ev_start = ctx->gemm_ev_stop = nullptr;
        }
        else
            rc = launch_qkv_latest_paged_simt(ctx, e->a.page_table, e->a.lengths, e->wk, e->wq, e->wv,
                                              e->q_out, B, S, d);
        if (rc) return rc;
        if (round == 0 && ev0) {
            ctx->attn_ev_start = ev0;
            ctx->attn_ev_stop = ev1;
        }
        // lengths[] was last written by the scheduler (round 0) or the previous round's decoder: at least
        // two kernels up the chain, complete before the attention can start
        ctx->attn_lengths_final = 1;
        rc = launch_decode_attention_paged(ctx, e->q_out, e->a.page_table, e->a.lengths, e->attn_out,
                                           nullptr, B, S, d);
        ctx->attn_lengths_final = 0;
        ctx->attn_ev_start = ctx->attn_ev_stop = nullptr;
        if (rc) return rc;
        int n_split = 1;
        if (ctx->gemm_mode == 0 && ctx->tc_available)
            rc = launch_logits_tc(ctx, e->attn_out, e->emb, e->score, B, V, d, &n_split, e->a.act_rows,
                                  e->a.counts);
        else
            rc = launch_logits_simt(ctx, e->attn_out, e->emb, e->score, B, V, d);
        if (rc) return rc;
        if ((rc = launch_paged_decoder(ctx, e->score, n_split, nullptr, e->a.dec, e->a.lengths,
                                       e->a.page_table, e->pos, e->emb, B, V, S, d, c.n_forward_rounds,
                                       round)))
            return rc;
    }
    return 0;
}

// one engine iteration: scheduler (the head of the step: plain launch, fully ordered after the
// previous step), then the model
// chained = true: the scheduler itself is launched with programmatic dependent launch (it follows
// the previous step's decoder inside one multi-step graph)
int enqueue_step(mli_engine* e, bool chained = false) {
    mli_ctx* ctx = e->ctx;
    e->a.trace = ctx->trace;
    if (chained && ctx->opt_pdl) {
        ctx->use_pdl = true;
        int rc = launch_kernel(ctx, sched_step_kernel, dim3(1), dim3(sched_threads(e->cfg.n_batch)),
                               sched_smem_bytes(e->cfg.n_batch), e->a);
        ctx->use_pdl = false;
        if (rc) return rc;
    } else {
        sched_step_kernel<<<1, sched_threads(e->cfg.n_batch), sched_smem_bytes(e->cfg.n_batch), ctx->stream>>>(e->a);
        MLI_LAUNCH_CHECK();
    }
    return enqueue_model(e, nullptr, nullptr);
}

void drop_graph(mli_engine* e) {
    if (e->graphn_exec) cudaGraphExecDestroy(e->graphn_exec);
    if (e->graphn) cudaGraphDestroy(e->graphn);
    e->graphn_exec = nullptr;
    e->graphn = nullptr;
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->graph) cudaGraphDestroy(e->graph);
    e->graph_exec = nullptr;
    e->graph = nullptr;
    if (e->froze_ws) {   // other engines of the context may still hold captured graphs
        e->ctx->ws_frozen -= 1;
        e->froze_ws = false;
    }
}

const char* sched_error_text(int code) {
    switch (code) {
        case 1: return "engine: a token arrived for a row that is not processing";
        case 2: return "engine: a prompt length is outside [1, n_sequence - 1]";
        case 3: return "engine: a prompt needs more KV pages than the pool holds (it could never be admitted)";
        case 4: return "engine: a pre-empted request outgrew the KV pool (it needs more pages than n_blocks) and can "
                       "never be admitted again; the job was ended with requests still queued";
    }
    return "engine: scheduler error";
}

// validate host prompts (the device path is validated by engine_append_kernel)
int check_host_prompts(const mli_engine* e, int n_req, const int* offs) {
    const int S = e->cfg.n_sequence, R = e->cfg.n_forward_rounds;
    for (int i = 0; i < n_req; ++i) {
        const int n = offs[i + 1] - offs[i];
        MLI_REQUIRE(n >= 1 && n + 1 <= S, "prompt length must be in [1, n_sequence-1]");
        const int need = std::max((n + R + kPage - 1) / kPage, MLI_DEFAULT_INIT_NUM_BLOCKS);
        if (need > e->cfg.n_blocks) {
            set_error("No enough block memories to return: a prompt needs more KV pages than the pool holds");
            return MLI_ERR_NO_BLOCKS;
        }
    }
    return 0;
}

// stage (host prompts) and append requests [first, first + n_req) on `st`, then publish the new total.
// Caller holds e->mu.  own_staging = true (mli_engine_enqueue: may run while ANOTHER thread is capturing or
// replaying the step graph on the engine's stream): the staging buffer is a stream-ordered allocation on `st`,
// so nothing here synchronises the device or touches the engine's stream; false (mli_engine_submit, on the
// engine's stream itself): the engine's persistent staging buffer.
int append_requests(mli_engine* e, int first, int n_req, const int* offs, const int* toks, int is_device,
                    cudaStream_t st, bool own_staging) {
    if (n_req <= 0) return 0;
    const int* d_offs = offs;
    const int* d_toks = toks;
    int* tmp = nullptr;
    if (!is_device) {
        const int total = offs[n_req] - offs[0];
        const size_t need = (size_t)n_req + 1 + (size_t)total;
        int* buf = nullptr;
        if (own_staging) {
            MLI_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), sizeof(int) * need, st));
            buf = tmp;
        } else {
            if (e->stage_ints < need) {
                MLI_CUDA(cudaStreamSynchronize(st));   // an earlier submit may still read the old buffer
                if (e->stage_buf) cudaFree(e->stage_buf);
                e->stage_buf = nullptr;
                e->stage_ints = 0;
                MLI_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->stage_buf), sizeof(int) * need));
                e->stage_ints = need;
            }
            buf = e->stage_buf;
        }
        MLI_CUDA(cudaMemcpyAsync(buf, offs, sizeof(int) * ((size_t)n_req + 1), cudaMemcpyHostToDevice, st));
        MLI_CUDA(cudaMemcpyAsync(buf + n_req + 1, toks + offs[0], sizeof(int) * (size_t)total,
                                 cudaMemcpyHostToDevice, st));
        d_offs = buf;
        d_toks = buf + n_req + 1 - offs[0];   // the kernel indexes tokens with the caller's offsets
    }
    engine_append_kernel<<<n_req, 128, 0, st>>>(e->a, d_offs, d_toks, first, n_req);
    MLI_LAUNCH_CHECK();
    engine_publish_kernel<<<1, 1, 0, st>>>(e->a, first + n_req);
    MLI_LAUNCH_CHECK();
    if (tmp) MLI_CUDA(cudaFreeAsync(tmp, st));
    return 0;
}

// how many requests have finished so far: a 4-byte copy on the io stream, concurrent with the step graphs
// (the scheduler publishes the count in device memory only -- a host-visible word would cost it a
// system-scope fence on every step that retires a request).  A request counted here has its complete token
// list in the request table: the scheduler instance that wrote both finished before the copy could read.
int finished_count(mli_engine* e, int* n_fin) {
    if (!e->cnt_host) MLI_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&e->cnt_host), 64, cudaHostAllocDefault));
    MLI_CUDA(cudaMemcpyAsync(e->cnt_host, &e->a.v->n_fin, sizeof(int), cudaMemcpyDeviceToHost, e->io));
    MLI_CUDA(cudaStreamSynchronize(e->io));
    *n_fin = *e->cnt_host;
    return 0;
}

// finished requests [first, first + n) packed into the mapped pinned staging area on the io stream:
// res_host = ids[n] | offsets[n + 1] | tokens.  Returns after the io stream has drained (the engine's
// step graphs keep running on their own stream).
int gather_finished(mli_engine* e, int first, int n, const int** ids, const int** offs, const int** toks) {
    const size_t S = (size_t)e->cfg.n_sequence, NR = (size_t)e->cfg.max_requests;
    const size_t need = 2 * NR + 1 + NR * S;
    if (e->res_ints < need) {
        if (e->res_host) cudaFreeHost(e->res_host);
        e->res_host = nullptr;
        e->res_ints = 0;
        MLI_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&e->res_host), sizeof(int) * need, cudaHostAllocMapped));
        MLI_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->res_dev), e->res_host, 0));
        e->res_ints = need;
    }
    int* d_ids = e->res_dev;
    int* d_offs = e->res_dev + NR;
    int* d_toks = e->res_dev + 2 * NR + 1;
    e->res_host[NR] = 0;
    if (n > 0) {
        engine_fin_offsets_kernel<<<1, 1024, 0, e->io>>>(e->a.fin_ids, e->a.req_cnt, first, n, d_ids, d_offs);
        MLI_LAUNCH_CHECK();
        engine_fin_gather_kernel<<<n, 256, 0, e->io>>>(e->a.req_tok, (int)S, d_ids, d_offs, d_toks);
        MLI_LAUNCH_CHECK();
        MLI_CUDA(cudaStreamSynchronize(e->io));
    }
    *ids = e->res_host;
    *offs = e->res_host + NR;
    *toks = e->res_host + 2 * NR + 1;
    return 0;
}

}  // namespace

extern "C" {

int mli_engine_create(mli_ctx* ctx, const mli_engine_cfg* cfg, const float* emb_table,
                      const float* pos_table, const float* wk, const float* wq, const float* wv,
                      mli_engine** out) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(cfg && out, "null argument");
    MLI_REQUIRE(cfg->n_batch > 0 && cfg->n_sequence % kPage == 0 && cfg->emb_dim % 4 == 0 &&
                    cfg->n_vocab > 0 && cfg->n_blocks > 0 && cfg->max_requests > 0,
                "bad engine dims");
    MLI_REQUIRE(cfg->n_forward_rounds >= 1 && cfg->n_forward_rounds <= kPage,
                "n_forward_rounds must be 1..16");
    MLI_REQUIRE(cfg->max_new_tokens >= 0 && cfg->max_prefill_positions >= 0 && cfg->prefill_chunk_positions >= 0,
                "negative policy value");
    MLI_REQUIRE(cfg->prefill_chunk_positions == 0 || (cfg->prefill_chunk_positions >= kPage && !cfg->compat_stale_lengths),
                "prefill_chunk_positions must be 0 or >= 16, and needs corrected lengths (compat_stale_lengths = 0)");
    MLI_REQUIRE(sched_smem_bytes(cfg->n_batch) <= 220 * 1024,
                "n_batch too large for the device scheduler's shared-memory mirrors (max ~10000 rows per GPU)");
    {
        int rc0 = ensure_dyn_smem(ctx, sched_step_kernel, sched_smem_bytes(cfg->n_batch));
        if (rc0) return rc0;
    }
    mli_engine* e = new mli_engine();
    e->ctx = ctx;
    e->cfg = *cfg;
    e->emb = emb_table; e->pos = pos_table; e->wk = wk; e->wq = wq; e->wv = wv;
    const int B = cfg->n_batch, S = cfg->n_sequence, d = cfg->emb_dim, V = cfg->n_vocab;
    const int W = S / kPage, R = cfg->n_forward_rounds, NR = cfg->max_requests;
    SchedArgs& a = e->a;
    a.B = B; a.S = S; a.W = W; a.R = R; a.n_blocks = cfg->n_blocks;
    a.compat = cfg->compat_stale_lengths;
    a.max_new = cfg->max_new_tokens;
    a.max_prefill = cfg->max_prefill_positions;
    a.chunk = cfg->prefill_chunk_positions / kPage * kPage;
    int rc = 0;
#define A(call) if ((rc = (call))) { mli_engine_destroy(e); return rc; }
    A(dev_alloc(e, &a.v, 1));
    A(dev_alloc(e, &a.req_tok, (size_t)NR * S));
    A(dev_alloc(e, &a.req_cnt, NR));
    A(dev_alloc(e, &a.req_plen, NR));
    A(dev_alloc(e, &a.queue, NR + 1));
    A(dev_alloc(e, &a.fin_ids, NR));
    A(dev_alloc(e, &a.row_req, B));
    A(dev_alloc(e, &a.lengths, B));
    A(dev_alloc(e, &a.len_shadow, B));
    A(dev_alloc(e, &a.used, B));
    A(dev_alloc(e, &a.npages, B));
    A(dev_alloc(e, &a.page_table, (size_t)B * W));
    A(dev_alloc(e, &a.free_ring, cfg->n_blocks));
    A(dev_alloc(e, &a.dec, (size_t)B * R));
    A(dev_alloc(e, &a.new_idx, B));
    A(dev_alloc(e, &a.pf_pos, B));
    A(dev_alloc(e, &a.act_rows, B));
    a.max_gran = B * W;
    A(dev_alloc(e, &a.gran, (size_t)a.max_gran));
    A(dev_alloc(e, &a.counts, 4));
    {
        int* p = nullptr;
        A(dev_alloc(e, &p, 4));
        a.n_avail = p;
    }
    A(dev_alloc(e, &e->q_out, (size_t)B * d));
    A(dev_alloc(e, &e->attn_out, (size_t)B * d));
    A(dev_alloc(e, &e->score, (size_t)kMaxLogitSplit * B * V));   // split-K partial logits
    e->max_tiles = B * ceil_div(S, kTileM);
    A(dev_alloc(e, &e->tiles, e->max_tiles));
    A(dev_alloc(e, &e->n_tiles, 4));
    // compact KV format: only with the tensor-core projection (it writes the bf16 rows)
    if (ctx->kv_bf16 && !(ctx->gemm_mode == 0 && ctx->tc_available && d % 128 == 0 && V % 128 == 0)) {
        set_error("engine: the compact KV format needs the tcgen05 GEMM mode, emb_dim % 128 == 0 and n_vocab % 128 == 0");
        mli_engine_destroy(e);
        return MLI_ERR_UNSUPPORTED;
    }
    e->kv_bf16 = ctx->kv_bf16;
    const size_t page_floats = (size_t)kPage * page_pos_floats(d, ctx->kv_bf16);
    if (cfg->page_pool) {
        e->pool = cfg->page_pool;
    } else {
        void* p = nullptr;
        cudaError_t ce = cudaMalloc(&p, sizeof(float) * page_floats * (size_t)cfg->n_blocks);
        if (ce != cudaSuccess) { mli_engine_destroy(e); return cuda_fail(ce, __FILE__, __LINE__); }
        e->pool = reinterpret_cast<float*>(p);
        e->own_pool = true;
    }
    {
        cudaError_t ce = cudaHostAlloc(reinterpret_cast<void**>(&e->done_host), 128, cudaHostAllocMapped);
        if (ce != cudaSuccess) { mli_engine_destroy(e); return cuda_fail(ce, __FILE__, __LINE__); }
        int* dptr = nullptr;
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), e->done_host, 0);
        a.done_host = dptr;
        e->done_host[0] = 0;
    }
    cudaEventCreate(&e->ev_submit);
    cudaEventCreate(&e->ev_end);
    for (auto& ev : e->ring_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->join_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->reset_ev, cudaEventDisableTiming);
    {
        cudaError_t ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->ingest, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->io, cudaStreamNonBlocking);
        if (ce != cudaSuccess) { mli_engine_destroy(e); return cuda_fail(ce, __FILE__, __LINE__); }
    }
#undef A
    StreamScope scope(e);
    // the engine owns the weight lifetime for as long as it lives: split them once (tcgen05 mode)
    if (ctx->tc_available) {
        if ((rc = tc_register_weights(ctx, wk, wq, wv, emb_table, d, V))) {
            mli_engine_destroy(e);
            return rc;
        }
        e->registered_weights = true;
    }
    // zero state so a warm-up step is harmless, then run one un-captured step on the empty engine:
    // it sizes every workspace the captured graph will later hold pointers to.
    engine_reset_kernel<<<64, 256, 0, ctx->stream>>>(a, e->pool, page_floats, NR);
    MLI_LAUNCH_CHECK();
    if ((rc = enqueue_step(e))) { mli_engine_destroy(e); return rc; }
    cudaEventRecord(e->reset_ev, ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) { mli_engine_destroy(e); return cuda_fail(ce, __FILE__, __LINE__); }
    *out = e;
    return MLI_OK;
}

int mli_engine_destroy(mli_engine* e) {
    if (!e) return MLI_OK;
    MLI_ENTER(e->ctx, "null ctx");
    cudaStreamSynchronize(e->ctx->stream);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->ingest) cudaStreamSynchronize(e->ingest);
    if (e->io) cudaStreamSynchronize(e->io);
    drop_graph(e);
    // only what this engine registered (reference-counted: another engine of the context may share
    // the same weight pointers)
    if (e->registered_weights) tc_unregister_weights(e->ctx, e->wk, e->emb);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->ingest) cudaStreamDestroy(e->ingest);
    if (e->io) cudaStreamDestroy(e->io);
    if (e->join_ev) cudaEventDestroy(e->join_ev);
    if (e->reset_ev) cudaEventDestroy(e->reset_ev);
    for (void* p : e->allocs) cudaFree(p);
    if (e->own_pool && e->pool) cudaFree(e->pool);
    if (e->stage_buf) cudaFree(e->stage_buf);
    if (e->done_host) cudaFreeHost(e->done_host);
    if (e->res_host) cudaFreeHost(e->res_host);
    if (e->cnt_host) cudaFreeHost(e->cnt_host);
    if (e->prof_lengths) cudaFreeHost(e->prof_lengths);
    for (auto& ev : e->prof_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : e->prof_gev)
        if (ev) cudaEventDestroy(ev);
    if (e->ev_submit) cudaEventDestroy(e->ev_submit);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    for (auto& ev : e->ring_ev)
        if (ev) cudaEventDestroy(ev);
    delete e;
    return MLI_OK;
}

int mli_engine_submit(mli_engine* e, int n_req, const int* prompt_offsets, const int* prompt_tokens,
                      int is_device) {
    MLI_REQUIRE(e && prompt_offsets && prompt_tokens, "null argument");
    MLI_ENTER(e->ctx, "null ctx");
    MLI_REQUIRE(n_req >= 0 && n_req <= e->cfg.max_requests, "too many requests");
    mli_ctx* ctx = e->ctx;
    int rc;
    // everything that can fail on the host is checked before the job clock starts
    if (!is_device && (rc = check_host_prompts(e, n_req, prompt_offsets))) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    StreamScope scope(e);
    const size_t page_floats = (size_t)kPage * page_pos_floats(e->cfg.emb_dim, e->kv_bf16);
    MLI_CUDA(cudaStreamSynchronize(e->ingest));   // no append of the previous job may still be in flight
    MLI_CUDA(cudaEventRecord(e->ev_submit, ctx->stream));
    e->submit_timed = true;
    e->done_host[0] = 0;
    engine_reset_kernel<<<64, 256, 0, ctx->stream>>>(e->a, e->pool, page_floats, e->cfg.max_requests);
    MLI_LAUNCH_CHECK();
    if ((rc = append_requests(e, 0, n_req, prompt_offsets, prompt_tokens, is_device, ctx->stream, false))) return rc;
    MLI_CUDA(cudaEventRecord(e->reset_ev, ctx->stream));
    e->n_req = n_req;
    e->n_polled = 0;
    e->poll_armed = false;
    e->pending_wake = false;
    e->stats = mli_engine_stats{};
    return MLI_OK;
}

int mli_engine_enqueue(mli_engine* e, int n_req, const int* prompt_offsets, const int* prompt_tokens,
                       int is_device, int* first_id) {
    MLI_REQUIRE(e && prompt_offsets && prompt_tokens, "null argument");
    MLI_ENTER(e->ctx, "null ctx");
    MLI_REQUIRE(n_req >= 0, "negative request count");
    int rc;
    if (!is_device && (rc = check_host_prompts(e, n_req, prompt_offsets))) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    MLI_REQUIRE(e->n_req + n_req <= e->cfg.max_requests, "request table full (mli_engine_cfg.max_requests)");
    // on the ingest stream, concurrent with the step graphs; ordered after the last reset only
    MLI_CUDA(cudaStreamWaitEvent(e->ingest, e->reset_ev, 0));
    if ((rc = append_requests(e, e->n_req, n_req, prompt_offsets, prompt_tokens, is_device, e->ingest, true))) return rc;
    // the caller may reuse its buffers as soon as this returns
    MLI_CUDA(cudaStreamSynchronize(e->ingest));
    if (first_id) *first_id = e->n_req;
    e->n_req += n_req;
    e->pending_wake = true;   // the next mli_engine_run looks at the device state, not at a stale done word
    return MLI_OK;
}

int mli_engine_run(mli_engine* e, long long max_steps, int profile_attention) {
    MLI_REQUIRE(e, "null engine");
    MLI_ENTER(e->ctx, "null ctx");
    mli_ctx* ctx = e->ctx;
    MLI_REQUIRE(ctx->kv_bf16 == e->kv_bf16, "MLI_OPT_KV_FORMAT was changed after the engine was created");
    StreamScope scope(e);
    int rc;
    const int B = e->cfg.n_batch, d = e->cfg.emb_dim;
    float attn_ms = 0.f, gemm_ms = 0.f;
    double attn_bytes = 0.0, gemm_flops = 0.0;
    long long attn_launches = 0, gemm_launches = 0;
    MLI_CUDA(cudaEventRecord(e->ring_ev[0], ctx->stream));  // make ring events valid
    {
        // requests enqueued since the last run: launch at least one step, whose scheduler either takes
        // them over or finds nothing new and re-asserts the done word itself
        std::lock_guard<std::mutex> lk(e->mu);
        if (e->pending_wake) {
            e->done_host[0] = 0;
            e->pending_wake = false;
        }
    }
    // device time of the job: from the start of the submit that fed this run (else from here)
    if (!e->submit_timed) MLI_CUDA(cudaEventRecord(e->ev_submit, ctx->stream));
    e->submit_timed = false;
    long long it = 0;
    if (profile_attention) {
        // un-captured: every fused-attention launch is bracketed by its own pair of CUDA events and
        // the lengths it saw are copied out (for the algorithmic bytes).  Steps are enqueued in
        // batches without a host sync in between, so the device stays busy as it does in the graph.
        constexpr int kBatch = 16;
        if (!e->prof_ev[0]) {
            for (auto& ev : e->prof_ev) MLI_CUDA(cudaEventCreate(&ev));
            for (auto& ev : e->prof_gev) MLI_CUDA(cudaEventCreate(&ev));
            MLI_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&e->prof_lengths),
                                   sizeof(int) * ((size_t)B + 2) * kBatch, cudaHostAllocDefault));
        }
        bool finished = false;
        while (!finished) {
            int nb = 0;
            for (; nb < kBatch; ++nb, ++it) {
                if (max_steps > 0 && it >= max_steps) { finished = true; break; }
                sched_step_kernel<<<1, sched_threads(e->cfg.n_batch), sched_smem_bytes(e->cfg.n_batch), ctx->stream>>>(e->a);
                MLI_LAUNCH_CHECK();
                int* slot = e->prof_lengths + (size_t)nb * (B + 2);
                MLI_CUDA(cudaMemcpyAsync(slot, e->a.lengths, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost,
                                         ctx->stream));
                MLI_CUDA(cudaMemcpyAsync(slot + B, e->a.counts, sizeof(int) * 2, cudaMemcpyDeviceToHost,
                                         ctx->stream));
                if ((rc = enqueue_model(e, e->prof_ev[2 * nb], e->prof_ev[2 * nb + 1], e->prof_gev[2 * nb],
                                        e->prof_gev[2 * nb + 1])))
                    return rc;
            }
            MLI_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int k = 0; k < nb; ++k) {
                const int* slot = e->prof_lengths + (size_t)k * (B + 2);
                const double bytes = attention_algorithmic_bytes(slot, B, d, e->kv_bf16);
                if (bytes <= 0.0) continue;   // a step past the end of the job: nothing to attend
                if (ctx->gemm_mode == 0 && ctx->tc_available && d % 128 == 0) {
                    // active rows: K, q, V of one position; granules: K, V of the new rows' earlier
                    // positions (a granule is 16 positions, the last one of a row partly used: upper bound)
                    float gms = 0.f;
                    MLI_CUDA(cudaEventElapsedTime(&gms, e->prof_gev[2 * k], e->prof_gev[2 * k + 1]));
                    gemm_ms += gms;
                    const double fl = (6.0 * slot[B] + 4.0 * kPage * slot[B + 1]) * (double)d * d;
                    gemm_flops += fl;
                    ++gemm_launches;
                    // the largest launch of the job (bulk prefill: the tensor-pipe regime)
                    if (fl > e->stats.gemm_max_flops) {
                        e->stats.gemm_max_flops = fl;
                        e->stats.gemm_max_ms = gms;
                    }
                }
                float ms = 0.f;
                MLI_CUDA(cudaEventElapsedTime(&ms, e->prof_ev[2 * k], e->prof_ev[2 * k + 1]));
                attn_ms += ms;
                attn_bytes += bytes;
                ++attn_launches;
            }
            if (*reinterpret_cast<volatile int*>(e->done_host)) finished = true;
        }
    } else {
        // Runs to completion replay a graph of kStepsPerGraph steps (every scheduler after the
        // first is chained to the previous decoder with programmatic dependent launch; steps past
        // the end of the job find an empty engine and cost a few microseconds); bounded runs replay
        // a one-step graph.
#ifndef MLI_STEPS_PER_GRAPH
#define MLI_STEPS_PER_GRAPH 4
#endif
#ifndef MLI_GRAPHS_AHEAD
#define MLI_GRAPHS_AHEAD 2
#endif
        constexpr int kStepsPerGraph = MLI_STEPS_PER_GRAPH;
        const int per = (max_steps > 0) ? 1 : kStepsPerGraph;
        cudaGraphExec_t& gexec = (per == 1) ? e->graph_exec : e->graphn_exec;
        cudaGraph_t& g = (per == 1) ? e->graph : e->graphn;
        if (!gexec) {
            MLI_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            const long long l0 = mli_kernel_launch_count();
            rc = 0;
            for (int k = 0; k < per && !rc; ++k) rc = enqueue_step(e, k > 0);
            e->launches_per_step = (int)(mli_kernel_launch_count() - l0) / per;
            count_launch(-e->launches_per_step * per);   // captured, not launched
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
            if (rc || ce != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                g = nullptr;
                cudaGetLastError();
                return rc ? rc : cuda_fail(ce, __FILE__, __LINE__);
            }
            MLI_CUDA(cudaGraphInstantiate(&gexec, g, 0));
            if (!e->froze_ws) {
                ctx->ws_frozen += 1;
                e->froze_ws = true;
            }
        }
        const int kAhead = (per == 1) ? 4 : MLI_GRAPHS_AHEAD;   // graphs in flight before the host looks at `done`
        for (;; it += per) {
            if (max_steps > 0 && it >= max_steps) break;
            const int slot = (int)((it / per) % kAhead);
            if (it / per >= kAhead) MLI_CUDA(cudaEventSynchronize(e->ring_ev[slot]));
            if (*reinterpret_cast<volatile int*>(e->done_host)) break;
            MLI_CUDA(cudaGraphLaunch(gexec, ctx->stream));
            count_launch(e->launches_per_step * per);
            MLI_CUDA(cudaEventRecord(e->ring_ev[slot], ctx->stream));
        }
    }
    MLI_CUDA(cudaEventRecord(e->ev_end, ctx->stream));
    MLI_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    MLI_CUDA(cudaEventElapsedTime(&ms, e->ev_submit, e->ev_end));
    SchedVars hv;
    MLI_CUDA(cudaMemcpy(&hv, e->a.v, sizeof(hv), cudaMemcpyDeviceToHost));
    e->stats.steps = hv.steps;
    e->stats.generated_tokens = hv.generated;
    e->stats.preemptions = hv.preemptions;
    e->stats.admitted = hv.admitted;
    e->stats.n_finished = hv.n_fin;
    e->stats.gpu_ms = ms;
    e->stats.attn_ms = attn_ms;
    e->stats.attn_bytes = attn_bytes;
    e->stats.attn_launches = attn_launches;
    e->stats.gemm_ms = gemm_ms;
    e->stats.gemm_flops = gemm_flops;
    e->stats.gemm_launches = gemm_launches;
    e->stats.peak_resident_rows = hv.max_used;
    e->stats.min_free_pages = hv.min_free;
    if (hv.error) {
        set_error(sched_error_text(hv.error));
        return (hv.error == 3 || hv.error == 4) ? MLI_ERR_NO_BLOCKS : (hv.error == 2 ? MLI_ERR_ARG : MLI_ERR_STATE);
    }
    return MLI_OK;
}

int mli_engine_results(mli_engine* e, int* finished_ids, int* finished_offsets, int* finished_tokens,
                       int* n_finished) {
    MLI_REQUIRE(e && finished_ids && finished_offsets && finished_tokens && n_finished,
                "null argument");
    MLI_ENTER(e->ctx, "null ctx");
    mli_ctx* ctx = e->ctx;
    MLI_CUDA(cudaStreamSynchronize(e->stream));
    MLI_CUDA(cudaStreamSynchronize(ctx->stream));
    std::lock_guard<std::mutex> lk(e->mu);
    int n_fin = 0;
    int rc0 = finished_count(e, &n_fin);
    if (rc0) return rc0;
    const int *ids, *offs, *toks;
    int rc = gather_finished(e, 0, n_fin, &ids, &offs, &toks);
    if (rc) return rc;
    // one packed device-to-host transfer happened (the gather kernels wrote pinned memory); hand it out
    memcpy(finished_ids, ids, sizeof(int) * (size_t)n_fin);
    memcpy(finished_offsets, offs, sizeof(int) * ((size_t)n_fin + 1));
    memcpy(finished_tokens, toks, sizeof(int) * (size_t)offs[n_fin]);
    *n_finished = n_fin;
    return MLI_OK;
}

int mli_engine_poll_finished(mli_engine* e, int max_out, int* ids_out, int* offsets_out, int* tokens_out,
                             long long tokens_capacity, int* n_out) {
    MLI_REQUIRE(e && ids_out && offsets_out && tokens_out && n_out, "null argument");
    MLI_ENTER(e->ctx, "null ctx");
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->poll_armed) {
        // tell the scheduler to order its finished lists before the count from now on (see sched_step_kernel);
        // the first poll after this returns nothing, so a count read before the flag landed is never used
        static const int one = 1;
        MLI_CUDA(cudaMemcpyAsync(&e->a.v->poll, &one, sizeof(int), cudaMemcpyHostToDevice, e->io));
        MLI_CUDA(cudaStreamSynchronize(e->io));
        e->poll_armed = true;
        *n_out = 0;
        offsets_out[0] = 0;
        return MLI_OK;
    }
    int n_fin = 0;
    int rc0 = finished_count(e, &n_fin);
    if (rc0) return rc0;
    int n = std::min(std::max(n_fin - e->n_polled, 0), std::max(max_out, 0));
    *n_out = 0;
    offsets_out[0] = 0;
    if (n == 0) return MLI_OK;
    const int *ids, *offs, *toks;
    int rc = gather_finished(e, e->n_polled, n, &ids, &offs, &toks);
    if (rc) return rc;
    while (n > 0 && (long long)offs[n] > tokens_capacity) --n;   // hand out only what fits
    memcpy(ids_out, ids, sizeof(int) * (size_t)n);
    memcpy(offsets_out, offs, sizeof(int) * ((size_t)n + 1));
    memcpy(tokens_out, toks, sizeof(int) * (size_t)offs[n]);
    e->n_polled += n;
    *n_out = n;
    return MLI_OK;
}

int mli_engine_copy_tokens(mli_engine* e, int* tokens_dev, int* counts_dev) {
    MLI_REQUIRE(e && tokens_dev && counts_dev, "null argument");
    MLI_ENTER(e->ctx, "null ctx");
    mli_ctx* ctx = e->ctx;
    MLI_CUDA(cudaStreamSynchronize(e->stream));
    MLI_CUDA(cudaMemcpyAsync(tokens_dev, e->a.req_tok,
                             sizeof(int) * (size_t)e->n_req * e->cfg.n_sequence,
                             cudaMemcpyDeviceToDevice, ctx->stream));
    MLI_CUDA(cudaMemcpyAsync(counts_dev, e->a.req_cnt, sizeof(int) * (size_t)e->n_req,
                             cudaMemcpyDeviceToDevice, ctx->stream));
    return MLI_OK;
}

int mli_engine_get_stats(mli_engine* e, mli_engine_stats* stats) {
    MLI_REQUIRE(e && stats, "null argument");
    *stats = e->stats;
    return MLI_OK;
}

}  // extern "C"

// internal to the library (comm.cu): what the token gather sends
namespace mli {
int engine_token_table(mli_engine* e, const int** tokens_dev, const int** counts_dev, int* capacity,
                       int* n_sequence, cudaStream_t* engine_stream, mli_ctx** ctx) {
    MLI_REQUIRE(e, "null engine");
    if (tokens_dev) *tokens_dev = e->a.req_tok;
    if (counts_dev) *counts_dev = e->a.req_cnt;
    if (capacity) *capacity = e->cfg.max_requests;
    if (n_sequence) *n_sequence = e->cfg.n_sequence;
    if (engine_stream) *engine_stream = e->stream;
    if (ctx) *ctx = e->ctx;
    return MLI_OK;
}
}  // namespace mli
