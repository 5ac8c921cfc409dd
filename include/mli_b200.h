/*
 * mli_b200.h -- C ABI of the B200-native (sm_100a) paged-attention decode path that drops in for
 * xyg-coder/min_llm_inference.  Plain pointers and sizes only; every pointer named *_dev / every
 * tensor argument is a DEVICE pointer unless the comment says HOST.  All functions return 0 on
 * success and a negative mli_status on failure; mli_last_error() gives the text.  Nothing here
 * falls back to the CPU: without a CUDA device every compute entry point fails with MLI_ERR_CUDA.
 *
 * Each entry point names the reference interface it replaces (file:line under /root/reference).
 * The C++ mirror of the reference's classes (Tensor, *Layer, *InferenceModel, start_*_engine ...)
 * lives in min_llm_inference_b200/host/ and is implemented on top of exactly these calls.
 *
 * Data contracts (reference include/utils.h:32-76, include/constants.h:3-18):
 *   page        float[16][3][d]   sub-row 0 = input embedding, 1 = K, 2 = V
 *   page_table  float*[B][S/16]   raw device pointers; entries past a row's allocation are never read
 *   lengths     int[B]            0 = empty row; mutated by the decoder stage
 *   weights     float[d][d]       row-major [in, out] (y = x . W)
 *   emb_table   float[V][d], pos_table float[S][d]
 *   tokens      EOF = 1023, empty-row marker = -1
 *   d % 4 == 0, S % 16 == 0, 1 <= n_forward_rounds <= 16
 */
#ifndef MLI_B200_H
#define MLI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLI_PAGE_BLOCK_SIZE 16
#define MLI_EOF_TOKEN_ID 1023
#define MLI_EMPTY_ROW_TOKEN_ID (-1)
#define MLI_DEFAULT_INIT_NUM_BLOCKS 4

typedef enum {
    MLI_OK = 0,
    MLI_ERR_CUDA = -1,      /* CUDA runtime/driver failure ("Cuda Failure" in the reference, src/utils.cpp:5-11) */
    MLI_ERR_ARG = -2,       /* contract violation (the reference asserts) */
    MLI_ERR_NO_BLOCKS = -3, /* "No enough block memories to return" (src/paged_item_storage.cpp:145-147) */
    MLI_ERR_UNSUPPORTED = -4,
    MLI_ERR_STATE = -5
} mli_status;

typedef struct mli_ctx mli_ctx; /* opaque: stream, SM count, workspaces, cached TMA descriptors */

typedef enum {
    /* how the dense contractions (latest-token QKV, prefill, logits) are computed */
    MLI_OPT_GEMM_MODE = 1,        /* 0 = tcgen05 3xTF32 (tensor cores, TMEM accumulators; default when
                                         available), 1 = SIMT fp32 with the reference's k-ascending
                                         FMA order (bit-exact with the reference's naive kernels) */
    MLI_OPT_ATTN_CHUNK_PAGES = 2, /* KV pages per split of the fused decode attention; 0 = auto */
    MLI_OPT_ATTN_CTAS_PER_SM = 3, /* persistent CTAs per SM for the fused decode attention; 0 = auto */
    MLI_OPT_PDL = 4,              /* 1 (default) = the engine's step graph chains its kernels with
                                         programmatic dependent launch (each kernel's prologue overlaps
                                         the tail of its predecessor); 0 = plain stream order */
    MLI_OPT_KV_FORMAT = 5         /* 0 (default) = the reference's page float[16][3][d]; 1 = compact page
                                         (opt-in because the layout is API-visible): a position
                                         is [inp f32 x d | K bf16 x d | V bf16 x d] = 8*d bytes, i.e. a page
                                         is 16*2*d floats.  K and V are rounded to bf16 (rel 2^-9) when they
                                         are written; q, scores, softmax and accumulation stay fp32.  Only the
                                         tensor-core GEMM mode and the single-launch attention support it. */
    ,
    MLI_OPT_ATTN_KERNEL = 6       /* consumer design of the single-launch decode attention: 0 (default) =
                                         auto, 1 = column-split consumers (two CTAs per SM, a position's
                                         columns spread over 8 warps), 2 = warp-per-position consumers (one
                                         CTA per SM, 16 warps with private online-softmax states; emb_dim in
                                         {128, 256, 512, 1024, 2048, 4096}) */
    ,
    MLI_OPT_ATTN_MIN_DYN = 7      /* positions per attention CTA (fair share) from which the last quarter of a
                                         launch is handed out in small dynamic slices; default 4096 (below
                                         that the merges of the cut rows cost more than the tail they
                                         remove).  Tests lower it to cover the dynamic path on small inputs */
} mli_option;

/* ---- context --------------------------------------------------------------------------- */
int mli_ctx_create(mli_ctx** out, int device, void* cuda_stream /* cudaStream_t, NULL = default */);
int mli_ctx_destroy(mli_ctx* ctx);
int mli_ctx_set_stream(mli_ctx* ctx, void* cuda_stream);
int mli_ctx_set_option(mli_ctx* ctx, int option, int value);
int mli_ctx_get_option(mli_ctx* ctx, int option, int* value);
int mli_ctx_synchronize(mli_ctx* ctx);
/* tcgen05 mode keeps tf32-split, K-major copies of the weight operands.  Registered pointers are
 * split once and trusted to stay unchanged until unregistered (what a layer that owns its weights
 * does: the reference's *Layer classes take them by &&, include/layers.h:54-100); unregistered
 * pointers are re-split on every call.  Any argument may be NULL. */
int mli_ctx_register_weights(mli_ctx* ctx, const float* wk, const float* wq, const float* wv,
                             const float* emb_table, int emb_dim, int n_vocab);
int mli_ctx_unregister_weights(mli_ctx* ctx);
const char* mli_last_error(void);
const char* mli_version(void);
/* number of kernels this library has launched in this process (bench.py reports it) */
long long mli_kernel_launch_count(void);
/* diagnostics (tools/gemm_timing.py): when stamps_dev != NULL every tcgen05 GEMM launch writes
 * clock64() phase stamps [cta][8] into it (>= 8 * 8 * n_ctas bytes); NULL switches it off */
int mli_debug_set_gemm_stamps(mli_ctx* ctx, void* stamps_dev);
/* diagnostics (tests): the plan of this context's last tcgen05 GEMM launch of one kind (0 latest-token QKV
 * stage, 1 prefill stage, 2 logits, 3 the engine's merged QKV + prefill step) into HOST int[5]:
 * [0] kernel: 0 static grid, 1 persistent with dynamic tiles, 2 persistent on CTA pairs (cta_group::2), -1 none
 * yet; [1] K split; [2] activation-tile width; [3] decode-tile width of the pair kernel (0 = same); [4] K passes
 * (emb_dim 4096 on CTA pairs: 2) */
int mli_debug_last_gemm_plan(mli_ctx* ctx, int kind, int* plan);
/* diagnostics (tools/step_timeline.py): when trace_dev != NULL (u64[8 + 8 * capacity], [0] = 0,
 * [1] = capacity in steps) the kernels of every engine step record the %globaltimer at which their
 * dependencies were satisfied; must be set before the engine captures its step graph */
int mli_debug_set_step_trace(mli_ctx* ctx, void* trace_dev);

/* ---- paged stages ------------------------------------------------------------------------ */
/* replaces launch_paged_attention_encoder_kernel (include/kernels/encoder.h:21-25,
 * src/kernels/encoder.cu:102-147): new rows, j < L: page[j].inp = E[tok_j] + P[j] */
int mli_paged_encoder(mli_ctx* ctx, const float* emb_table, const float* pos_table, const int* inp,
                      float** page_table, const int* lengths, const int* new_item_indices,
                      int n_batch, int n_sequence, int emb_dim, int n_new_items);

/* replaces launch_fill_new_k_v_cache_paged_attention{,_warp_tiling}
 * (include/kernels/paged_attention.h:28-30, :65-67; src/kernels/paged_attention.cu:20-115,
 * src/kernels/paged_attention_cublas.cu:112-246): prefill K,V of new rows into their pages */
int mli_prefill_kv_paged(mli_ctx* ctx, float** page_table, const int* new_batch_idx,
                         const int* lengths, const float* wk, const float* wv, int n_new_items,
                         int n_batch, int n_sequence, int emb_dim);

/* replaces launch_get_latest_k_q_v_paged_attention{,_cublas} (paged_attention.h:32-35, :57-63;
 * paged_attention.cu:126-199, paged_attention_cublas.cu:16-99): x = page[L-1].inp;
 * k,v -> same page, q -> q_output[B,d]; rows with L == 0 untouched */
int mli_qkv_latest_paged(mli_ctx* ctx, float** page_table, const int* lengths, const float* wk,
                         const float* wq, const float* wv, float* q_output, int n_batch,
                         int n_sequence, int emb_dim);

/* replaces launch_qkt_paged_attention + launch_softmax_in_place_with_lengths +
 * launch_softmax_v_paged_attention (paged_attention.h:37-43; paged_attention.cu:208-345;
 * self_attention_inference_optimized.cu:191-242) with ONE fused, length-aware, split-KV kernel.
 * attention_result[B,d]; rows with L == 0 get zeros.  softmax_out may be NULL; when given it
 * receives the [B,S] probabilities the reference leaves in qkt_output (zeros past L). */
int mli_decode_attention_paged(mli_ctx* ctx, const float* q, float* const* page_table,
                               const int* lengths, float* attention_result, float* softmax_out,
                               int n_batch, int n_sequence, int emb_dim);

/* The reference's three unfused stages, one launch each, for callers (and the reference's kernel tests,
 * tests/paged_attention_kernels_test.cpp:115-169) written against them.  Not used by the product path.
 * They keep the reference's arithmetic order, so scores, probabilities and P.V are bit-identical to its kernels.
 * replaces launch_qkt_paged_attention (paged_attention.h:37-39; paged_attention.cu:208-280):
 * qkt_output[r][j] = (q[r] . K[r][j]) / sqrtf(d) for j < lengths[r]; other entries untouched */
int mli_qkt_paged(mli_ctx* ctx, const float* q, float* const* page_table, const int* lengths,
                  float* qkt_output, int n_batch, int n_sequence, int emb_dim);
/* replaces launch_softmax_in_place_with_lengths (self_attention_inference_optimized.h:21-22;
 * self_attention_inference_optimized.cu:191-242, :360-368): softmax over the first lengths[r]
 * entries of every row, zeros up to n_sequence */
int mli_softmax_in_place_with_lengths(mli_ctx* ctx, float* qkt_output, const int* lengths, int n_batch,
                                      int n_sequence);
/* replaces launch_softmax_v_paged_attention (paged_attention.h:41-43; paged_attention.cu:287-345):
 * attention_result[r][c] = sum_{j < lengths[r]} softmax_result[r][j] * V[r][j][c] */
int mli_softmax_v_paged(mli_ctx* ctx, const float* softmax_result, float* const* page_table,
                        float* attention_result, const int* lengths, int n_batch, int n_sequence,
                        int emb_dim);

/* replaces paged_attention / paged_attention_with_cublas (paged_attention.h:17-26, :46-55;
 * paged_attention.cu:358-377): prefill(new rows) -> latest QKV -> fused decode attention.
 * qkt_output may be NULL (the fused kernel needs no [B,S] scratch). */
int mli_paged_attention(mli_ctx* ctx, float** page_table, const int* lengths, const float* wk,
                        const float* wq, const float* wv, const int* new_batch_idx, float* q_output,
                        float* qkt_output, float* attention_result, int n_new_items, int n_batch,
                        int n_sequence, int emb_dim);

/* replaces launch_paged_attention_decoder_multi_rounds / ..._cublas_... (include/kernels/decoder.h:27-37;
 * src/kernels/decoder.cu:128-255): logits = attn . E^T, device-rule argmax, token ->
 * decoder_result[r*n_decoder_results + i_decoder], lengths update, next embedding into page[L].
 * emb_score[B,V] may be NULL (logits then stay in an internal workspace). */
int mli_paged_decoder(mli_ctx* ctx, const float* batch_result, const float* emb_table,
                      float* emb_score, const float* pos_table, float** page_table, int* lengths,
                      int* decoder_result, int n_batch, int n_vocab, int n_sequence, int emb_dim,
                      int n_decoder_results, int i_decoder);

/* replaces PagedAttentionInferenceModel::forward / PagedAttentionCublasInferenceModel::forward
 * (include/inference_model.h:34-74, src/inference_model.cpp:52-124): n_forward_rounds x
 * (encoder -> attention -> decoder); new rows only in round 0.  attention_result[B,d] and
 * q_output[B,d] are caller scratch (the reference's layer-owned tensors); may be NULL. */
int mli_paged_forward(mli_ctx* ctx, const int* inp, int* lengths, const int* new_item_indices,
                      int* decoder_result, int n_new_items, const float* emb_table,
                      const float* pos_table, float** page_table, const float* wk, const float* wq,
                      const float* wv, float* q_output, float* attention_result, int n_batch,
                      int n_sequence, int emb_dim, int n_vocab, int n_forward_rounds);

/* ---- dense (non-paged) stages: BASELINE config C1 ------------------------------------------ */
/* replaces launch_inference_optimized_encoder_kernel (encoder.h:16-19; encoder.cu:56-92) */
int mli_dense_encoder(mli_ctx* ctx, const float* emb_table, const float* pos_table, const int* inp,
                      float* inp_embedding, const int* lengths, const int* new_item_indices,
                      int n_batch, int n_sequence, int emb_dim, int n_new_items);

/* replaces inference_self_attention (include/kernels/self_attention_inference_optimized.h:37-47;
 * self_attention_inference_optimized.cu:27-383).  kt_cache is TRANSPOSED [B,d_out,S], v_cache
 * [B,S,d_out].  qkt_output[B,S] receives the softmax probabilities (zeros past L) as in the
 * reference; may be NULL. */
int mli_self_attention(mli_ctx* ctx, const float* inp_embedding, const int* lengths, const float* wk,
                       const float* wq, const float* wv, const int* new_batch_idx, float* kt_cache,
                       float* v_cache, float* q_output, float* qkt_output, float* attention_result,
                       int n_new_items, int n_batch, int n_sequence, int input_dim, int output_dim);

/* replaces launch_decoder (decoder.h:19-24; decoder.cu:25-112) */
int mli_dense_decoder(mli_ctx* ctx, const float* batch_result, const float* emb_table,
                      float* emb_score, const float* pos_table, float* inp_embedding, int* lengths,
                      int* decoder_result, int n_batch, int n_vocab, int n_sequence, int emb_dim);

/* replaces InferenceModel::forward (inference_model.h:8-30; inference_model.cpp:14-39) */
int mli_dense_forward(mli_ctx* ctx, const int* inp, int* lengths, const int* new_item_indices,
                      int* decoder_result, int n_new_items, const float* emb_table,
                      const float* pos_table, const float* wk, const float* wq, const float* wv,
                      float* inp_embedding, float* kt_cache, float* v_cache, float* q_output,
                      float* attention_result, int n_batch, int n_sequence, int emb_dim, int n_vocab);

/* ---- on-device continuous-batching engine --------------------------------------------------- */
/* replaces start_paged_attention_inference_engine / ..._cublas_... (include/inferencer.h:23-32;
 * src/inferencer.cpp:43-133) together with the host scheduler it drives: paged insert_new_items,
 * allocate_or_free_memory_blocks_if_needed, process_decoder_result, MemoryBlockManager and
 * PagedAttentionsManager (src/paged_item_storage.cpp:14-203, src/item_storage.cpp:97-139).
 * Requests, the page free list, the page table and all admission / retirement / growth /
 * pre-emption decisions live on the device; the host only launches a CUDA graph per step and
 * polls a pinned completion word.
 * (SURVEY 8b calls these entry points mli_sched_*: create / admit / step / poll / destroy map to
 * mli_engine_create / _submit + _enqueue / _run / _poll_finished + _results / _destroy.) */
typedef struct {
    int n_batch, n_sequence, emb_dim, n_vocab;
    int n_blocks;          /* KV pages in the pool */
    int n_forward_rounds;  /* 1..16 */
    int compat_stale_lengths; /* 1 = reproduce the reference's stale-lengths behaviour
                                 (paged_item_storage.cpp:73-75,:114-118) so
                                 token lists match the reference engine; 0 = corrected */
    int max_requests;      /* capacity of the device request table */
    float* page_pool;      /* optional caller-owned slab of n_blocks*16*3*d floats (e.g. the
                              reference's MemoryBlockManager slab); NULL = engine allocates */
    /* opt-in scheduling policies that the reference does not have (0 = off = the reference's
     * behaviour); the CPU oracle implements the same two flags so decisions stay comparable */
    int max_new_tokens;        /* > 0: a request is finished once it has generated this many tokens
                                  (besides EOF / n_sequence, src/item_storage.cpp:120-124) */
    int max_prefill_positions; /* > 0: admission throttle: one step admits queued prompts
                                  only while their positions add up to at most this many (the first
                                  admission of a step always passes), so an admission burst is spread
                                  over several steps instead of stalling every decoding row behind one
                                  huge prefill GEMM (the reference admits everything that fits,
                                  src/paged_item_storage.cpp:84-113) */
    int prefill_chunk_positions; /* > 0: chunked prefill: a step prefills at most this many prompt
                                  positions (a multiple of 16) over all rows whose prompts are still being
                                  prefilled, in admission order; an admitted row stays inactive until the step
                                  that schedules its last chunk, in which it also emits its first token -- the
                                  decoding rows are never stalled behind one huge prefill GEMM (the reference
                                  prefills a whole prompt in the step that admits it, src/inference_model.cpp:
                                  52-82).  Tokens do not depend on the chunking; needs compat_stale_lengths = 0.
                                  Tensor-core mode computes the chunks step by step; exact-order mode keeps the
                                  same schedule but prefills a row when it becomes active */
} mli_engine_cfg;

typedef struct {
    long long steps;            /* engine iterations (each = n_forward_rounds decode rounds) */
    long long generated_tokens; /* tokens appended to requests */
    long long preemptions;
    long long admitted;
    int n_finished;
    float gpu_ms;               /* device time (CUDA events on the engine's stream) from the start
                                   of the mli_engine_submit that fed it to the end of the last
                                   mli_engine_run */
    float attn_ms;              /* device time spent in the fused decode-attention kernels, only
                                   when profiling was requested (else 0) */
    double attn_bytes;          /* algorithmic bytes of those launches (SURVEY 8d ATTN_BYTES) */
    long long attn_launches;
    float gemm_ms;              /* profiling only: device time of the merged latest-QKV + prefill GEMM */
    double gemm_flops;          /* fp32-equivalent FLOPs of those launches: 6*d*d per active row +
                                   4*d*d per prefill position (the tensor cores execute 3x that in tf32) */
    long long gemm_launches;
    double gemm_max_flops;      /* profiling only: the largest merged-GEMM launch of the job (the bulk
                                   prefill, i.e. the tensor-pipe regime) and its device time */
    float gemm_max_ms;
    int peak_resident_rows;     /* most rows occupied at once */
    int min_free_pages;         /* fewest free KV pages seen (pool occupancy = 1 - min_free / n_blocks) */
} mli_engine_stats;

typedef struct mli_engine mli_engine;

int mli_engine_create(mli_ctx* ctx, const mli_engine_cfg* cfg, const float* emb_table,
                      const float* pos_table, const float* wk, const float* wq, const float* wv,
                      mli_engine** out);
int mli_engine_destroy(mli_engine* e);
/* upload requests; prompt_offsets[n_req+1] / prompt_tokens are HOST pointers (is_device = 0) or
 * DEVICE pointers (is_device = 1, used by bench.py's device-resident leg).  Resets the engine. */
int mli_engine_submit(mli_engine* e, int n_req, const int* prompt_offsets, const int* prompt_tokens,
                      int is_device);
/* streaming ingestion (replaces ItemStorage::add_new_item being called while the loop of
 * src/inferencer.cpp:43-85 runs): append requests to a live engine WITHOUT resetting it.  Ids continue
 * from the last submitted / enqueued request (*first_id, optional, receives the first new id).  The
 * upload runs on the engine's ingest stream, concurrent with the step graphs; the device scheduler
 * queues the new requests at its next iteration.  May be called from another host thread while
 * mli_engine_run executes; requests that arrive after the engine has gone idle are processed by the
 * next mli_engine_run.  Fails when max_requests would be exceeded. */
int mli_engine_enqueue(mli_engine* e, int n_req, const int* prompt_offsets, const int* prompt_tokens,
                       int is_device, int* first_id);
/* run to completion (max_steps <= 0) or for at most max_steps iterations.
 * profile_attention != 0 brackets every fused-attention launch with CUDA events (no graph).
 * Returns MLI_ERR_NO_BLOCKS when the job cannot complete: a prompt that needs more pages than the pool
 * holds, or a pre-empted request that has outgrown the pool (nothing resident, every page free and the
 * queue head still not admissible).  The reference's loop (src/inferencer.cpp:43-85 over
 * src/paged_item_storage.cpp:84-113) spins for ever in the second case; here the job ends, the requests
 * that finished are available from mli_engine_results and the engine accepts the next mli_engine_submit. */
int mli_engine_run(mli_engine* e, long long max_steps, int profile_attention);
/* download finished requests in finish order (HOST buffers): ids[n_req], offsets[n_req+1],
 * tokens[n_req * n_sequence] (prompt + generated) */
int mli_engine_results(mli_engine* e, int* finished_ids, int* finished_offsets, int* finished_tokens,
                       int* n_finished);
/* asynchronous token return (replaces process_decoder_result's blocking cudaMemcpy +
 * ItemStorage::pop_finished_items, src/item_storage.cpp:97-139): the requests that finished since the
 * last poll, in finish order, HOST buffers: ids[max_out], offsets[max_out+1], tokens[tokens_capacity]
 * (prompt + generated, packed).  Never waits for the engine: the finished count is read and the token
 * lists are packed into mapped pinned memory on a side stream while the step graphs keep running.  Callable from another host thread during mli_engine_run.  *n_out may be 0; it always is on
 * the first call after a submit, which only tells the device scheduler that polling is in use (from then on
 * it orders a step's finished lists before the count it publishes; a job that never polls does not pay for
 * that fence). */
int mli_engine_poll_finished(mli_engine* e, int max_out, int* ids, int* offsets, int* tokens,
                             long long tokens_capacity, int* n_out);
/* device-to-device copy of the request table (tokens[n_req][n_sequence], counts[n_req]) into
 * caller buffers on the context's stream -- what the multi-GPU token gather (NCCL all-gather,
 * the only collective of the path) sends */
int mli_engine_copy_tokens(mli_engine* e, int* tokens_dev, int* counts_dev);
int mli_engine_get_stats(mli_engine* e, mli_engine_stats* stats);

/* ---- multi-GPU: request sharding + the final token gather ------------------------------------- */
/* The reference is single-GPU (no collective anywhere under /root/reference); its engine
 * loop (include/inferencer.h:23-32) shards by request with no exchange inside a decode step, so each
 * GPU runs its own mli_engine on its share of the requests and ONE collective -- an NCCL all-gather of
 * the per-rank request tables over NVLink / NVSwitch -- returns every finished token list to every
 * rank.  NCCL is dlopen'ed on first use (libnccl.so.2); single-GPU callers never need it.
 *   multi-process (one process per GPU, e.g. torchrun): rank 0 calls mli_comm_get_unique_id, ships the
 *     MLI_COMM_ID_BYTES bytes to the other ranks by any means, every rank calls mli_comm_init_rank;
 *   single process driving n GPUs: mli_comm_init_all over one context per device; collective calls of
 *     the n communicators then go between mli_comm_group_start / mli_comm_group_end. */
#define MLI_COMM_ID_BYTES 128
typedef struct mli_comm mli_comm;
int mli_comm_get_unique_id(void* id_out, size_t id_bytes);
int mli_comm_init_rank(mli_ctx* ctx, int world_size, int rank, const void* unique_id, mli_comm** out);
int mli_comm_init_all(mli_ctx* const* ctxs, int n, mli_comm** comms_out /* [n] */);
int mli_comm_group_start(void);
int mli_comm_group_end(void);
/* all-gather the engine's request table: all_tokens_dev[world*per_rank][n_sequence] and
 * all_counts_dev[world*per_rank] (DEVICE buffers of the caller); rank r's requests land at rows
 * [r*per_rank, (r+1)*per_rank), count 0 = slot unused.  per_rank <= the engine's max_requests.
 * Stream-ordered on the context's stream after the engine's own stream; no host synchronisation. */
int mli_comm_gather_tokens(mli_comm* comm, mli_engine* e, int per_rank, int* all_tokens_dev,
                           int* all_counts_dev);
int mli_comm_info(mli_comm* comm, int* world_size, int* rank, int* nccl_version);
int mli_comm_destroy(mli_comm* comm);

#ifdef __cplusplus
}
#endif
#endif
