/*
 * oracle.h -- CPU restatement (plain C) of the reference's paged-attention decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (min_llm_inference_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker or the reported CPU baseline.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.  The
 * reference has NO golden vectors; the oracle is pinned two ways:
 *   (1) on a GPU box, tests/test_gpu_oracle_pin.py runs the reference's own CUDA kernels
 *       (oracle/_ref/libmli_ref.so, compiled from the sources where they lie) on the same seeded
 *       inputs and compares: K/V/q/logits/tokens must be BIT-EXACT (both sides are k-ascending
 *       single-accumulator fp32 FMA chains), softmax outputs within 2e-6 rel (expf differs);
 *   (2) tests/golden/ holds outputs of that reference CUDA path generated on a B200 by
 *       tests/golden/make_golden.py; the CPU-only suite checks the oracle against them.
 *
 * Conventions (reference include/utils.h:32-76, include/constants.h:12-18):
 *   page           = float[16][3][d]; sub-row 0 = input embedding, 1 = K, 2 = V
 *   page_table     = float*[B][W], W = S/16, raw pointers (host pointers here)
 *   element(r,j,off,c) = page_table[r*W + j/16][(j%16)*3*d + off*d + c]
 *   lengths[r] == 0 marks an empty row; EOF token = 1023; empty-row token = -1
 */
#ifndef MLI_ORACLE_H
#define MLI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_PAGE_BLOCK 16
#define ORC_EOF_TOKEN 1023
#define ORC_EMPTY_TOKEN (-1)
#define ORC_INIT_BLOCKS 4

/* number of threads used by the row-parallel loops (1 = the reference's own single-threaded
 * host behaviour); returns the value actually in effect (OpenMP may be compiled out). */
int orc_set_threads(int n);
int orc_get_threads(void);

/* ---- paged stage functions ------------------------------------------------------------ */
void orc_paged_encoder(const float* emb_table, const float* pos_table, const int* inp,
                       float** page_table, const int* lengths, const int* new_item_indices,
                       int n_batch, int n_sequence, int emb_dim, int n_new_items);

void orc_prefill_kv_paged(float** page_table, const int* new_batch_idx, const int* lengths,
                          const float* wk, const float* wv, int n_new_items, int n_sequence,
                          int emb_dim);

void orc_qkv_latest_paged(float** page_table, const int* lengths, const float* wk,
                          const float* wq, const float* wv, float* q_output, int n_batch,
                          int n_sequence, int emb_dim);

void orc_qkt_paged(const float* q, float** page_table, const int* lengths, float* qkt,
                   int n_batch, int n_sequence, int emb_dim);

void orc_softmax_in_place_with_lengths(float* qkt, const int* lengths, int n_batch,
                                       int n_sequence);

void orc_softmax_v_paged(const float* softmax_result, float** page_table,
                         float* attention_result, const int* lengths, int n_batch,
                         int n_sequence, int emb_dim);

/* a15: prefill -> latest -> qkt -> softmax -> softmax_v */
void orc_paged_attention(float** page_table, const int* lengths, const float* wk,
                         const float* wq, const float* wv, const int* new_batch_idx,
                         float* q_output, float* qkt_output, float* attention_result,
                         int n_new_items, int n_batch, int n_sequence, int emb_dim);

/* logits[B,V] = attn[B,d] . emb_table[V,d]^T */
void orc_logits(const float* batch_result, const float* emb_table, float* emb_score,
                int n_batch, int n_vocab, int emb_dim);

/* device argmax rule (decoder.cu:146-172): winner = min over maxima of
 * (bitreverse8(index mod 256), index) -- the tree keeps the lower thread at every level */
int orc_argmax_device_rule(const float* score, int n_vocab);

void orc_paged_decoder(const float* emb_score, int* decoder_result, int* lengths,
                       float** page_table, const float* pos_table, const float* emb_table,
                       int n_batch, int n_vocab, int n_sequence, int emb_dim,
                       int n_decoder_results, int i_decoder);

/* a7: R x (encoder -> attention -> logits+decoder); scratch may be NULL (allocated inside) */
void orc_paged_forward(const int* inp, int* lengths, const int* new_item_indices,
                       int* decoder_result, int n_new_items, const float* emb_table,
                       const float* pos_table, float** page_table, const float* wk,
                       const float* wq, const float* wv, float* attention_result /*[B,d] or NULL*/,
                       float* emb_score /*[B,V] or NULL*/, int n_batch, int n_sequence,
                       int emb_dim, int n_vocab, int n_forward_rounds);

/* ---- dense (non-paged) functions: config C1 -------------------------------------------- */
void orc_dense_encoder(const float* emb_table, const float* pos_table, const int* inp,
                       float* inp_embedding, const int* lengths, const int* new_item_indices,
                       int n_batch, int n_sequence, int emb_dim, int n_new_items);

void orc_self_attention(const float* inp_embedding, const int* lengths, const float* wk,
                        const float* wq, const float* wv, const int* new_batch_idx,
                        float* kt_cache, float* v_cache, float* q_output, float* qkt_output,
                        float* attention_result, int n_new_items, int n_batch, int n_sequence,
                        int input_dim, int output_dim);

void orc_dense_decoder(const float* emb_score, int* decoder_result, int* lengths,
                       float* inp_embedding, const float* pos_table, const float* emb_table,
                       int n_batch, int n_vocab, int n_sequence, int emb_dim);

void orc_dense_forward(const int* inp, int* lengths, const int* new_item_indices,
                       int* decoder_result, int n_new_items, const float* emb_table,
                       const float* pos_table, const float* wk, const float* wq,
                       const float* wv, float* inp_embedding, float* kt_cache, float* v_cache,
                       int n_batch, int n_sequence, int emb_dim, int n_vocab);

/* ---- engines (continuous batching loops) ------------------------------------------------ */
typedef struct {
    int n_batch, n_sequence, emb_dim, n_vocab;
    int n_blocks;          /* KV pages in the pool (paged only) */
    int n_forward_rounds;  /* paged only; 1..16 */
    int fix_stale_lengths; /* 0 = reproduce reference quirk Q1 (paged_item_storage.cpp:62-118),
                              1 = refresh host lengths from the device before every insert */
    int max_steps;         /* safety cap on loop iterations; <=0 = unlimited */
    /* opt-in scheduling policies of the product's engine that the reference does not have (0 = off =
     * the reference's behaviour); restated here so scheduler decisions stay comparable */
    int max_new_tokens;        /* > 0: a request is finished once it has generated this many tokens */
    int max_prefill_positions; /* > 0: one insert_new_items call admits prompts only while their
                                  positions add up to at most this many (the first always passes) */
    int prefill_chunk_positions; /* > 0: chunked prefill -- a step schedules at most this many prompt positions
                                  (granules of 16, rows in admission order); an admitted row becomes active in
                                  the step that schedules its last chunk.  Needs fix_stale_lengths = 1 */
} orc_engine_cfg;

typedef struct {
    long long steps;            /* loop iterations executed */
    long long generated_tokens; /* tokens appended by process_decoder_result */
    long long preemptions;
    int n_finished;
} orc_engine_stats;

/* Requests are given as prompt_offsets[n_req+1] into prompt_tokens; request id = index.
 * Finished requests are returned in finish order: finished_ids[i], tokens
 * finished_tokens[finished_offsets[i] .. finished_offsets[i+1]) (prompt + generated).
 * finished_tokens must hold n_req * n_sequence ints.  Returns 0, or <0 on error
 * (-2 = "No enough block memories to return"). */
int orc_paged_engine_run(const orc_engine_cfg* cfg, const float* emb_table,
                         const float* pos_table, const float* wk, const float* wq,
                         const float* wv, int n_req, const int* prompt_offsets,
                         const int* prompt_tokens, int* finished_ids, int* finished_offsets,
                         int* finished_tokens, orc_engine_stats* stats);

int orc_dense_engine_run(const orc_engine_cfg* cfg, const float* emb_table,
                         const float* pos_table, const float* wk, const float* wq,
                         const float* wv, int n_req, const int* prompt_offsets,
                         const int* prompt_tokens, int* finished_ids, int* finished_offsets,
                         int* finished_tokens, orc_engine_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
