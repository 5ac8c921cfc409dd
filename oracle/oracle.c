/*
 * oracle.c -- CPU restatement of the reference's paged-attention decode path.  See oracle.h.
 *
 * TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline); never linked into the product.
 *
 * Arithmetic contract: every dot product in the reference's CUDA path is a k-ascending, single
 * accumulator chain of fp32 FMAs starting from 0 (nvcc contracts `acc += a*b` to FFMA), e.g.
 * src/kernels/paged_attention.cu:64-66, :167-171, :254-256, :318-320 and src/kernels/gemm.cu:42-44.
 * This file reproduces those chains with fmaf() in the same order, so K/V/q/qkt/logits are
 * bit-identical to the reference's naive CUDA kernels.  expf differs between libm and CUDA, so
 * softmax (and what follows it) agrees to ~1e-6 relative, not bitwise.
 *
 * Build: gcc -O3 -mfma -mavx2 -ffp-contract=off -fno-math-errno -fopenmp (see oracle/Makefile).
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

int orc_set_threads(int n) {
    if (n < 1) n = 1;
#ifdef _OPENMP
    g_threads = n;
#else
    g_threads = 1;
#endif
    return g_threads;
}
int orc_get_threads(void) { return g_threads; }

static inline int ceil_div_i(int a, int b) { return (a + b - 1) / b; }

/* include/utils.h:32-60: address of element (r, j, off, 0) */
static inline float* page_row(float** page_table, int r, int W, int j, int d, int off) {
    float* page = page_table[(size_t)r * W + j / ORC_PAGE_BLOCK];
    return page + (size_t)(j % ORC_PAGE_BLOCK) * d * 3 + (size_t)off * d;
}

/* y[c] = fma(x, w[c], y[c]) for c < n : one k-step of n independent chains */
static inline void axpy_fma(float x, const float* __restrict__ w, float* __restrict__ y, int n) {
    for (int c = 0; c < n; ++c) y[c] = __builtin_fmaf(x, w[c], y[c]);
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/encoder.cu:102-147 (paged_attention_encoder); host twin tests/test_utils.cpp:559-573
 * for new rows, j < L: page[r][j].inp = emb_table[tok_j] + pos_table[j]
 * ---------------------------------------------------------------------------------------- */
void orc_paged_encoder(const float* emb_table, const float* pos_table, const int* inp,
                       float** page_table, const int* lengths, const int* new_item_indices,
                       int n_batch, int n_sequence, int emb_dim, int n_new_items) {
    (void)n_batch;
    int W = n_sequence / ORC_PAGE_BLOCK;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int i = 0; i < n_new_items; ++i) {
        int r = new_item_indices[i];
        int L = lengths[r];
        for (int j = 0; j < L; ++j) {
            int tok = inp[(size_t)r * n_sequence + j];
            const float* e = emb_table + (size_t)tok * emb_dim;
            const float* p = pos_table + (size_t)j * emb_dim;
            float* x = page_row(page_table, r, W, j, emb_dim, 0);
            for (int c = 0; c < emb_dim; ++c) x[c] = e[c] + p[c];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/paged_attention.cu:20-87 (fill_new_k_v_cache_paged_attention); host twin
 * tests/test_utils.cpp:29-66.  K[j] = x[j].Wk, V[j] = x[j].Wv for new rows, j < L, written to
 * sub-rows 1 and 2 of the same page.  k-ascending FMA chains.
 * ---------------------------------------------------------------------------------------- */
void orc_prefill_kv_paged(float** page_table, const int* new_batch_idx, const int* lengths,
                          const float* wk, const float* wv, int n_new_items, int n_sequence,
                          int emb_dim) {
    int W = n_sequence / ORC_PAGE_BLOCK;
    int d = emb_dim;
    /* flatten (new row, position) so threads balance over ragged prompts */
    long long total = 0;
    for (int i = 0; i < n_new_items; ++i) total += lengths[new_batch_idx[i]];
    if (total == 0) return;
    int* rows = (int*)malloc(sizeof(int) * (size_t)total);
    int* poss = (int*)malloc(sizeof(int) * (size_t)total);
    long long t = 0;
    for (int i = 0; i < n_new_items; ++i) {
        int r = new_batch_idx[i];
        for (int j = 0; j < lengths[r]; ++j) { rows[t] = r; poss[t] = j; ++t; }
    }
#pragma omp parallel num_threads(g_threads)
    {
        float* kacc = (float*)malloc(sizeof(float) * (size_t)d * 2);
        float* vacc = kacc + d;
#pragma omp for schedule(dynamic, 4)
        for (long long u = 0; u < total; ++u) {
            int r = rows[u], j = poss[u];
            const float* x = page_row(page_table, r, W, j, d, 0);
            memset(kacc, 0, sizeof(float) * (size_t)d * 2);
            for (int w = 0; w < d; ++w) {
                axpy_fma(x[w], wk + (size_t)w * d, kacc, d);
                axpy_fma(x[w], wv + (size_t)w * d, vacc, d);
            }
            memcpy(page_row(page_table, r, W, j, d, 1), kacc, sizeof(float) * (size_t)d);
            memcpy(page_row(page_table, r, W, j, d, 2), vacc, sizeof(float) * (size_t)d);
        }
        free(kacc);
    }
    free(rows);
    free(poss);
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/paged_attention.cu:126-180 (get_latest_k_q_v_paged_attention); host twin
 * tests/test_utils.cpp:363-403.  x = page[(L-1)].inp; k,v -> same page; q -> q_output.
 * Rows with L == 0 are skipped (q_output untouched).
 * ---------------------------------------------------------------------------------------- */
void orc_qkv_latest_paged(float** page_table, const int* lengths, const float* wk,
                          const float* wq, const float* wv, float* q_output, int n_batch,
                          int n_sequence, int emb_dim) {
    int W = n_sequence / ORC_PAGE_BLOCK;
    int d = emb_dim;
#pragma omp parallel num_threads(g_threads)
    {
        float* acc = (float*)malloc(sizeof(float) * (size_t)d * 3);
#pragma omp for schedule(dynamic)
        for (int r = 0; r < n_batch; ++r) {
            int L = lengths[r];
            if (L == 0) continue;
            int j = L - 1;
            const float* x = page_row(page_table, r, W, j, d, 0);
            memset(acc, 0, sizeof(float) * (size_t)d * 3);
            for (int w = 0; w < d; ++w) {
                axpy_fma(x[w], wk + (size_t)w * d, acc, d);
                axpy_fma(x[w], wv + (size_t)w * d, acc + d, d);
                axpy_fma(x[w], wq + (size_t)w * d, acc + 2 * d, d);
            }
            memcpy(page_row(page_table, r, W, j, d, 1), acc, sizeof(float) * (size_t)d);
            memcpy(page_row(page_table, r, W, j, d, 2), acc + d, sizeof(float) * (size_t)d);
            memcpy(q_output + (size_t)r * d, acc + 2 * d, sizeof(float) * (size_t)d);
        }
        free(acc);
    }
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/paged_attention.cu:208-263 (qkt_paged_attention); host twin
 * tests/test_utils.cpp:410-436.  qkt[r,j] = (sum_k q[r,k]*K[r,j,k]) / sqrtf(d), j < L_r.
 * Entries j >= L_r are left untouched (the softmax zero-fills them).
 * The device divides by sqrtf((float)d) in fp32 (paged_attention.cu:261); the host test divides
 * by a double sqrt -- the CUDA form is the one tokens must match, so it is used here.
 * ---------------------------------------------------------------------------------------- */
void orc_qkt_paged(const float* q, float** page_table, const int* lengths, float* qkt,
                   int n_batch, int n_sequence, int emb_dim) {
    int W = n_sequence / ORC_PAGE_BLOCK;
    int d = emb_dim;
    float denom = sqrtf((float)d);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int r = 0; r < n_batch; ++r) {
        int L = lengths[r];
        const float* qr = q + (size_t)r * d;
        for (int j = 0; j < L; ++j) {
            const float* k = page_row(page_table, r, W, j, d, 1);
            float s = 0.0f;
            for (int c = 0; c < d; ++c) s = __builtin_fmaf(qr[c], k[c], s);
            qkt[(size_t)r * n_sequence + j] = s / denom;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/self_attention_inference_optimized.cu:191-242 (softmax_in_place_with_lengths);
 * host twin tests/test_utils.cpp:443-469.  softmax over the first L_r entries, zeros to S.
 * Uses the device's final form expf(x - max) * (1.f / sum) (:226-236).  The device accumulates
 * the sum lane-wise with online rescaling and a warp reduce; this is a plain ascending sum, so
 * agreement with the device is to rounding (~1e-6 rel), which is what the tests allow.
 * ---------------------------------------------------------------------------------------- */
void orc_softmax_in_place_with_lengths(float* qkt, const int* lengths, int n_batch,
                                       int n_sequence) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int r = 0; r < n_batch; ++r) {
        float* row = qkt + (size_t)r * n_sequence;
        int L = lengths[r];
        float maxv = -FLT_MAX;
        for (int j = 0; j < L; ++j) maxv = fmaxf(maxv, row[j]);
        float sum = 0.0f;
        for (int j = 0; j < L; ++j) sum += expf(row[j] - maxv);
        float norm = 1.f / sum;
        for (int j = 0; j < n_sequence; ++j) row[j] = (j < L) ? expf(row[j] - maxv) * norm : 0.0f;
    }
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/paged_attention.cu:287-326 (softmax_v_paged_attention); host twin
 * tests/test_utils.cpp:476-500.  out[r,c] = sum_{j<L_r} p[r,j]*V[r,j,c], j-ascending FMA chain.
 * Rows with L == 0 get zeros (:289 `result = 0.0` is still stored at :323-325).
 * ---------------------------------------------------------------------------------------- */
void orc_softmax_v_paged(const float* softmax_result, float** page_table,
                         float* attention_result, const int* lengths, int n_batch,
                         int n_sequence, int emb_dim) {
    int W = n_sequence / ORC_PAGE_BLOCK;
    int d = emb_dim;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int r = 0; r < n_batch; ++r) {
        int L = lengths[r];
        float* out = attention_result + (size_t)r * d;
        memset(out, 0, sizeof(float) * (size_t)d);
        const float* p = softmax_result + (size_t)r * n_sequence;
        for (int j = 0; j < L; ++j) axpy_fma(p[j], page_row(page_table, r, W, j, d, 2), out, d);
    }
}

/* src/kernels/paged_attention.cu:358-377 (paged_attention) */
void orc_paged_attention(float** page_table, const int* lengths, const float* wk,
                         const float* wq, const float* wv, const int* new_batch_idx,
                         float* q_output, float* qkt_output, float* attention_result,
                         int n_new_items, int n_batch, int n_sequence, int emb_dim) {
    if (n_new_items > 0)
        orc_prefill_kv_paged(page_table, new_batch_idx, lengths, wk, wv, n_new_items, n_sequence,
                             emb_dim);
    orc_qkv_latest_paged(page_table, lengths, wk, wq, wv, q_output, n_batch, n_sequence, emb_dim);
    orc_qkt_paged(q_output, page_table, lengths, qkt_output, n_batch, n_sequence, emb_dim);
    orc_softmax_in_place_with_lengths(qkt_output, lengths, n_batch, n_sequence);
    orc_softmax_v_paged(qkt_output, page_table, attention_result, lengths, n_batch, n_sequence,
                        emb_dim);
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/gemm.cu:13-60 (gemm_transpose_kernel) as called at src/kernels/decoder.cu:222;
 * host twin tests/test_utils.cpp:575-578.  logits[r,v] = sum_k attn[r,k]*E[v,k], k-ascending.
 * ---------------------------------------------------------------------------------------- */
void orc_logits(const float* batch_result, const float* emb_table, float* emb_score,
                int n_batch, int n_vocab, int emb_dim) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int r = 0; r < n_batch; ++r) {
        const float* a = batch_result + (size_t)r * emb_dim;
        for (int v = 0; v < n_vocab; ++v) {
            const float* e = emb_table + (size_t)v * emb_dim;
            float s = 0.0f;
            for (int c = 0; c < emb_dim; ++c) s = __builtin_fmaf(a[c], e[c], s);
            emb_score[(size_t)r * n_vocab + v] = s;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/decoder.cu:146-172: 256 threads each scan indices = t (mod 256) keeping the first
 * strict maximum, then a tree reduction where the lower thread wins ties.  Simulated literally.
 * (The host test uses "lowest index wins", tests/test_utils.cpp:607-614; they differ only on
 * exact ties -- SURVEY App. A Q4.)
 * ---------------------------------------------------------------------------------------- */
int orc_argmax_device_rule(const float* score, int n_vocab) {
    enum { BD = 256 };
    float mv[BD];
    int mi[BD];
    for (int t = 0; t < BD; ++t) {
        float lm = -FLT_MAX;
        int li = -1;
        for (int i = t; i < n_vocab; i += BD) {
            if (score[i] > lm) { lm = score[i]; li = i; }
        }
        mv[t] = lm;
        mi[t] = li;
    }
    for (int gap = BD / 2; gap > 0; gap >>= 1) {
        for (int t = 0; t < gap; ++t) {
            if (mv[t + gap] > mv[t]) { mv[t] = mv[t + gap]; mi[t] = mi[t + gap]; }
        }
    }
    return mi[0];
}

/* ------------------------------------------------------------------------------------------
 * src/kernels/decoder.cu:128-205 (paged_attention_decoder_kernel_with_multi_decoder); host twin
 * tests/test_utils.cpp:593-630.  token -> decoder_result[r, i_decoder]; lengths = L+1, or 0 when
 * token == EOF or L+1 >= S; otherwise page[r][L].inp = E[token] + P[L].
 * ---------------------------------------------------------------------------------------- */
void orc_paged_decoder(const float* emb_score, int* decoder_result, int* lengths,
                       float** page_table, const float* pos_table, const float* emb_table,
                       int n_batch, int n_vocab, int n_sequence, int emb_dim,
                       int n_decoder_results, int i_decoder) {
    int W = n_sequence / ORC_PAGE_BLOCK;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int r = 0; r < n_batch; ++r) {
        int L = lengths[r];
        int* out = decoder_result + (size_t)r * n_decoder_results + i_decoder;
        if (L == 0) { *out = ORC_EMPTY_TOKEN; continue; }
        int tok = orc_argmax_device_rule(emb_score + (size_t)r * n_vocab, n_vocab);
        *out = tok;
        lengths[r] = L + 1;
        if (tok == ORC_EOF_TOKEN || L + 1 >= n_sequence) { lengths[r] = 0; continue; }
        const float* e = emb_table + (size_t)tok * emb_dim;
        const float* p = pos_table + (size_t)L * emb_dim;
        float* x = page_row(page_table, r, W, L, emb_dim, 0);
        for (int c = 0; c < emb_dim; ++c) x[c] = e[c] + p[c];
    }
}

/* src/inference_model.cpp:52-82 (PagedAttentionInferenceModel::forward) */
void orc_paged_forward(const int* inp, int* lengths, const int* new_item_indices,
                       int* decoder_result, int n_new_items, const float* emb_table,
                       const float* pos_table, float** page_table, const float* wk,
                       const float* wq, const float* wv, float* attention_result,
                       float* emb_score, int n_batch, int n_sequence, int emb_dim, int n_vocab,
                       int n_forward_rounds) {
    float* q_output = (float*)malloc(sizeof(float) * (size_t)n_batch * emb_dim);
    float* qkt_output = (float*)calloc((size_t)n_batch * n_sequence, sizeof(float));
    float* attn = attention_result ? attention_result
                                   : (float*)malloc(sizeof(float) * (size_t)n_batch * emb_dim);
    float* score =
        emb_score ? emb_score : (float*)malloc(sizeof(float) * (size_t)n_batch * n_vocab);
    for (int i = 0; i < n_forward_rounds; ++i) {
        if (i > 0) n_new_items = 0;
        if (n_new_items > 0)
            orc_paged_encoder(emb_table, pos_table, inp, page_table, lengths, new_item_indices,
                              n_batch, n_sequence, emb_dim, n_new_items);
        orc_paged_attention(page_table, lengths, wk, wq, wv, new_item_indices, q_output,
                            qkt_output, attn, n_new_items, n_batch, n_sequence, emb_dim);
        orc_logits(attn, emb_table, score, n_batch, n_vocab, emb_dim);
        orc_paged_decoder(score, decoder_result, lengths, page_table, pos_table, emb_table,
                          n_batch, n_vocab, n_sequence, emb_dim, n_forward_rounds, i);
    }
    free(q_output);
    free(qkt_output);
    if (!attention_result) free(attn);
    if (!emb_score) free(score);
}

/* ==========================================================================================
 * Dense (non-paged) path: config C1.  kt_cache is TRANSPOSED [B, d_out, S]; v_cache [B, S, d_out]
 * ======================================================================================== */

/* src/kernels/encoder.cu:56-92 (inference_optimized_encoder); tests/test_utils.cpp:559-573 */
void orc_dense_encoder(const float* emb_table, const float* pos_table, const int* inp,
                       float* inp_embedding, const int* lengths, const int* new_item_indices,
                       int n_batch, int n_sequence, int emb_dim, int n_new_items) {
    (void)n_batch;
    for (int i = 0; i < n_new_items; ++i) {
        int r = new_item_indices[i];
        for (int j = 0; j < lengths[r]; ++j) {
            int tok = inp[(size_t)r * n_sequence + j];
            float* x = inp_embedding + ((size_t)r * n_sequence + j) * emb_dim;
            for (int c = 0; c < emb_dim; ++c)
                x[c] = emb_table[(size_t)tok * emb_dim + c] + pos_table[(size_t)j * emb_dim + c];
        }
    }
}

/* src/kernels/self_attention_inference_optimized.cu:282-301 (inference_self_attention) =
 * fill_new_kt_v_cache (:27-85) -> get_latest_kt_q_v (:100-143) -> qkt (:150-184) ->
 * softmax_in_place_with_lengths (:191-242) -> softmax_v (:249-279);
 * host twin tests/test_utils.cpp:502-519. */
void orc_self_attention(const float* inp_embedding, const int* lengths, const float* wk,
                        const float* wq, const float* wv, const int* new_batch_idx,
                        float* kt_cache, float* v_cache, float* q_output, float* qkt_output,
                        float* attention_result, int n_new_items, int n_batch, int n_sequence,
                        int input_dim, int output_dim) {
    int S = n_sequence, di = input_dim, dn = output_dim;
    float* acc = (float*)malloc(sizeof(float) * (size_t)dn * 3);
    /* prefill for new rows (:27-85) */
    for (int i = 0; i < n_new_items; ++i) {
        int r = new_batch_idx[i];
        for (int j = 0; j < lengths[r]; ++j) {
            const float* x = inp_embedding + ((size_t)r * S + j) * di;
            memset(acc, 0, sizeof(float) * (size_t)dn * 2);
            for (int w = 0; w < di; ++w) {
                axpy_fma(x[w], wk + (size_t)w * dn, acc, dn);
                axpy_fma(x[w], wv + (size_t)w * dn, acc + dn, dn);
            }
            for (int c = 0; c < dn; ++c) kt_cache[((size_t)r * dn + c) * S + j] = acc[c];
            memcpy(v_cache + ((size_t)r * S + j) * dn, acc + dn, sizeof(float) * (size_t)dn);
        }
    }
    /* latest token (:100-143) */
    for (int r = 0; r < n_batch; ++r) {
        int L = lengths[r];
        if (L == 0) continue;
        int j = L - 1;
        const float* x = inp_embedding + ((size_t)r * S + j) * di;
        memset(acc, 0, sizeof(float) * (size_t)dn * 3);
        for (int w = 0; w < di; ++w) {
            axpy_fma(x[w], wk + (size_t)w * dn, acc, dn);
            axpy_fma(x[w], wv + (size_t)w * dn, acc + dn, dn);
            axpy_fma(x[w], wq + (size_t)w * dn, acc + 2 * dn, dn);
        }
        for (int c = 0; c < dn; ++c) kt_cache[((size_t)r * dn + c) * S + j] = acc[c];
        memcpy(v_cache + ((size_t)r * S + j) * dn, acc + dn, sizeof(float) * (size_t)dn);
        memcpy(q_output + (size_t)r * dn, acc + 2 * dn, sizeof(float) * (size_t)dn);
    }
    free(acc);
    /* qkt (:150-184) */
    float denom = sqrtf((float)dn);
    for (int r = 0; r < n_batch; ++r) {
        int L = lengths[r];
        for (int j = 0; j < L; ++j) {
            float s = 0.0f;
            for (int c = 0; c < dn; ++c)
                s = __builtin_fmaf(q_output[(size_t)r * dn + c],
                                   kt_cache[((size_t)r * dn + c) * S + j], s);
            qkt_output[(size_t)r * S + j] = s / denom;
        }
    }
    orc_softmax_in_place_with_lengths(qkt_output, lengths, n_batch, S);
    /* softmax_v (:249-279) */
    for (int r = 0; r < n_batch; ++r) {
        float* out = attention_result + (size_t)r * dn;
        memset(out, 0, sizeof(float) * (size_t)dn);
        for (int j = 0; j < lengths[r]; ++j)
            axpy_fma(qkt_output[(size_t)r * S + j], v_cache + ((size_t)r * S + j) * dn, out, dn);
    }
}

/* src/kernels/decoder.cu:25-91 (decoder_kernel); tests/test_utils.cpp:593-630 */
void orc_dense_decoder(const float* emb_score, int* decoder_result, int* lengths,
                       float* inp_embedding, const float* pos_table, const float* emb_table,
                       int n_batch, int n_vocab, int n_sequence, int emb_dim) {
    for (int r = 0; r < n_batch; ++r) {
        int L = lengths[r];
        if (L == 0) { decoder_result[r] = ORC_EMPTY_TOKEN; continue; }
        int tok = orc_argmax_device_rule(emb_score + (size_t)r * n_vocab, n_vocab);
        decoder_result[r] = tok;
        lengths[r] = L + 1;
        if (L + 1 >= n_sequence || tok == ORC_EOF_TOKEN) { lengths[r] = 0; continue; }
        float* x = inp_embedding + ((size_t)r * n_sequence + L) * emb_dim;
        for (int c = 0; c < emb_dim; ++c)
            x[c] = emb_table[(size_t)tok * emb_dim + c] + pos_table[(size_t)L * emb_dim + c];
    }
}

/* src/inference_model.cpp:14-39 (InferenceModel::forward) */
void orc_dense_forward(const int* inp, int* lengths, const int* new_item_indices,
                       int* decoder_result, int n_new_items, const float* emb_table,
                       const float* pos_table, const float* wk, const float* wq,
                       const float* wv, float* inp_embedding, float* kt_cache, float* v_cache,
                       int n_batch, int n_sequence, int emb_dim, int n_vocab) {
    float* q_output = (float*)malloc(sizeof(float) * (size_t)n_batch * emb_dim);
    float* qkt_output = (float*)calloc((size_t)n_batch * n_sequence, sizeof(float));
    float* attn = (float*)malloc(sizeof(float) * (size_t)n_batch * emb_dim);
    float* score = (float*)malloc(sizeof(float) * (size_t)n_batch * n_vocab);
    orc_dense_encoder(emb_table, pos_table, inp, inp_embedding, lengths, new_item_indices, n_batch,
                      n_sequence, emb_dim, n_new_items);
    orc_self_attention(inp_embedding, lengths, wk, wq, wv, new_item_indices, kt_cache, v_cache,
                       q_output, qkt_output, attn, n_new_items, n_batch, n_sequence, emb_dim,
                       emb_dim);
    orc_logits(attn, emb_table, score, n_batch, n_vocab, emb_dim);
    orc_dense_decoder(score, decoder_result, lengths, inp_embedding, pos_table, emb_table, n_batch,
                      n_vocab, n_sequence, emb_dim);
    free(q_output);
    free(qkt_output);
    free(attn);
    free(score);
}

/* ==========================================================================================
 * Engines.  Request bookkeeping restates src/item_storage.cpp and src/paged_item_storage.cpp with
 * arrays instead of std::list / unordered_map; ORDER semantics are kept exactly (queue order,
 * used-block list order, FIFO free list, tail pre-emption).
 * ======================================================================================== */
typedef struct {
    int n_req, S;
    int* tok;  /* [n_req][S] prompt + generated */
    int* cnt;  /* [n_req] */
    int* plen; /* [n_req] prompt length at submission */
    int max_new; /* opt-in policy (not in the reference): > 0 = finish after this many generated tokens */
    /* Storage new_items_ (src/item_storage.cpp:12-95): deque of request ids */
    int* q;
    int qcap, qhead, qcount;
    /* finished (in finish order) */
    int* fin;
    int nfin;
} req_store;

static void rs_init(req_store* rs, int n_req, int S, const int* off, const int* toks) {
    rs->n_req = n_req;
    rs->S = S;
    rs->tok = (int*)calloc((size_t)n_req * S, sizeof(int));
    rs->cnt = (int*)calloc((size_t)n_req, sizeof(int));
    rs->plen = (int*)calloc((size_t)n_req, sizeof(int));
    rs->max_new = 0;
    rs->qcap = n_req + 1;
    rs->q = (int*)malloc(sizeof(int) * (size_t)rs->qcap);
    rs->qhead = 0;
    rs->qcount = 0;
    rs->fin = (int*)malloc(sizeof(int) * (size_t)(n_req > 0 ? n_req : 1));
    rs->nfin = 0;
    for (int i = 0; i < n_req; ++i) {
        int n = off[i + 1] - off[i];
        memcpy(rs->tok + (size_t)i * S, toks + off[i], sizeof(int) * (size_t)n);
        rs->cnt[i] = n;
        rs->plen[i] = n;
        rs->q[(rs->qhead + rs->qcount++) % rs->qcap] = i; /* add_new_item: push_back */
    }
}
static void rs_free(req_store* rs) {
    free(rs->tok); free(rs->cnt); free(rs->plen); free(rs->q); free(rs->fin);
}
static int rs_pop_front(req_store* rs) {
    int id = rs->q[rs->qhead];
    rs->qhead = (rs->qhead + 1) % rs->qcap;
    rs->qcount--;
    return id;
}
static void rs_push_front(req_store* rs, int id) { /* add_new_item_to_head (item_storage.cpp:194) */
    rs->qhead = (rs->qhead - 1 + rs->qcap) % rs->qcap;
    rs->q[rs->qhead] = id;
    rs->qcount++;
}
static int rs_head_len(const req_store* rs) { return rs->cnt[rs->q[rs->qhead]]; }

static void emit_finished(const req_store* rs, int* finished_ids, int* finished_offsets,
                          int* finished_tokens) {
    int o = 0;
    for (int i = 0; i < rs->nfin; ++i) {
        int id = rs->fin[i];
        finished_ids[i] = id;
        finished_offsets[i] = o;
        memcpy(finished_tokens + o, rs->tok + (size_t)id * rs->S, sizeof(int) * (size_t)rs->cnt[id]);
        o += rs->cnt[id];
    }
    finished_offsets[rs->nfin] = o;
}

/* src/item_storage.cpp:97-139 (process_decoder_result).  row_req[r] = request in row r or -1.
 * Returns number of finished_indices written; *phantom is set if a token arrives for a row that
 * is not processing (the reference would default-construct an entry there, :117). */
static int process_decoder_result(const int* dec, int B, int R, int S, req_store* rs, int* row_req,
                                  int* finished_indices, long long* gen, int* phantom, const int* pf_pos) {
    int nf = 0;
    for (int i = 0; i < B; ++i) {
        int empty = 0, finished = 0;
        for (int j = 0; j < R; ++j) {
            int t = dec[(size_t)i * R + j];
            if (t == ORC_EMPTY_TOKEN) {
                /* chunked prefill: a row whose prompt is still being prefilled has no token yet and is not free */
                if (pf_pos && row_req[i] >= 0 && pf_pos[i] >= 0) break;
                empty = 1;
            } else {
                int id = row_req[i];
                if (id < 0) { *phantom = 1; break; }
                if (rs->cnt[id] < S) rs->tok[(size_t)id * S + rs->cnt[id]] = t;
                rs->cnt[id]++;
                (*gen)++;
                if (rs->cnt[id] >= S || t == ORC_EOF_TOKEN) finished = 1;
                if (rs->max_new > 0 && rs->cnt[id] - rs->plen[id] >= rs->max_new) finished = 1; /* opt-in */
            }
            if (finished || empty) break;
        }
        if (finished || empty) finished_indices[nf++] = i;
        if (finished) { /* move_to_finished (:72-76) */
            int id = row_req[i];
            if (rs->cnt[id] > S) rs->cnt[id] = S;
            rs->fin[rs->nfin++] = id;
            row_req[i] = -1;
        }
    }
    return nf;
}

/* ---- paged scheduler state (src/paged_item_storage.cpp) -------------------------------- */
typedef struct {
    int B, S, W, d, R, n_blocks;
    float* slab;          /* MemoryBlockManager::block_memory_ (:125-133) */
    int* freeq;           /* FIFO of page ids (:136-153) */
    int fhead, fcount;
    int* used_rows;       /* used_blocks_ list order (:155-194) */
    int n_used;
    int* row_pages;       /* [B][W] page ids in LIST order (front = index 0) */
    int* row_npages;      /* [B] */
    float** pt_host;      /* [B][W] */
    float** pt_dev;
    int needs_sync;
    int *inp_host, *inp_dev, *len_host, *len_dev, *idx_host, *idx_dev;
    int* row_req;
    /* opt-in chunked prefill of the product's engine (not in the reference; 0 = off) */
    int chunk;            /* prompt positions prefilled per step (multiple of 16) */
    int* pf_pos;          /* [B] positions of the row's prompt already scheduled, -1 = not prefilling */
    int* pf_len;          /* [B] prompt length of a prefilling row */
} paged_state;

static float* page_ptr(const paged_state* ps, int id) {
    return ps->slab + (size_t)id * ORC_PAGE_BLOCK * 3 * ps->d;
}
static void free_push_back(paged_state* ps, int id) {
    ps->freeq[(ps->fhead + ps->fcount++) % ps->n_blocks] = id;
}
static int free_pop_front(paged_state* ps) {
    int id = ps->freeq[ps->fhead];
    ps->fhead = (ps->fhead + 1) % ps->n_blocks;
    ps->fcount--;
    return id;
}
static void return_row_pages(paged_state* ps, int row) { /* return_free_blocks: splice to end */
    for (int k = 0; k < ps->row_npages[row]; ++k) free_push_back(ps, ps->row_pages[(size_t)row * ps->W + k]);
    ps->row_npages[row] = 0;
}
static void used_erase_at(paged_state* ps, int pos) {
    for (int k = pos; k + 1 < ps->n_used; ++k) ps->used_rows[k] = ps->used_rows[k + 1];
    ps->n_used--;
}
static void move_to_new(paged_state* ps, req_store* rs, int row) { /* item_storage.cpp:75-79 */
    rs_push_front(rs, ps->row_req[row]);
    ps->row_req[row] = -1;
    if (ps->pf_pos) ps->pf_pos[row] = -1; /* a pre-empted prompt starts over when it is re-admitted */
}

/* src/paged_item_storage.cpp:14-60 (allocate_or_free_memory_blocks_if_needed) */
static void allocate_or_free(paged_state* ps, req_store* rs, const int* finished_indices, int nf,
                             long long* preemptions) {
    char* fin = (char*)calloc((size_t)ps->B, 1);
    for (int i = 0; i < nf; ++i) fin[finished_indices[i]] = 1;
    /* 1. free finished rows, in used-list order (:23-32) */
    for (int p = 0; p < ps->n_used;) {
        int row = ps->used_rows[p];
        if (fin[row]) { return_row_pages(ps, row); used_erase_at(ps, p); } else { ++p; }
    }
    free(fin);
    /* 2. grow / pre-empt (:36-59); note the iterator is NOT advanced after a successful
     *    allocation or a tail pre-emption, so the same row is re-examined. */
    for (int p = 0; p < ps->n_used;) {
        int row = ps->used_rows[p];
        int id = ps->row_req[row];
        if (rs->cnt[id] + ps->R > ps->row_npages[row] * ORC_PAGE_BLOCK) {
            if (ps->fcount > 0) {
                /* allocate_memory_block (:196-203): push_front, table index = size-1 */
                int pg = free_pop_front(ps);
                int n = ps->row_npages[row];
                int* lst = ps->row_pages + (size_t)row * ps->W;
                if (n < ps->W) {
                    memmove(lst + 1, lst, sizeof(int) * (size_t)n);
                    lst[0] = pg;
                    ps->row_npages[row] = n + 1;
                    ps->pt_host[(size_t)row * ps->W + n] = page_ptr(ps, pg);
                    ps->needs_sync = 1;
                } else {
                    /* table row is full (only reachable when tokens+R > S); keep the page
                     * accounted to the row without touching the table */
                    free_push_back(ps, pg);
                    ++p;
                }
            } else if (p + 1 == ps->n_used) {
                move_to_new(ps, rs, row);
                return_row_pages(ps, row);
                used_erase_at(ps, p);
                (*preemptions)++;
            } else {
                int tail = ps->used_rows[ps->n_used - 1];
                ps->n_used--;
                move_to_new(ps, rs, tail);
                return_row_pages(ps, tail);
                (*preemptions)++;
            }
        } else {
            ++p;
        }
    }
}

/* src/paged_item_storage.cpp:62-122 (paged insert_new_items).  Returns n_new; the row indices
 * are in ps->idx_dev[0..n_new). */
static int paged_insert_new_items(paged_state* ps, req_store* rs, int fix_stale_lengths,
                                  int max_prefill) {
    int B = ps->B, S = ps->S, W = ps->W, R = ps->R;
    int admitted_positions = 0; /* opt-in admission throttle (not in the reference; 0 = off) */
    int n_admitted_chunk = 0;   /* chunk mode: admissions of this call (they are not "new items" yet) */
    char* occ = (char*)calloc((size_t)B, 1);
    for (int p = 0; p < ps->n_used; ++p) occ[ps->used_rows[p]] = 1;
    if (fix_stale_lengths) memcpy(ps->len_host, ps->len_dev, sizeof(int) * (size_t)B);
    int need_copy = 0, n_new = 0;
    for (int i = 0; i < B; ++i) {
        if (occ[i]) continue;
        if (ps->fcount >= ORC_INIT_BLOCKS && rs->qcount > 0 &&
            ps->fcount >= ceil_div_i(rs_head_len(rs) + R, ORC_PAGE_BLOCK) &&
            (max_prefill <= 0 || n_new + n_admitted_chunk == 0 ||
             admitted_positions + rs_head_len(rs) <= max_prefill)) {
            int id = rs_pop_front(rs);
            int len = rs->cnt[id];
            admitted_positions += len;
            memcpy(ps->inp_host + (size_t)i * S, rs->tok + (size_t)id * S, sizeof(int) * (size_t)len);
            if (ps->chunk > 0) { /* inactive until its last chunk is scheduled (plan_prefill_chunks) */
                ps->len_host[i] = 0;
                ps->pf_pos[i] = 0;
                ps->pf_len[i] = len;
                n_admitted_chunk++;
            } else {
                ps->len_host[i] = len;
                ps->idx_host[n_new++] = i;
            }
            int nb = ceil_div_i(len + R, ORC_PAGE_BLOCK);
            if (nb < ORC_INIT_BLOCKS) nb = ORC_INIT_BLOCKS;
            ps->row_req[i] = id;
            /* pop_free_blocks(nb) + add_batch_block_pair (:176-189) */
            for (int k = 0; k < nb; ++k) {
                int pg = free_pop_front(ps);
                if (k < W) {
                    ps->row_pages[(size_t)i * W + k] = pg;
                    ps->pt_host[(size_t)i * W + k] = page_ptr(ps, pg);
                } else {
                    free_push_back(ps, pg); /* cannot be represented in the table */
                }
            }
            ps->row_npages[i] = nb < W ? nb : W;
            ps->used_rows[ps->n_used++] = i;
            ps->needs_sync = 1;
            need_copy = 1;
        } else {
            ps->len_host[i] = 0;
            need_copy = 1;
        }
    }
    free(occ);
    if (need_copy) { /* :113-118 -- copies the WHOLE (possibly stale) lengths array: quirk Q1 */
        memcpy(ps->inp_dev, ps->inp_host, sizeof(int) * (size_t)B * S);
        memcpy(ps->len_dev, ps->len_host, sizeof(int) * (size_t)B);
        memcpy(ps->idx_dev, ps->idx_host, sizeof(int) * (size_t)B);
    }
    if (ps->needs_sync) { /* maybe_flush_changes (:167-172) */
        memcpy(ps->pt_dev, ps->pt_host, sizeof(float*) * (size_t)B * W);
        ps->needs_sync = 0;
    }
    return n_new;
}

/* Chunked prefill (opt-in policy of the product's engine, csrc/engine.cu; the reference prefills a whole prompt in
 * the step that admits it): rows in admission order take 16-position granules of their remaining prompt until
 * the step's budget is spent.  A row whose last granule is scheduled becomes active in this step -- it is
 * appended to the new-item list, so the forward below prefills it (K and V do not depend on how the positions
 * were cut into chunks) and emits its first token.  Returns the number of rows that became active. */
static int plan_prefill_chunks(paged_state* ps, int n_new) {
    int budget = ps->chunk / ORC_PAGE_BLOCK;
    for (int p = 0; p < ps->n_used && budget > 0; ++p) {
        int row = ps->used_rows[p];
        if (ps->pf_pos[row] < 0) continue;
        int g_rem = ceil_div_i(ps->pf_len[row] - ps->pf_pos[row], ORC_PAGE_BLOCK);
        int take = g_rem < budget ? g_rem : budget;
        budget -= take;
        if (take == g_rem) {
            ps->pf_pos[row] = -1;
            ps->len_host[row] = ps->pf_len[row];
            ps->len_dev[row] = ps->pf_len[row];
            ps->idx_dev[n_new] = row;
            ps->idx_host[n_new] = row;
            n_new++;
        } else {
            ps->pf_pos[row] += take * ORC_PAGE_BLOCK;
        }
    }
    return n_new;
}

/* src/inferencer.cpp:43-85 (start_paged_attention_inference_engine) */
int orc_paged_engine_run(const orc_engine_cfg* cfg, const float* emb_table,
                         const float* pos_table, const float* wk, const float* wq,
                         const float* wv, int n_req, const int* prompt_offsets,
                         const int* prompt_tokens, int* finished_ids, int* finished_offsets,
                         int* finished_tokens, orc_engine_stats* stats) {
    int B = cfg->n_batch, S = cfg->n_sequence, d = cfg->emb_dim, V = cfg->n_vocab;
    int R = cfg->n_forward_rounds;
    if (S % ORC_PAGE_BLOCK != 0 || R < 1 || R > ORC_PAGE_BLOCK) return -1;
    int W = S / ORC_PAGE_BLOCK;
    paged_state ps;
    memset(&ps, 0, sizeof(ps));
    ps.B = B; ps.S = S; ps.W = W; ps.d = d; ps.R = R; ps.n_blocks = cfg->n_blocks;
    ps.slab = (float*)calloc((size_t)cfg->n_blocks * ORC_PAGE_BLOCK * 3 * d, sizeof(float));
    ps.freeq = (int*)malloc(sizeof(int) * (size_t)cfg->n_blocks);
    for (int i = 0; i < cfg->n_blocks; ++i) free_push_back(&ps, i);
    ps.used_rows = (int*)malloc(sizeof(int) * (size_t)B);
    ps.row_pages = (int*)malloc(sizeof(int) * (size_t)B * W);
    ps.row_npages = (int*)calloc((size_t)B, sizeof(int));
    ps.pt_host = (float**)calloc((size_t)B * W, sizeof(float*));
    ps.pt_dev = (float**)calloc((size_t)B * W, sizeof(float*));
    ps.inp_host = (int*)calloc((size_t)B * S, sizeof(int));
    ps.inp_dev = (int*)calloc((size_t)B * S, sizeof(int));
    ps.len_host = (int*)calloc((size_t)B, sizeof(int));
    ps.len_dev = (int*)calloc((size_t)B, sizeof(int));
    ps.idx_host = (int*)calloc((size_t)B, sizeof(int));
    ps.idx_dev = (int*)calloc((size_t)B, sizeof(int));
    ps.row_req = (int*)malloc(sizeof(int) * (size_t)B);
    for (int i = 0; i < B; ++i) ps.row_req[i] = -1;
    ps.chunk = cfg->prefill_chunk_positions / ORC_PAGE_BLOCK * ORC_PAGE_BLOCK;
    if (ps.chunk > 0 && !cfg->fix_stale_lengths) return -1; /* needs corrected lengths */
    if (ps.chunk > 0) {
        ps.pf_pos = (int*)malloc(sizeof(int) * (size_t)B);
        ps.pf_len = (int*)calloc((size_t)B, sizeof(int));
        for (int i = 0; i < B; ++i) ps.pf_pos[i] = -1;
    }
    req_store rs;
    rs_init(&rs, n_req, S, prompt_offsets, prompt_tokens);
    int* dec = (int*)malloc(sizeof(int) * (size_t)B * R);
    int* finished_indices = (int*)malloc(sizeof(int) * (size_t)B);
    float* attn = (float*)malloc(sizeof(float) * (size_t)B * d);
    float* score = (float*)malloc(sizeof(float) * (size_t)B * V);
    long long steps = 0, gen = 0, pre = 0;
    int rc = 0, phantom = 0;

    rs.max_new = cfg->max_new_tokens;
    int n_new = paged_insert_new_items(&ps, &rs, cfg->fix_stale_lengths, cfg->max_prefill_positions);
    if (ps.chunk > 0) n_new = plan_prefill_chunks(&ps, n_new);
    for (;;) {
        int processing = 0;
        for (int i = 0; i < B; ++i) processing += (ps.row_req[i] >= 0);
        if (processing + rs.qcount == 0) break; /* is_done (item_storage.cpp:186-188) */
        /* nothing resident, so the whole pool is free, and the queue head still was not admitted: a
         * pre-empted request that outgrew the pool.  The reference spins here for ever
         * (paged_item_storage.cpp:84-113 returns 0 new items again and again); oracle and engine stop. */
        if (processing == 0) { rc = -5; break; }
        if (cfg->max_steps > 0 && steps >= cfg->max_steps) { rc = -4; break; }
        orc_paged_forward(ps.inp_dev, ps.len_dev, ps.idx_dev, dec, n_new, emb_table, pos_table,
                          ps.pt_dev, wk, wq, wv, attn, score, B, S, d, V, R);
        int nf = process_decoder_result(dec, B, R, S, &rs, ps.row_req, finished_indices, &gen,
                                        &phantom, ps.pf_pos);
        if (phantom) { rc = -3; break; }
        allocate_or_free(&ps, &rs, finished_indices, nf, &pre);
        n_new = paged_insert_new_items(&ps, &rs, cfg->fix_stale_lengths, cfg->max_prefill_positions);
        if (ps.chunk > 0) n_new = plan_prefill_chunks(&ps, n_new);
        ++steps;
    }
    emit_finished(&rs, finished_ids, finished_offsets, finished_tokens);
    if (stats) {
        stats->steps = steps; stats->generated_tokens = gen; stats->preemptions = pre;
        stats->n_finished = rs.nfin;
    }
    free(dec); free(finished_indices); free(attn); free(score);
    rs_free(&rs);
    free(ps.slab); free(ps.freeq); free(ps.used_rows); free(ps.row_pages); free(ps.row_npages);
    free(ps.pt_host); free(ps.pt_dev); free(ps.inp_host); free(ps.inp_dev); free(ps.len_host);
    free(ps.len_dev); free(ps.idx_host); free(ps.idx_dev); free(ps.row_req);
    free(ps.pf_pos); free(ps.pf_len);
    return rc;
}

/* src/inferencer.cpp:11-41 (start_inference_engine) with the non-paged insert_new_items
 * (src/item_storage.cpp:141-180), which DOES refresh host lengths from the device (:153-154). */
int orc_dense_engine_run(const orc_engine_cfg* cfg, const float* emb_table,
                         const float* pos_table, const float* wk, const float* wq,
                         const float* wv, int n_req, const int* prompt_offsets,
                         const int* prompt_tokens, int* finished_ids, int* finished_offsets,
                         int* finished_tokens, orc_engine_stats* stats) {
    int B = cfg->n_batch, S = cfg->n_sequence, d = cfg->emb_dim, V = cfg->n_vocab;
    req_store rs;
    rs_init(&rs, n_req, S, prompt_offsets, prompt_tokens);
    rs.max_new = cfg->max_new_tokens;
    int* inp = (int*)calloc((size_t)B * S, sizeof(int));
    int* len = (int*)calloc((size_t)B, sizeof(int));
    int* idx = (int*)calloc((size_t)B, sizeof(int));
    int* dec = (int*)malloc(sizeof(int) * (size_t)B);
    int* row_req = (int*)malloc(sizeof(int) * (size_t)B);
    int* finished_indices = (int*)malloc(sizeof(int) * (size_t)B);
    float* emb = (float*)calloc((size_t)B * S * d, sizeof(float));
    float* kt = (float*)calloc((size_t)B * S * d, sizeof(float));
    float* vc = (float*)calloc((size_t)B * S * d, sizeof(float));
    for (int i = 0; i < B; ++i) { row_req[i] = -1; finished_indices[i] = i; }
    int nf = B, rc = 0, phantom = 0;
    long long steps = 0, gen = 0;
    int n_new = 0;
    for (;;) {
        /* insert_new_items (item_storage.cpp:141-180) */
        n_new = 0;
        if (nf > 0) {
            int npop = nf < rs.qcount ? nf : rs.qcount;
            for (int i = 0; i < nf; ++i) {
                int row = finished_indices[i];
                idx[i] = row;
                if (i >= npop) {
                    len[row] = 0;
                } else {
                    int id = rs_pop_front(&rs);
                    len[row] = rs.cnt[id];
                    memcpy(inp + (size_t)row * S, rs.tok + (size_t)id * S,
                           sizeof(int) * (size_t)rs.cnt[id]);
                    row_req[row] = id;
                }
            }
            n_new = npop;
        }
        int processing = 0;
        for (int i = 0; i < B; ++i) processing += (row_req[i] >= 0);
        if (processing + rs.qcount == 0) break;
        if (cfg->max_steps > 0 && steps >= cfg->max_steps) { rc = -4; break; }
        orc_dense_forward(inp, len, idx, dec, n_new, emb_table, pos_table, wk, wq, wv, emb, kt, vc,
                          B, S, d, V);
        nf = process_decoder_result(dec, B, 1, S, &rs, row_req, finished_indices, &gen, &phantom, NULL);
        if (phantom) { rc = -3; break; }
        ++steps;
    }
    emit_finished(&rs, finished_ids, finished_offsets, finished_tokens);
    if (stats) {
        stats->steps = steps; stats->generated_tokens = gen; stats->preemptions = 0;
        stats->n_finished = rs.nfin;
    }
    free(inp); free(len); free(idx); free(dec); free(row_req); free(finished_indices);
    free(emb); free(kt); free(vc);
    rs_free(&rs);
    return rc;
}
