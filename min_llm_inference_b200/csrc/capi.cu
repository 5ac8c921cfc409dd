// capi.cu -- the C ABI (include/mli_b200.h): context, stage entry points, forward compositions.
// Every entry point only enqueues work on the context's stream; nothing here synchronises the
// device except workspace growth and mli_ctx_synchronize.
#include "common.cuh"
#include "kernels.h"

#include <atomic>
#include <cstdio>
#include <new>

namespace mli {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "[CUDA ERROR] at file %s:%d: %s", file, line, cudaGetErrorString(e));
    g_last_error = buf;
    return MLI_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ws_get(mli_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes < 256) bytes = 256;
    if (ctx->ws_bytes[slot] < bytes) {
        if (ctx->ws_frozen) {
            set_error("workspace would have to grow while a captured graph holds it");
            return MLI_ERR_STATE;
        }
        // grow: the old buffer may still be in use by enqueued kernels
        MLI_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->ws[slot]) MLI_CUDA(cudaFree(ctx->ws[slot]));
        ctx->ws[slot] = nullptr;
        ctx->ws_bytes[slot] = 0;
        size_t want = bytes + bytes / 4;
        MLI_CUDA(cudaMalloc(&ctx->ws[slot], want));
        ctx->ws_bytes[slot] = want;
    }
    *out = ctx->ws[slot];
    return 0;
}

int ws_get_zeroed(mli_ctx* ctx, int slot, size_t bytes, void** out) {
    const size_t had = ctx->ws_bytes[slot];
    int rc = ws_get(ctx, slot, bytes, out);
    if (rc) return rc;
    if (ctx->ws_bytes[slot] != had)
        MLI_CUDA(cudaMemsetAsync(ctx->ws[slot], 0, ctx->ws_bytes[slot], ctx->stream));
    return 0;
}

int ensure_dyn_smem_impl(mli_ctx* ctx, const void* func, size_t bytes) {
    if (bytes <= 48 * 1024) return 0;
    for (auto& e : ctx->smem_attr)
        if (e.first == func) {
            if (e.second >= bytes) return 0;
            MLI_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            e.second = bytes;
            return 0;
        }
    MLI_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->smem_attr.push_back({func, bytes});
    return 0;
}

static bool use_tc(mli_ctx* ctx) { return ctx->gemm_mode == 0 && ctx->tc_available; }

// tile-list workspace: [int n_tiles | pad to 16 B | TileDesc tiles[max_tiles]]
static int tiles_ws(mli_ctx* ctx, int max_tiles, int** n_tiles, TileDesc** tiles) {
    void* p = nullptr;
    int rc = ws_get(ctx, WS_TILES, 16 + sizeof(TileDesc) * (size_t)(max_tiles > 0 ? max_tiles : 1), &p);
    if (rc) return rc;
    *n_tiles = reinterpret_cast<int*>(p);
    *tiles = reinterpret_cast<TileDesc*>(reinterpret_cast<char*>(p) + 16);
    return 0;
}

static int check_paged_dims(int n_batch, int n_sequence, int emb_dim) {
    MLI_REQUIRE(n_batch > 0, "n_batch must be positive");
    MLI_REQUIRE(n_sequence > 0 && n_sequence % kPage == 0, "n_sequence must be a multiple of 16");
    MLI_REQUIRE(emb_dim > 0 && emb_dim % 4 == 0, "emb_dim must be a multiple of 4");
    return 0;
}

static int prefill_paged(mli_ctx* ctx, float** page_table, const TileDesc* tiles, const int* n_tiles,
                         int max_tiles, const int* lengths, const float* wk, const float* wv, int S,
                         int d) {
    if (use_tc(ctx))
        return launch_prefill_kv_paged_tc(ctx, page_table, tiles, n_tiles, max_tiles, lengths, wk, wv,
                                          S, d);
    return launch_prefill_kv_paged_simt(ctx, page_table, tiles, n_tiles, max_tiles, lengths, wk, wv, S,
                                        d);
}

static int latest_paged(mli_ctx* ctx, float** page_table, const int* lengths, const float* wk,
                        const float* wq, const float* wv, float* q_output, int B, int S, int d) {
    if (use_tc(ctx))
        return launch_qkv_latest_paged_tc(ctx, page_table, lengths, wk, wq, wv, q_output, B, S, d);
    return launch_qkv_latest_paged_simt(ctx, page_table, lengths, wk, wq, wv, q_output, B, S, d);
}

// logits + decoder.  Tensor-core mode: split-K partial planes in a workspace, summed by the decoder
// (which also writes them to emb_score when the caller wants the logits); exact mode: plain logits.
static int logits_and_decode(mli_ctx* ctx, const float* attn, const float* emb, float* emb_score,
                             const float* pos, float* const* page_table, int* lengths,
                             int* decoder_result, int B, int V, int S, int d, int n_dec, int i_dec) {
    int rc;
    void* p;
    if (use_tc(ctx)) {
        if ((rc = ws_get(ctx, WS_LOGITS, sizeof(float) * (size_t)kMaxLogitSplit * B * V, &p))) return rc;
        float* part = reinterpret_cast<float*>(p);
        int n_split = 1;
        if ((rc = launch_logits_tc(ctx, attn, emb, part, B, V, d, &n_split))) return rc;
        return launch_paged_decoder(ctx, part, n_split, emb_score, decoder_result, lengths, page_table,
                                    pos, emb, B, V, S, d, n_dec, i_dec);
    }
    if (!emb_score) {
        if ((rc = ws_get(ctx, WS_LOGITS, sizeof(float) * (size_t)B * V, &p))) return rc;
        emb_score = reinterpret_cast<float*>(p);
    }
    if ((rc = launch_logits_simt(ctx, attn, emb, emb_score, B, V, d))) return rc;
    return launch_paged_decoder(ctx, emb_score, 1, nullptr, decoder_result, lengths, page_table, pos,
                                emb, B, V, S, d, n_dec, i_dec);
}

}  // namespace mli

using namespace mli;

extern "C" {

const char* mli_last_error(void) { return g_last_error.c_str(); }
const char* mli_version(void) { return "min_llm_inference_b200 0.1 (sm_100a)"; }
long long mli_kernel_launch_count(void) { return g_launches.load(); }

int mli_ctx_create(mli_ctx** out, int device, void* cuda_stream) {
    if (!out) return MLI_ERR_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device: this library has no CPU path");
        return MLI_ERR_CUDA;
    }
    MLI_REQUIRE(device >= 0 && device < n, "bad device ordinal");
    MLI_CUDA(cudaSetDevice(device));
    mli_ctx* ctx = new (std::nothrow) mli_ctx();
    if (!ctx) return MLI_ERR_STATE;
    ctx->device = device;
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    cudaDeviceProp prop;
    MLI_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    ctx->tc_available = (prop.major == 10) && tcgen05_supported(ctx);
    ctx->gemm_mode = ctx->tc_available ? 0 : 1;
    *out = ctx;
    return MLI_OK;
}

int mli_ctx_destroy(mli_ctx* ctx) {
    if (!ctx) return MLI_OK;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->tc_available) tc_unregister_all(ctx);
    for (int i = 0; i < WS_NUM_SLOTS; ++i)
        if (ctx->ws[i]) cudaFree(ctx->ws[i]);
    delete ctx;
    return MLI_OK;
}

int mli_ctx_set_stream(mli_ctx* ctx, void* cuda_stream) {
    MLI_ENTER(ctx, "null ctx");
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return MLI_OK;
}

int mli_ctx_set_option(mli_ctx* ctx, int option, int value) {
    MLI_ENTER(ctx, "null ctx");
    switch (option) {
        case MLI_OPT_GEMM_MODE:
            MLI_REQUIRE(value == 0 || value == 1, "gemm mode must be 0 or 1");
            if (value == 0 && !ctx->tc_available) {
                set_error("tcgen05 GEMM path is not available on this device/build");
                return MLI_ERR_UNSUPPORTED;
            }
            ctx->gemm_mode = value;
            return MLI_OK;
        case MLI_OPT_ATTN_CHUNK_PAGES:
            MLI_REQUIRE(value >= 0 && value <= 32, "chunk pages must be in [0,32]");
            ctx->attn_chunk_pages = value;
            return MLI_OK;
        case MLI_OPT_ATTN_CTAS_PER_SM:
            MLI_REQUIRE(value >= 0 && value <= 2, "CTAs per SM must be in [0,2]");
            ctx->attn_ctas_per_sm = value;
            return MLI_OK;
        case MLI_OPT_PDL:
            MLI_REQUIRE(value == 0 || value == 1, "pdl must be 0 or 1");
            ctx->opt_pdl = value;
            return MLI_OK;
        case MLI_OPT_KV_FORMAT:
            MLI_REQUIRE(value == 0 || value == 1, "kv format must be 0 (fp32 pages) or 1 (bf16 K/V)");
            if (value == 1 && !ctx->tc_available) {
                set_error("the compact KV format needs the tcgen05 GEMM path");
                return MLI_ERR_UNSUPPORTED;
            }
            ctx->kv_bf16 = value;
            return MLI_OK;
        case MLI_OPT_ATTN_KERNEL:
            MLI_REQUIRE(value >= 0 && value <= 2, "attention kernel must be 0 (auto), 1 or 2");
            ctx->attn_kernel = value;
            return MLI_OK;
        case MLI_OPT_ATTN_MIN_DYN:
            MLI_REQUIRE(value >= 1, "the dynamic-slice threshold must be positive");
            ctx->attn_min_dyn = value;
            return MLI_OK;
    }
    set_error("unknown option");
    return MLI_ERR_ARG;
}

int mli_ctx_get_option(mli_ctx* ctx, int option, int* value) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(value, "null argument");
    switch (option) {
        case MLI_OPT_GEMM_MODE: *value = ctx->gemm_mode; return MLI_OK;
        case MLI_OPT_ATTN_CHUNK_PAGES: *value = ctx->attn_chunk_pages; return MLI_OK;
        case MLI_OPT_ATTN_CTAS_PER_SM: *value = ctx->attn_ctas_per_sm; return MLI_OK;
        case MLI_OPT_PDL: *value = ctx->opt_pdl; return MLI_OK;
        case MLI_OPT_KV_FORMAT: *value = ctx->kv_bf16; return MLI_OK;
        case MLI_OPT_ATTN_KERNEL: *value = ctx->attn_kernel; return MLI_OK;
        case MLI_OPT_ATTN_MIN_DYN: *value = ctx->attn_min_dyn; return MLI_OK;
    }
    set_error("unknown option");
    return MLI_ERR_ARG;
}

int mli_ctx_register_weights(mli_ctx* ctx, const float* wk, const float* wq, const float* wv,
                             const float* emb_table, int emb_dim, int n_vocab) {
    MLI_ENTER(ctx, "null ctx");
    if (!ctx->tc_available) return MLI_OK;  // the SIMT path reads the weights in place
    return tc_register_weights(ctx, wk, wq, wv, emb_table, emb_dim, n_vocab);
}

int mli_ctx_unregister_weights(mli_ctx* ctx) {
    MLI_ENTER(ctx, "null ctx");
    if (ctx->tc_available) tc_unregister_all(ctx);
    return MLI_OK;
}

int mli_debug_set_gemm_stamps(mli_ctx* ctx, void* stamps_dev) {
    MLI_ENTER(ctx, "null ctx");
    ctx->tc_dbg = stamps_dev;
    return MLI_OK;
}

int mli_debug_last_gemm_plan(mli_ctx* ctx, int kind, int* plan) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(plan != nullptr && kind >= 0 && kind < 4, "kind must be 0..3 and plan non-null");
    for (int i = 0; i < 5; ++i) plan[i] = ctx->tc_last_plan[kind][i];
    return MLI_OK;
}

int mli_debug_set_step_trace(mli_ctx* ctx, void* trace_dev) {
    MLI_ENTER(ctx, "null ctx");
    ctx->trace = reinterpret_cast<unsigned long long*>(trace_dev);
    return MLI_OK;
}

int mli_ctx_synchronize(mli_ctx* ctx) {
    MLI_ENTER(ctx, "null ctx");
    MLI_CUDA(cudaStreamSynchronize(ctx->stream));
    return MLI_OK;
}

// ---- paged stages ---------------------------------------------------------------------------
int mli_paged_encoder(mli_ctx* ctx, const float* emb_table, const float* pos_table, const int* inp,
                      float** page_table, const int* lengths, const int* new_item_indices,
                      int n_batch, int n_sequence, int emb_dim, int n_new_items) {
    MLI_ENTER(ctx, "null ctx");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    if (n_new_items <= 0) return MLI_OK;  // encoder.cu:138-140
    const int max_tiles = n_new_items * ceil_div(n_sequence, kTileM);
    int* n_tiles;
    TileDesc* tiles;
    if ((rc = tiles_ws(ctx, max_tiles, &n_tiles, &tiles))) return rc;
    if ((rc = launch_build_new_row_tiles(ctx, new_item_indices, lengths, n_new_items, nullptr, tiles,
                                         n_tiles, max_tiles)))
        return rc;
    return launch_paged_encoder_tiles(ctx, emb_table, pos_table, inp, nullptr, nullptr, page_table,
                                      tiles, n_tiles, max_tiles, lengths, n_sequence, emb_dim);
}

int mli_prefill_kv_paged(mli_ctx* ctx, float** page_table, const int* new_batch_idx,
                         const int* lengths, const float* wk, const float* wv, int n_new_items,
                         int n_batch, int n_sequence, int emb_dim) {
    MLI_ENTER(ctx, "null ctx");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    if (n_new_items <= 0) return MLI_OK;  // paged_attention.cu:100-102
    const int max_tiles = n_new_items * ceil_div(n_sequence, kTileM);
    int* n_tiles;
    TileDesc* tiles;
    if ((rc = tiles_ws(ctx, max_tiles, &n_tiles, &tiles))) return rc;
    if ((rc = launch_build_new_row_tiles(ctx, new_batch_idx, lengths, n_new_items, nullptr, tiles,
                                         n_tiles, max_tiles)))
        return rc;
    return prefill_paged(ctx, page_table, tiles, n_tiles, max_tiles, lengths, wk, wv, n_sequence,
                         emb_dim);
}

int mli_qkv_latest_paged(mli_ctx* ctx, float** page_table, const int* lengths, const float* wk,
                         const float* wq, const float* wv, float* q_output, int n_batch,
                         int n_sequence, int emb_dim) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(q_output, "null argument");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    return latest_paged(ctx, page_table, lengths, wk, wq, wv, q_output, n_batch, n_sequence, emb_dim);
}

int mli_decode_attention_paged(mli_ctx* ctx, const float* q, float* const* page_table,
                               const int* lengths, float* attention_result, float* softmax_out,
                               int n_batch, int n_sequence, int emb_dim) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(q && attention_result, "null argument");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    return launch_decode_attention_paged(ctx, q, page_table, lengths, attention_result, softmax_out,
                                         n_batch, n_sequence, emb_dim);
}

int mli_qkt_paged(mli_ctx* ctx, const float* q, float* const* page_table, const int* lengths,
                  float* qkt_output, int n_batch, int n_sequence, int emb_dim) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(q && page_table && lengths && qkt_output, "null argument");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    MLI_REQUIRE(!ctx->kv_bf16, "the unfused stages read the reference's fp32 page format only");
    return launch_qkt_unfused(ctx, q, page_table, lengths, qkt_output, n_batch, n_sequence, emb_dim);
}

int mli_softmax_in_place_with_lengths(mli_ctx* ctx, float* qkt_output, const int* lengths, int n_batch,
                                      int n_sequence) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(qkt_output && lengths && n_batch > 0 && n_sequence > 0, "bad argument");
    return launch_softmax_lengths_unfused(ctx, qkt_output, lengths, n_batch, n_sequence);
}

int mli_softmax_v_paged(mli_ctx* ctx, const float* softmax_result, float* const* page_table,
                        float* attention_result, const int* lengths, int n_batch, int n_sequence,
                        int emb_dim) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(softmax_result && page_table && attention_result && lengths, "null argument");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    MLI_REQUIRE(!ctx->kv_bf16, "the unfused stages read the reference's fp32 page format only");
    return launch_softmax_v_unfused(ctx, softmax_result, page_table, attention_result, lengths, n_batch,
                                    n_sequence, emb_dim);
}

int mli_paged_attention(mli_ctx* ctx, float** page_table, const int* lengths, const float* wk,
                        const float* wq, const float* wv, const int* new_batch_idx, float* q_output,
                        float* qkt_output, float* attention_result, int n_new_items, int n_batch,
                        int n_sequence, int emb_dim) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(attention_result, "null argument");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    if (!q_output) {
        void* p;
        if ((rc = ws_get(ctx, WS_QOUT, sizeof(float) * (size_t)n_batch * emb_dim, &p))) return rc;
        q_output = reinterpret_cast<float*>(p);
    }
    if ((rc = mli_prefill_kv_paged(ctx, page_table, new_batch_idx, lengths, wk, wv, n_new_items,
                                   n_batch, n_sequence, emb_dim)))
        return rc;
    if ((rc = latest_paged(ctx, page_table, lengths, wk, wq, wv, q_output, n_batch, n_sequence,
                           emb_dim)))
        return rc;
    return launch_decode_attention_paged(ctx, q_output, page_table, lengths, attention_result,
                                         qkt_output, n_batch, n_sequence, emb_dim);
}

int mli_paged_decoder(mli_ctx* ctx, const float* batch_result, const float* emb_table,
                      float* emb_score, const float* pos_table, float** page_table, int* lengths,
                      int* decoder_result, int n_batch, int n_vocab, int n_sequence, int emb_dim,
                      int n_decoder_results, int i_decoder) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(batch_result && decoder_result, "null argument");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    MLI_REQUIRE(n_vocab > 0 && n_decoder_results > 0 && i_decoder >= 0 &&
                    i_decoder < n_decoder_results,
                "bad decoder dims");
    return logits_and_decode(ctx, batch_result, emb_table, emb_score, pos_table, page_table, lengths,
                             decoder_result, n_batch, n_vocab, n_sequence, emb_dim, n_decoder_results,
                             i_decoder);
}

int mli_paged_forward(mli_ctx* ctx, const int* inp, int* lengths, const int* new_item_indices,
                      int* decoder_result, int n_new_items, const float* emb_table,
                      const float* pos_table, float** page_table, const float* wk, const float* wq,
                      const float* wv, float* q_output, float* attention_result, int n_batch,
                      int n_sequence, int emb_dim, int n_vocab, int n_forward_rounds) {
    MLI_ENTER(ctx, "null ctx");
    int rc = check_paged_dims(n_batch, n_sequence, emb_dim);
    if (rc) return rc;
    MLI_REQUIRE(n_forward_rounds >= 1 && n_forward_rounds <= kPage, "n_forward_rounds must be 1..16");
    void* p;
    if (!q_output) {
        if ((rc = ws_get(ctx, WS_QOUT, sizeof(float) * (size_t)n_batch * emb_dim, &p))) return rc;
        q_output = reinterpret_cast<float*>(p);
    }
    if (!attention_result) {
        if ((rc = ws_get(ctx, WS_ATTN_OUT, sizeof(float) * (size_t)n_batch * emb_dim, &p))) return rc;
        attention_result = reinterpret_cast<float*>(p);
    }
    for (int round = 0; round < n_forward_rounds; ++round) {
        if (round == 0 && n_new_items > 0) {
            const int max_tiles = n_new_items * ceil_div(n_sequence, kTileM);
            int* n_tiles;
            TileDesc* tiles;
            if ((rc = tiles_ws(ctx, max_tiles, &n_tiles, &tiles))) return rc;
            if ((rc = launch_build_new_row_tiles(ctx, new_item_indices, lengths, n_new_items, nullptr,
                                                 tiles, n_tiles, max_tiles)))
                return rc;
            if ((rc = launch_paged_encoder_tiles(ctx, emb_table, pos_table, inp, nullptr, nullptr,
                                                 page_table, tiles, n_tiles, max_tiles, lengths,
                                                 n_sequence, emb_dim)))
                return rc;
            if ((rc = prefill_paged(ctx, page_table, tiles, n_tiles, max_tiles, lengths, wk, wv,
                                    n_sequence, emb_dim)))
                return rc;
        }
        if ((rc = latest_paged(ctx, page_table, lengths, wk, wq, wv, q_output, n_batch, n_sequence,
                               emb_dim)))
            return rc;
        if ((rc = launch_decode_attention_paged(ctx, q_output, page_table, lengths, attention_result,
                                                nullptr, n_batch, n_sequence, emb_dim)))
            return rc;
        if ((rc = logits_and_decode(ctx, attention_result, emb_table, nullptr, pos_table, page_table,
                                    lengths, decoder_result, n_batch, n_vocab, n_sequence, emb_dim,
                                    n_forward_rounds, round)))
            return rc;
    }
    return MLI_OK;
}

// ---- dense stages ---------------------------------------------------------------------------
int mli_dense_encoder(mli_ctx* ctx, const float* emb_table, const float* pos_table, const int* inp,
                      float* inp_embedding, const int* lengths, const int* new_item_indices,
                      int n_batch, int n_sequence, int emb_dim, int n_new_items) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(n_batch > 0 && n_sequence > 0 && emb_dim > 0, "bad dims");
    return launch_dense_encoder(ctx, emb_table, pos_table, inp, inp_embedding, lengths,
                                new_item_indices, n_sequence, emb_dim, n_new_items);
}

int mli_self_attention(mli_ctx* ctx, const float* inp_embedding, const int* lengths, const float* wk,
                       const float* wq, const float* wv, const int* new_batch_idx, float* kt_cache,
                       float* v_cache, float* q_output, float* qkt_output, float* attention_result,
                       int n_new_items, int n_batch, int n_sequence, int input_dim, int output_dim) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(q_output && attention_result, "null argument");
    MLI_REQUIRE(n_batch > 0 && n_sequence > 0 && input_dim > 0 && output_dim > 0, "bad dims");
    int rc;
    if (n_new_items > 0) {
        const int max_tiles = n_new_items * ceil_div(n_sequence, kTileM);
        int* n_tiles;
        TileDesc* tiles;
        if ((rc = tiles_ws(ctx, max_tiles, &n_tiles, &tiles))) return rc;
        if ((rc = launch_build_new_row_tiles(ctx, new_batch_idx, lengths, n_new_items, nullptr, tiles,
                                             n_tiles, max_tiles)))
            return rc;
        if ((rc = launch_prefill_kv_dense_simt(ctx, inp_embedding, tiles, n_tiles, max_tiles, lengths,
                                               wk, wv, kt_cache, v_cache, n_sequence, input_dim,
                                               output_dim)))
            return rc;
    }
    if ((rc = launch_qkv_latest_dense_simt(ctx, inp_embedding, lengths, wk, wq, wv, kt_cache, v_cache,
                                           q_output, n_batch, n_sequence, input_dim, output_dim)))
        return rc;
    return launch_decode_attention_dense(ctx, q_output, kt_cache, v_cache, lengths, attention_result,
                                         qkt_output, n_batch, n_sequence, output_dim);
}

int mli_dense_decoder(mli_ctx* ctx, const float* batch_result, const float* emb_table,
                      float* emb_score, const float* pos_table, float* inp_embedding, int* lengths,
                      int* decoder_result, int n_batch, int n_vocab, int n_sequence, int emb_dim) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(batch_result && decoder_result, "null argument");
    int rc;
    if (!emb_score) {
        void* p;
        if ((rc = ws_get(ctx, WS_LOGITS, sizeof(float) * (size_t)n_batch * n_vocab, &p))) return rc;
        emb_score = reinterpret_cast<float*>(p);
    }
    if ((rc = launch_logits_simt(ctx, batch_result, emb_table, emb_score, n_batch, n_vocab, emb_dim)))
        return rc;
    return launch_dense_decoder(ctx, emb_score, decoder_result, lengths, inp_embedding, pos_table,
                                emb_table, n_batch, n_vocab, n_sequence, emb_dim);
}

int mli_dense_forward(mli_ctx* ctx, const int* inp, int* lengths, const int* new_item_indices,
                      int* decoder_result, int n_new_items, const float* emb_table,
                      const float* pos_table, const float* wk, const float* wq, const float* wv,
                      float* inp_embedding, float* kt_cache, float* v_cache, float* q_output,
                      float* attention_result, int n_batch, int n_sequence, int emb_dim, int n_vocab) {
    MLI_ENTER(ctx, "null ctx");
    int rc;
    void* p;
    if (!q_output) {
        if ((rc = ws_get(ctx, WS_QOUT, sizeof(float) * (size_t)n_batch * emb_dim, &p))) return rc;
        q_output = reinterpret_cast<float*>(p);
    }
    if (!attention_result) {
        if ((rc = ws_get(ctx, WS_ATTN_OUT, sizeof(float) * (size_t)n_batch * emb_dim, &p))) return rc;
        attention_result = reinterpret_cast<float*>(p);
    }
    if ((rc = mli_dense_encoder(ctx, emb_table, pos_table, inp, inp_embedding, lengths,
                                new_item_indices, n_batch, n_sequence, emb_dim, n_new_items)))
        return rc;
    if ((rc = mli_self_attention(ctx, inp_embedding, lengths, wk, wq, wv, new_item_indices, kt_cache,
                                 v_cache, q_output, nullptr, attention_result, n_new_items, n_batch,
                                 n_sequence, emb_dim, emb_dim)))
        return rc;
    return mli_dense_decoder(ctx, attention_result, emb_table, nullptr, pos_table, inp_embedding,
                             lengths, decoder_result, n_batch, n_vocab, n_sequence, emb_dim);
}

}  // extern "C"
