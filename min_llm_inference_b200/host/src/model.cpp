// model.cpp -- kernel-level launchers, layers and models of the reference API, each a thin call
// into the C ABI (include/mli_b200.h).  Shapes are read from the Tensors exactly where the reference
// reads them (src/layers.cpp, src/kernels/*.cu launchers).
#include <cassert>

#include "mli/compat.hpp"

namespace {
inline mli_ctx* C() { return mli::host_context(); }
inline int i(size_t v) { return static_cast<int>(v); }
}  // namespace

// ---- launchers --------------------------------------------------------------------------------------
// src/kernels/paged_attention.cu:358-377
void paged_attention(TensorFloatPoint& page_table, const TensorInt& lengths, const TensorFloat& wk,
                     const TensorFloat& wq, const TensorFloat& wv, const TensorInt& new_batch_idx,
                     TensorFloat& q_output, TensorFloat& qkt_output, TensorFloat& attention_result,
                     int n_new_items, int n_sequence) {
    const int B = i(page_table.shape()[0]), d = i(wk.shape()[0]);
    mli::check(mli_paged_attention(C(), page_table.data(), lengths.data(), wk.data(), wq.data(),
                                   wv.data(), new_batch_idx.data(), q_output.data(),
                                   qkt_output.data(), attention_result.data(), n_new_items, B,
                                   n_sequence, d));
}

// src/kernels/paged_attention.cu:96-115
void launch_fill_new_k_v_cache_paged_attention(TensorFloatPoint page_table,
                                               const TensorInt& new_batch_idx,
                                               const TensorInt& lengths, const TensorFloat& wk,
                                               const TensorFloat& wv, int n_new_items,
                                               int n_sequence) {
    const int B = i(page_table.shape()[0]), d = i(wk.shape()[0]);
    assert(i(page_table.shape()[1]) == n_sequence / PAGE_BLOCK_SIZE);
    mli::check(mli_prefill_kv_paged(C(), page_table.data(), new_batch_idx.data(), lengths.data(),
                                    wk.data(), wv.data(), n_new_items, B, n_sequence, d));
}

// src/kernels/paged_attention_cublas.cu:225-246 -- same contract, one implementation
void launch_fill_new_k_v_cache_paged_attention_warp_tiling(TensorFloatPoint page_table,
                                                           const TensorInt& new_batch_idx,
                                                           const TensorInt& lengths,
                                                           const TensorFloat& wk, const TensorFloat& wv,
                                                           int n_new_items, int n_sequence) {
    launch_fill_new_k_v_cache_paged_attention(page_table, new_batch_idx, lengths, wk, wv, n_new_items,
                                              n_sequence);
}

// src/kernels/paged_attention.cu:188-199
void launch_get_latest_k_q_v_paged_attention(TensorFloatPoint& page_table, const TensorInt& lengths,
                                             const TensorFloat& wk, const TensorFloat& wq,
                                             const TensorFloat& wv, TensorFloat& q_output,
                                             int n_sequence) {
    const int B = i(page_table.shape()[0]), d = i(wq.shape()[0]);
    mli::check(mli_qkv_latest_paged(C(), page_table.data(), lengths.data(), wk.data(), wq.data(),
                                    wv.data(), q_output.data(), B, n_sequence, d));
}

// src/kernels/paged_attention_cublas.cu:76-99; latest_emb / temp_placeholder / handle are scratch the
// fused GEMM epilogue does not need
void launch_get_latest_k_q_v_paged_attention_cublas(TensorFloatPoint& page_table,
                                                    const TensorInt& lengths, TensorFloat&,
                                                    const TensorFloat& wk, const TensorFloat& wq,
                                                    const TensorFloat& wv, TensorFloat& q_output,
                                                    TensorFloat&, cublasHandle_t&, int n_sequence) {
    launch_get_latest_k_q_v_paged_attention(page_table, lengths, wk, wq, wv, q_output, n_sequence);
}

void launch_fused_decode_attention(const TensorFloat& q_output, const TensorFloatPoint& page_table,
                                   const TensorInt& lengths, TensorFloat& attention_result,
                                   TensorFloat* softmax_result, int n_sequence) {
    const int B = i(q_output.shape()[0]), d = i(q_output.shape()[1]);
    mli::check(mli_decode_attention_paged(C(), q_output.data(), page_table.data(), lengths.data(),
                                          attention_result.data(),
                                          softmax_result ? softmax_result->data() : nullptr, B,
                                          n_sequence, d));
}

// the reference's three unfused stages (src/kernels/paged_attention.cu:208-345,
// src/kernels/self_attention_inference_optimized.cu:360-368), same signatures
void launch_qkt_paged_attention(const TensorFloat& q_output, const TensorFloatPoint& page_table,
                                const TensorInt& lengths, TensorFloat& qkt_output) {
    const int B = i(q_output.shape()[0]), d = i(q_output.shape()[1]), S = i(qkt_output.shape()[1]);
    mli::check(mli_qkt_paged(C(), q_output.data(), page_table.data(), lengths.data(), qkt_output.data(), B, S, d));
}

void launch_softmax_in_place_with_lengths(TensorFloat& qkt_output, const TensorInt& lengths) {
    mli::check(mli_softmax_in_place_with_lengths(C(), qkt_output.data(), lengths.data(), i(qkt_output.shape()[0]),
                                                 i(qkt_output.shape()[1])));
}

void launch_softmax_v_paged_attention(const TensorFloat& softmax_result, const TensorFloatPoint& page_table,
                                      TensorFloat& attention_result, const TensorInt& lengths) {
    const int B = i(softmax_result.shape()[0]), S = i(softmax_result.shape()[1]), d = i(attention_result.shape()[1]);
    mli::check(mli_softmax_v_paged(C(), softmax_result.data(), page_table.data(), attention_result.data(),
                                   lengths.data(), B, S, d));
}

// src/kernels/paged_attention_cublas.cu:260-280
void paged_attention_with_cublas(TensorFloatPoint& page_table, const TensorInt& lengths,
                                 const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv,
                                 const TensorInt& new_batch_idx, TensorFloat& q_output,
                                 TensorFloat& qkt_output, TensorFloat& attention_result, TensorFloat&,
                                 TensorFloat&, int n_new_items, int n_sequence, cublasHandle_t&) {
    paged_attention(page_table, lengths, wk, wq, wv, new_batch_idx, q_output, qkt_output,
                    attention_result, n_new_items, n_sequence);
}

// src/kernels/decoder.cu:94-112
void launch_decoder(const TensorFloat& batch_result, const TensorFloat& emb_table,
                    TensorFloat& emb_score, const TensorFloat& wpe_table, TensorFloat& inp_embedding,
                    TensorInt& lengths, TensorInt& decoder_result) {
    const int B = i(batch_result.shape()[0]), d = i(batch_result.shape()[1]);
    const int V = i(emb_table.shape()[0]), S = i(wpe_table.shape()[0]);
    mli::check(mli_dense_decoder(C(), batch_result.data(), emb_table.data(), emb_score.data(),
                                 wpe_table.data(), inp_embedding.data(), lengths.data(),
                                 decoder_result.data(), B, V, S, d));
}

// src/kernels/decoder.cu:207-229
void launch_paged_attention_decoder_multi_rounds(const TensorFloat& batch_result,
                                                 const TensorFloat& emb_table, TensorFloat& emb_score,
                                                 const TensorFloat& wpe_table,
                                                 TensorFloatPoint& page_table, TensorInt& lengths,
                                                 TensorInt& decoder_result, int i_decoder) {
    const int B = i(batch_result.shape()[0]), d = i(batch_result.shape()[1]);
    const int V = i(emb_table.shape()[0]), S = i(wpe_table.shape()[0]);
    const int n_dec = decoder_result.shape().size() == 2 ? i(decoder_result.shape()[1]) : 1;
    mli::check(mli_paged_decoder(C(), batch_result.data(), emb_table.data(), emb_score.data(),
                                 wpe_table.data(), page_table.data(), lengths.data(),
                                 decoder_result.data(), B, V, S, d, n_dec, i_decoder));
}

// src/kernels/decoder.cu:232-255
void launch_paged_attention_cublas_decoder_multi_rounds(
    const TensorFloat& batch_result, const TensorFloat& emb_table, TensorFloat& emb_score,
    const TensorFloat& wpe_table, TensorFloatPoint& page_table, TensorInt& lengths,
    TensorInt& decoder_result, int i_decoder, cublasHandle_t&) {
    launch_paged_attention_decoder_multi_rounds(batch_result, emb_table, emb_score, wpe_table,
                                                page_table, lengths, decoder_result, i_decoder);
}

// src/kernels/encoder.cu:80-92 and :134-147 take raw pointers in the reference too
void launch_inference_optimized_encoder_kernel(const float* emb_table, const float* wpe,
                                               const int* inp, float* inp_embedding,
                                               const int* lengths, const int* new_item_indices,
                                               int batch_size, int n_sequence, int embedding_dim,
                                               int n_new_items) {
    mli::check(mli_dense_encoder(C(), emb_table, wpe, inp, inp_embedding, lengths, new_item_indices,
                                 batch_size, n_sequence, embedding_dim, n_new_items));
}

void launch_paged_attention_encoder_kernel(const float* emb_table, const float* wpe, const int* inp,
                                           float** page_table, const int* lengths,
                                           const int* new_item_indices, int batch_size,
                                           int n_sequence, int embedding_dim, int n_new_items) {
    mli::check(mli_paged_encoder(C(), emb_table, wpe, inp, page_table, lengths, new_item_indices,
                                 batch_size, n_sequence, embedding_dim, n_new_items));
}

// src/kernels/self_attention_inference_optimized.cu:282-301
void inference_self_attention(const TensorFloat& inp_embedding, const TensorInt& lengths,
                              const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv,
                              const TensorInt& new_batch_idx, TensorFloat& kt_cache,
                              TensorFloat& v_cache, TensorFloat& q_output, TensorFloat& qkt_output,
                              TensorFloat& attention_result, int n_new_items) {
    const int B = i(inp_embedding.shape()[0]), S = i(inp_embedding.shape()[1]);
    const int di = i(inp_embedding.shape()[2]), dn = i(wk.shape()[1]);
    mli::check(mli_self_attention(C(), inp_embedding.data(), lengths.data(), wk.data(), wq.data(),
                                  wv.data(), new_batch_idx.data(), kt_cache.data(), v_cache.data(),
                                  q_output.data(), qkt_output.data(), attention_result.data(),
                                  n_new_items, B, S, di, dn));
}

// ---- layers (src/layers.cpp:54-154) ------------------------------------------------------------------------
SelfAttentionLayer::SelfAttentionLayer(TensorFloat&& wk, TensorFloat&& wq, TensorFloat&& wv,
                                       size_t n_batch, size_t input_dim, size_t n_sequence)
    : wk_(std::move(wk)), wq_(std::move(wq)), wv_(std::move(wv)),
      kt_cache_({n_batch, input_dim, n_sequence}, DeviceType::DEVICE),
      v_cache_({n_batch, n_sequence, input_dim}, DeviceType::DEVICE),
      q_output_({n_batch, input_dim}, DeviceType::DEVICE),
      qkt_output_({n_batch, n_sequence}, DeviceType::DEVICE) {}

void SelfAttentionLayer::forward(const TensorFloat& inp_embedding, const TensorInt& lengths,
                                 const TensorInt& new_batch_idx, TensorFloat& attention_result,
                                 int n_new_items) {
    inference_self_attention(inp_embedding, lengths, wk_, wq_, wv_, new_batch_idx, kt_cache_, v_cache_,
                             q_output_, qkt_output_, attention_result, n_new_items);
}

PagedAttentionLayer::PagedAttentionLayer(TensorFloat&& wk, TensorFloat&& wq, TensorFloat&& wv,
                                         size_t n_batch, size_t emb_dim, size_t n_sequence)
    : wk_(std::move(wk)), wq_(std::move(wq)), wv_(std::move(wv)),
      q_output_({n_batch, emb_dim}, DeviceType::DEVICE),
      qkt_output_({n_batch, n_sequence}, DeviceType::DEVICE) {}

void PagedAttentionLayer::forward(TensorFloatPoint& page_table, const TensorInt& lengths,
                                  const TensorInt& new_batch_idx, TensorFloat& attention_result,
                                  int n_new_items) {
    paged_attention(page_table, lengths, wk_, wq_, wv_, new_batch_idx, q_output_, qkt_output_,
                    attention_result, n_new_items, i(qkt_output_.shape()[1]));
}

PagedAttentionCublasLayer::PagedAttentionCublasLayer(TensorFloat&& wk, TensorFloat&& wq,
                                                     TensorFloat&& wv, size_t n_batch, size_t emb_dim,
                                                     size_t n_sequence)
    : wk_(std::move(wk)), wq_(std::move(wq)), wv_(std::move(wv)),
      q_output_({n_batch, emb_dim}, DeviceType::DEVICE),
      qkt_output_({n_batch, n_sequence}, DeviceType::DEVICE),
      latest_emb_({1}, DeviceType::DEVICE), temp_placeholder_({1}, DeviceType::DEVICE) {}

void PagedAttentionCublasLayer::forward(TensorFloatPoint& page_table, const TensorInt& lengths,
                                        const TensorInt& new_batch_idx, TensorFloat& attention_result,
                                        int n_new_items, cublasHandle_t& handle) {
    paged_attention_with_cublas(page_table, lengths, wk_, wq_, wv_, new_batch_idx, q_output_,
                                qkt_output_, attention_result, latest_emb_, temp_placeholder_,
                                n_new_items, i(qkt_output_.shape()[1]), handle);
}

void EncoderLayer::forward(const TensorFloat& emb_table, const TensorFloat& pos_emb,
                           const TensorInt& inp, TensorFloat& inp_embedding, const TensorInt& lengths,
                           const TensorInt& new_item_indices, int n_new_items) {
    launch_inference_optimized_encoder_kernel(emb_table.data(), pos_emb.data(), inp.data(),
                                              inp_embedding.data(), lengths.data(),
                                              new_item_indices.data(), i(inp_embedding.shape()[0]),
                                              i(inp_embedding.shape()[1]), i(inp_embedding.shape()[2]),
                                              n_new_items);
}

void PagedEncoderLayer::forward(const TensorFloat& emb_table, const TensorFloat& pos_emb,
                                const TensorInt& inp, TensorFloatPoint& page_table,
                                const TensorInt& lengths, const TensorInt& new_item_indices,
                                int n_new_items) {
    launch_paged_attention_encoder_kernel(emb_table.data(), pos_emb.data(), inp.data(),
                                          page_table.data(), lengths.data(), new_item_indices.data(),
                                          i(inp.shape()[0]), i(inp.shape()[1]), i(emb_table.shape()[1]),
                                          n_new_items);
}

DecoderLayer::DecoderLayer(size_t n_batch, size_t n_vocab)
    : emb_score_({n_batch, n_vocab}, DeviceType::DEVICE) {}
void DecoderLayer::forward(const TensorFloat& batch_result, const TensorFloat& emb_table,
                           const TensorFloat& wpe_table, TensorFloat& inp_embedding, TensorInt& lengths,
                           TensorInt& decoder_result) {
    launch_decoder(batch_result, emb_table, emb_score_, wpe_table, inp_embedding, lengths, decoder_result);
}

PagedDecoderLayer::PagedDecoderLayer(size_t n_batch, size_t n_vocab)
    : emb_score_({n_batch, n_vocab}, DeviceType::DEVICE) {}
void PagedDecoderLayer::forward(const TensorFloat& batch_result, const TensorFloat& emb_table,
                                const TensorFloat& wpe_table, TensorFloatPoint& page_table,
                                TensorInt& lengths, TensorInt& decoder_result, int i_decoder_round) {
    launch_paged_attention_decoder_multi_rounds(batch_result, emb_table, emb_score_, wpe_table,
                                                page_table, lengths, decoder_result, i_decoder_round);
}

PagedCublasDecoderLayer::PagedCublasDecoderLayer(size_t n_batch, size_t n_vocab)
    : emb_score_({n_batch, n_vocab}, DeviceType::DEVICE) {}
void PagedCublasDecoderLayer::forward(const TensorFloat& batch_result, const TensorFloat& emb_table,
                                      const TensorFloat& wpe_table, TensorFloatPoint& page_table,
                                      TensorInt& lengths, TensorInt& decoder_result,
                                      int i_decoder_round, cublasHandle_t& handle) {
    launch_paged_attention_cublas_decoder_multi_rounds(batch_result, emb_table, emb_score_, wpe_table,
                                                       page_table, lengths, decoder_result,
                                                       i_decoder_round, handle);
}

// ---- models (src/inference_model.cpp) ----------------------------------------------------------------------
InferenceModel::InferenceModel(SelfAttentionLayer&& attention, EncoderLayer&& encoder,
                               DecoderLayer&& decoder, size_t n_batch, size_t n_sequence,
                               size_t emb_dim)
    : attention_layer_(std::move(attention)), encoder_layer_(std::move(encoder)),
      decoder_layer_(std::move(decoder)), n_batch_(n_batch), n_sequence_(n_sequence),
      emb_dim_(emb_dim), inp_embedding_({n_batch, n_sequence, emb_dim}, DeviceType::DEVICE),
      attention_result_({n_batch, emb_dim}, DeviceType::DEVICE) {}

void InferenceModel::forward(const TensorInt& inp, TensorInt& lengths,
                             const TensorInt& new_item_indices, TensorInt& decoder_result,
                             int n_new_items, const TensorFloat& emb_table,
                             const TensorFloat& pos_emb_table) {
    encoder_layer_.forward(emb_table, pos_emb_table, inp, inp_embedding_, lengths, new_item_indices,
                           n_new_items);
    attention_layer_.forward(inp_embedding_, lengths, new_item_indices, attention_result_, n_new_items);
    decoder_layer_.forward(attention_result_, emb_table, pos_emb_table, inp_embedding_, lengths,
                           decoder_result);
}

PagedAttentionInferenceModel::PagedAttentionInferenceModel(PagedAttentionLayer&& attention,
                                                           PagedEncoderLayer&& encoder,
                                                           PagedDecoderLayer&& decoder, size_t n_batch,
                                                           size_t n_sequence, size_t emb_dim,
                                                           int n_forward_rounds)
    : paged_attention_layer_(std::move(attention)), paged_encoder_layer_(std::move(encoder)),
      paged_decoder_layer_(std::move(decoder)), n_batch_(n_batch), n_sequence_(n_sequence),
      emb_dim_(emb_dim), attention_result_({n_batch, emb_dim}, DeviceType::DEVICE),
      n_forward_rounds_(n_forward_rounds) {}

void PagedAttentionInferenceModel::forward(const TensorInt& inp, TensorInt& lengths,
                                           const TensorInt& new_item_indices,
                                           TensorInt& decoder_result, int n_new_items,
                                           const TensorFloat& emb_table,
                                           const TensorFloat& pos_emb_table,
                                           TensorFloatPoint& page_table) {
    // rounds after the first see no new rows (src/inference_model.cpp:56-59)
    for (int round = 0; round < n_forward_rounds_; ++round) {
        const int fresh = round == 0 ? n_new_items : 0;
        paged_encoder_layer_.forward(emb_table, pos_emb_table, inp, page_table, lengths,
                                     new_item_indices, fresh);
        paged_attention_layer_.forward(page_table, lengths, new_item_indices, attention_result_, fresh);
        paged_decoder_layer_.forward(attention_result_, emb_table, pos_emb_table, page_table, lengths,
                                     decoder_result, round);
    }
}

PagedAttentionCublasInferenceModel::PagedAttentionCublasInferenceModel(
    PagedAttentionCublasLayer&& attention, PagedEncoderLayer&& encoder,
    PagedCublasDecoderLayer&& decoder, size_t n_batch, size_t n_sequence, size_t emb_dim,
    int n_forward_rounds)
    : paged_attention_layer_(std::move(attention)), paged_encoder_layer_(std::move(encoder)),
      paged_decoder_layer_(std::move(decoder)), n_batch_(n_batch), n_sequence_(n_sequence),
      emb_dim_(emb_dim), attention_result_({n_batch, emb_dim}, DeviceType::DEVICE),
      n_forward_rounds_(n_forward_rounds) {}

void PagedAttentionCublasInferenceModel::forward(const TensorInt& inp, TensorInt& lengths,
                                                 const TensorInt& new_item_indices,
                                                 TensorInt& decoder_result, int n_new_items,
                                                 const TensorFloat& emb_table,
                                                 const TensorFloat& pos_emb_table,
                                                 TensorFloatPoint& page_table, cublasHandle_t handle) {
    for (int round = 0; round < n_forward_rounds_; ++round) {
        const int fresh = round == 0 ? n_new_items : 0;
        paged_encoder_layer_.forward(emb_table, pos_emb_table, inp, page_table, lengths,
                                     new_item_indices, fresh);
        paged_attention_layer_.forward(page_table, lengths, new_item_indices, attention_result_, fresh,
                                       handle);
        paged_decoder_layer_.forward(attention_result_, emb_table, pos_emb_table, page_table, lengths,
                                     decoder_result, round, handle);
    }
}
