// ref_shim.cu -- C-ABI window onto the UNMODIFIED reference implementation.
//
// TEST INFRASTRUCTURE ONLY.  This file is OUR code; it #includes the reference's headers from
// /root/reference/include (nothing is copied) and is linked with the reference's own sources,
// compiled where they lie, into oracle/_ref/libmli_ref.so (see oracle/Makefile).  It lets the
// parity tests and bench.py drive the reference's CUDA kernels, engines and host test
// implementation with caller-supplied (seeded) tensors instead of std::random_device ones.
//
// Every entry point copies caller buffers into reference `Tensor`s (they own their memory,
// include/tensor.hpp:97-120), calls the reference function, and copies results back.  Page tables
// hold raw device pointers into the CALLER's page pool, so the reference kernels read and write the
// caller's pages directly.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <list>
#include <string>
#include <vector>

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include "constants.h"
#include "inference_model.h"
#include "inferencer.h"
#include "item_storage.h"
#include "kernels/decoder.h"
#include "kernels/encoder.h"
#include "kernels/gemm.h"
#include "kernels/paged_attention.h"
#include "kernels/self_attention_inference_optimized.h"
#include "layers.h"
#include "paged_item_storage.h"
#include "tensor.hpp"
// reference host test implementation (tests/test_utils.cpp, tests/include/*.h)
#include "self_attention_inference_optimized_host.h"
#include "test_utils.h"

namespace {

thread_local std::string g_err;

template <typename T>
Tensor<T> dev_from(const T* src, std::vector<size_t> shape) {
    Tensor<T> t(shape, DeviceType::DEVICE);
    if (src) cudaMemcpy(t.data(), src, t.get_total_size() * sizeof(T), cudaMemcpyDefault);
    return t;
}
template <typename T>
Tensor<T> host_from(const T* src, std::vector<size_t> shape) {
    Tensor<T> t(shape, DeviceType::HOST);
    if (src) cudaMemcpy(t.data(), src, t.get_total_size() * sizeof(T), cudaMemcpyDefault);
    return t;
}
template <typename T>
void copy_out(T* dst, const Tensor<T>& t) {
    if (dst) cudaMemcpy(dst, t.data(), t.get_total_size() * sizeof(T), cudaMemcpyDefault);
}

template <typename F>
int guarded(F&& f) {
    try {
        f();
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            g_err = cudaGetErrorString(e);
            return -1;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

cublasHandle_t& handle() {
    static cublasHandle_t h = [] {
        cublasHandle_t x;
        cublasCreate(&x);
        return x;
    }();
    return h;
}

void fill_finished(const std::list<IdTokensPair>& fin, int* ids, int* offs, int* toks, int* n_fin) {
    int o = 0, i = 0;
    for (const auto& p : fin) {
        ids[i] = p.first;
        offs[i] = o;
        for (int t : p.second) toks[o++] = t;
        ++i;
    }
    offs[i] = o;
    *n_fin = i;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// ---- stage level ------------------------------------------------------------------------
// src/kernels/encoder.cu:134-147
int ref_paged_encoder(const float* emb, const float* pos, const int* inp, float** page_table,
                      const int* lengths, const int* new_idx, int B, int S, int d, int n_new) {
    return guarded([&] {
        launch_paged_attention_encoder_kernel(emb, pos, inp, page_table, lengths, new_idx, B, S, d,
                                              n_new);
    });
}

// variant 0: src/kernels/paged_attention.cu:96-115; 1: paged_attention_cublas.cu:225-246
int ref_prefill_kv_paged(float** page_table, const int* new_idx, const int* lengths,
                         const float* wk, const float* wv, int n_new, int B, int S, int d,
                         int variant) {
    return guarded([&] {
        auto pt = dev_from<float*>(page_table, {(size_t)B, (size_t)S / PAGE_BLOCK_SIZE});
        auto ni = dev_from<int>(new_idx, {(size_t)B});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto k = dev_from<float>(wk, {(size_t)d, (size_t)d});
        auto v = dev_from<float>(wv, {(size_t)d, (size_t)d});
        if (variant == 0)
            launch_fill_new_k_v_cache_paged_attention(pt, ni, ln, k, v, n_new, S);
        else
            launch_fill_new_k_v_cache_paged_attention_warp_tiling(pt, ni, ln, k, v, n_new, S);
    });
}

// variant 0: paged_attention.cu:188-199; 1: paged_attention_cublas.cu:76-99
int ref_qkv_latest_paged(float** page_table, const int* lengths, const float* wk,
                         const float* wq, const float* wv, float* q_output, int B, int S, int d,
                         int variant) {
    return guarded([&] {
        auto pt = dev_from<float*>(page_table, {(size_t)B, (size_t)S / PAGE_BLOCK_SIZE});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto k = dev_from<float>(wk, {(size_t)d, (size_t)d});
        auto q = dev_from<float>(wq, {(size_t)d, (size_t)d});
        auto v = dev_from<float>(wv, {(size_t)d, (size_t)d});
        auto qo = dev_from<float>(q_output, {(size_t)B, (size_t)d});
        if (variant == 0) {
            launch_get_latest_k_q_v_paged_attention(pt, ln, k, q, v, qo, S);
        } else {
            TensorFloat latest({(size_t)B, (size_t)d}, DeviceType::DEVICE);
            TensorFloat tmp({(size_t)B, (size_t)d}, DeviceType::DEVICE);
            launch_get_latest_k_q_v_paged_attention_cublas(pt, ln, latest, k, q, v, qo, tmp,
                                                           handle(), S);
        }
        cudaDeviceSynchronize();
        copy_out(q_output, qo);
    });
}

// paged_attention.cu:270-280
int ref_qkt_paged(const float* q, float** page_table, const int* lengths, float* qkt, int B,
                  int S, int d) {
    return guarded([&] {
        auto pt = dev_from<float*>(page_table, {(size_t)B, (size_t)S / PAGE_BLOCK_SIZE});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto qo = dev_from<float>(q, {(size_t)B, (size_t)d});
        auto out = dev_from<float>(qkt, {(size_t)B, (size_t)S});
        launch_qkt_paged_attention(qo, pt, ln, out);
        cudaDeviceSynchronize();
        copy_out(qkt, out);
    });
}

// self_attention_inference_optimized.cu:360-368
int ref_softmax_in_place_with_lengths(float* qkt, const int* lengths, int B, int S) {
    return guarded([&] {
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto out = dev_from<float>(qkt, {(size_t)B, (size_t)S});
        launch_softmax_in_place_with_lengths(out, ln);
        cudaDeviceSynchronize();
        copy_out(qkt, out);
    });
}

// paged_attention.cu:333-345
int ref_softmax_v_paged(const float* p, float** page_table, float* attention_result,
                        const int* lengths, int B, int S, int d) {
    return guarded([&] {
        auto pt = dev_from<float*>(page_table, {(size_t)B, (size_t)S / PAGE_BLOCK_SIZE});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto pp = dev_from<float>(p, {(size_t)B, (size_t)S});
        auto out = dev_from<float>(attention_result, {(size_t)B, (size_t)d});
        launch_softmax_v_paged_attention(pp, pt, out, ln);
        cudaDeviceSynchronize();
        copy_out(attention_result, out);
    });
}

// variant 0: paged_attention.cu:358-377; 1: paged_attention_cublas.cu:260-280
int ref_paged_attention(float** page_table, const int* lengths, const float* wk, const float* wq,
                        const float* wv, const int* new_idx, float* q_output, float* qkt_output,
                        float* attention_result, int n_new, int B, int S, int d, int variant) {
    return guarded([&] {
        auto pt = dev_from<float*>(page_table, {(size_t)B, (size_t)S / PAGE_BLOCK_SIZE});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto ni = dev_from<int>(new_idx, {(size_t)B});
        auto k = dev_from<float>(wk, {(size_t)d, (size_t)d});
        auto q = dev_from<float>(wq, {(size_t)d, (size_t)d});
        auto v = dev_from<float>(wv, {(size_t)d, (size_t)d});
        auto qo = dev_from<float>(q_output, {(size_t)B, (size_t)d});
        auto qkt = dev_from<float>(qkt_output, {(size_t)B, (size_t)S});
        auto out = dev_from<float>(attention_result, {(size_t)B, (size_t)d});
        if (variant == 0) {
            paged_attention(pt, ln, k, q, v, ni, qo, qkt, out, n_new, S);
        } else {
            TensorFloat latest({(size_t)B, (size_t)d}, DeviceType::DEVICE);
            TensorFloat tmp({(size_t)B, (size_t)d}, DeviceType::DEVICE);
            paged_attention_with_cublas(pt, ln, k, q, v, ni, qo, qkt, out, latest, tmp, n_new, S,
                                        handle());
        }
        cudaDeviceSynchronize();
        copy_out(q_output, qo);
        copy_out(qkt_output, qkt);
        copy_out(attention_result, out);
    });
}

// variant 0: decoder.cu:207-229 (gemm_transpose + decoder kernel); 1: decoder.cu:232-255 (cuBLAS)
int ref_paged_decoder(const float* attn, const float* emb, float* emb_score, const float* pos,
                      float** page_table, int* lengths, int* decoder_result, int B, int V, int S,
                      int d, int n_dec, int i_dec, int variant) {
    return guarded([&] {
        auto pt = dev_from<float*>(page_table, {(size_t)B, (size_t)S / PAGE_BLOCK_SIZE});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto a = dev_from<float>(attn, {(size_t)B, (size_t)d});
        auto e = dev_from<float>(emb, {(size_t)V, (size_t)d});
        auto p = dev_from<float>(pos, {(size_t)S, (size_t)d});
        auto sc = dev_from<float>(nullptr, {(size_t)B, (size_t)V});
        auto dr = dev_from<int>(decoder_result, {(size_t)B, (size_t)n_dec});
        if (variant == 0)
            launch_paged_attention_decoder_multi_rounds(a, e, sc, p, pt, ln, dr, i_dec);
        else
            launch_paged_attention_cublas_decoder_multi_rounds(a, e, sc, p, pt, ln, dr, i_dec,
                                                               handle());
        cudaDeviceSynchronize();
        copy_out(emb_score, sc);
        copy_out(lengths, ln);
        copy_out(decoder_result, dr);
    });
}

// src/inference_model.cpp:52-82 / :94-124 through the reference's own model classes
int ref_paged_forward(const int* inp, int* lengths, const int* new_idx, int* decoder_result,
                      int n_new, const float* emb, const float* pos, float** page_table,
                      const float* wk, const float* wq, const float* wv, int B, int S, int d,
                      int V, int R, int variant) {
    return guarded([&] {
        auto pt = dev_from<float*>(page_table, {(size_t)B, (size_t)S / PAGE_BLOCK_SIZE});
        auto in = dev_from<int>(inp, {(size_t)B, (size_t)S});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto ni = dev_from<int>(new_idx, {(size_t)B});
        auto e = dev_from<float>(emb, {(size_t)V, (size_t)d});
        auto p = dev_from<float>(pos, {(size_t)S, (size_t)d});
        auto dr = dev_from<int>(nullptr, {(size_t)B, (size_t)R});
        if (variant == 0) {
            PagedAttentionInferenceModel model(
                PagedAttentionLayer(dev_from<float>(wk, {(size_t)d, (size_t)d}),
                                    dev_from<float>(wq, {(size_t)d, (size_t)d}),
                                    dev_from<float>(wv, {(size_t)d, (size_t)d}), B, d, S),
                PagedEncoderLayer(), PagedDecoderLayer(B, V), B, S, d, R);
            model.forward(in, ln, ni, dr, n_new, e, p, pt);
        } else {
            PagedAttentionCublasInferenceModel model(
                PagedAttentionCublasLayer(dev_from<float>(wk, {(size_t)d, (size_t)d}),
                                          dev_from<float>(wq, {(size_t)d, (size_t)d}),
                                          dev_from<float>(wv, {(size_t)d, (size_t)d}), B, d, S),
                PagedEncoderLayer(), PagedCublasDecoderLayer(B, V), B, S, d, R);
            model.forward(in, ln, ni, dr, n_new, e, p, pt, handle());
        }
        cudaDeviceSynchronize();
        copy_out(lengths, ln);
        copy_out(decoder_result, dr);
    });
}

// ---- dense (non-paged) stage level ------------------------------------------------------
// encoder.cu:80-92
int ref_dense_encoder(const float* emb, const float* pos, const int* inp, float* inp_embedding,
                      const int* lengths, const int* new_idx, int B, int S, int d, int n_new) {
    return guarded([&] {
        launch_inference_optimized_encoder_kernel(emb, pos, inp, inp_embedding, lengths, new_idx,
                                                  B, S, d, n_new);
    });
}

// self_attention_inference_optimized.cu:282-301
int ref_self_attention(const float* inp_embedding, const int* lengths, const float* wk,
                       const float* wq, const float* wv, const int* new_idx, float* kt_cache,
                       float* v_cache, float* q_output, float* qkt_output,
                       float* attention_result, int n_new, int B, int S, int di, int dn) {
    return guarded([&] {
        auto x = dev_from<float>(inp_embedding, {(size_t)B, (size_t)S, (size_t)di});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto ni = dev_from<int>(new_idx, {(size_t)B});
        auto k = dev_from<float>(wk, {(size_t)di, (size_t)dn});
        auto q = dev_from<float>(wq, {(size_t)di, (size_t)dn});
        auto v = dev_from<float>(wv, {(size_t)di, (size_t)dn});
        auto kt = dev_from<float>(kt_cache, {(size_t)B, (size_t)dn, (size_t)S});
        auto vc = dev_from<float>(v_cache, {(size_t)B, (size_t)S, (size_t)dn});
        auto qo = dev_from<float>(q_output, {(size_t)B, (size_t)dn});
        auto qkt = dev_from<float>(qkt_output, {(size_t)B, (size_t)S});
        auto out = dev_from<float>(attention_result, {(size_t)B, (size_t)dn});
        inference_self_attention(x, ln, k, q, v, ni, kt, vc, qo, qkt, out, n_new);
        cudaDeviceSynchronize();
        copy_out(kt_cache, kt);
        copy_out(v_cache, vc);
        copy_out(q_output, qo);
        copy_out(qkt_output, qkt);
        copy_out(attention_result, out);
    });
}

// decoder.cu:94-112
int ref_dense_decoder(const float* attn, const float* emb, float* emb_score, const float* pos,
                      float* inp_embedding, int* lengths, int* decoder_result, int B, int V, int S,
                      int d) {
    return guarded([&] {
        auto a = dev_from<float>(attn, {(size_t)B, (size_t)d});
        auto e = dev_from<float>(emb, {(size_t)V, (size_t)d});
        auto p = dev_from<float>(pos, {(size_t)S, (size_t)d});
        auto sc = dev_from<float>(nullptr, {(size_t)B, (size_t)V});
        auto x = dev_from<float>(inp_embedding, {(size_t)B, (size_t)S, (size_t)d});
        auto ln = dev_from<int>(lengths, {(size_t)B});
        auto dr = dev_from<int>(nullptr, {(size_t)B});
        launch_decoder(a, e, sc, p, x, ln, dr);
        cudaDeviceSynchronize();
        copy_out(emb_score, sc);
        copy_out(inp_embedding, x);
        copy_out(lengths, ln);
        copy_out(decoder_result, dr);
    });
}

// ---- engines ----------------------------------------------------------------------------
// variant 0: start_paged_attention_inference_engine (src/inferencer.cpp:43-85)
// variant 1: start_paged_attention_cublas_inference_engine (:87-133)
// weights/tables are HOST pointers here; seconds = wall time of the engine call (the reference's
// own metric, ThroughputCounter, also counts host work: src/throughput_counter.cpp:8-30).
int ref_run_paged_engine(int variant, int B, int S, int d, int V, int n_blocks, int R,
                         const float* emb, const float* pos, const float* wk, const float* wq,
                         const float* wv, int n_req, const int* prompt_offsets,
                         const int* prompt_tokens, int* finished_ids, int* finished_offsets,
                         int* finished_tokens, int* n_finished, double* seconds) {
    return guarded([&] {
        PagedAttentionsManager pam(B, S, d);
        MemoryBlockManager mbm(n_blocks, (size_t)PAGE_BLOCK_SIZE * 3 * d);
        ProcessingStorage ps;
        ItemStorage is;
        for (int i = 0; i < n_req; ++i)
            is.add_new_item(IdTokensPair(
                i, std::vector<int>(prompt_tokens + prompt_offsets[i],
                                    prompt_tokens + prompt_offsets[i + 1])));
        auto e = dev_from<float>(emb, {(size_t)V, (size_t)d});
        auto p = dev_from<float>(pos, {(size_t)S, (size_t)d});
        cudaDeviceSynchronize();
        auto t0 = std::chrono::high_resolution_clock::now();
        if (variant == 0) {
            PagedAttentionInferenceModel model(
                PagedAttentionLayer(dev_from<float>(wk, {(size_t)d, (size_t)d}),
                                    dev_from<float>(wq, {(size_t)d, (size_t)d}),
                                    dev_from<float>(wv, {(size_t)d, (size_t)d}), B, d, S),
                PagedEncoderLayer(), PagedDecoderLayer(B, V), B, S, d, R);
            cudaDeviceSynchronize();
            t0 = std::chrono::high_resolution_clock::now();
            start_paged_attention_inference_engine(e, p, is, ps, mbm, pam, model, B, S, R);
        } else {
            PagedAttentionCublasInferenceModel model(
                PagedAttentionCublasLayer(dev_from<float>(wk, {(size_t)d, (size_t)d}),
                                          dev_from<float>(wq, {(size_t)d, (size_t)d}),
                                          dev_from<float>(wv, {(size_t)d, (size_t)d}), B, d, S),
                PagedEncoderLayer(), PagedCublasDecoderLayer(B, V), B, S, d, R);
            cudaDeviceSynchronize();
            t0 = std::chrono::high_resolution_clock::now();
            start_paged_attention_cublas_inference_engine(e, p, is, ps, mbm, pam, model, B, S, R);
        }
        cudaDeviceSynchronize();
        auto t1 = std::chrono::high_resolution_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        fill_finished(is.get_finished_items(), finished_ids, finished_offsets, finished_tokens,
                      n_finished);
    });
}

// start_inference_engine (src/inferencer.cpp:11-41)
int ref_run_dense_engine(int B, int S, int d, int V, const float* emb, const float* pos,
                         const float* wk, const float* wq, const float* wv, int n_req,
                         const int* prompt_offsets, const int* prompt_tokens, int* finished_ids,
                         int* finished_offsets, int* finished_tokens, int* n_finished,
                         double* seconds) {
    return guarded([&] {
        ProcessingStorage ps;
        ItemStorage is;
        for (int i = 0; i < n_req; ++i)
            is.add_new_item(IdTokensPair(
                i, std::vector<int>(prompt_tokens + prompt_offsets[i],
                                    prompt_tokens + prompt_offsets[i + 1])));
        auto e = dev_from<float>(emb, {(size_t)V, (size_t)d});
        auto p = dev_from<float>(pos, {(size_t)S, (size_t)d});
        InferenceModel model(
            SelfAttentionLayer(dev_from<float>(wk, {(size_t)d, (size_t)d}),
                               dev_from<float>(wq, {(size_t)d, (size_t)d}),
                               dev_from<float>(wv, {(size_t)d, (size_t)d}), B, d, S),
            EncoderLayer(), DecoderLayer(B, V), B, S, d);
        cudaDeviceSynchronize();
        auto t0 = std::chrono::high_resolution_clock::now();
        start_inference_engine(e, p, is, ps, model, B, S);
        cudaDeviceSynchronize();
        auto t1 = std::chrono::high_resolution_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        fill_finished(is.get_finished_items(), finished_ids, finished_offsets, finished_tokens,
                      n_finished);
    });
}

// The reference's HOST implementation of the path: one forward =
// inference_optimized_encoder_host (tests/test_utils.cpp:559-573) + self_attention_inference_host
// (:502-519) + decoder_host (:633-647), driven by the reference's own non-paged
// insert_new_items / process_decoder_result (src/item_storage.cpp:97-180).  Single-threaded as
// written.  max_steps bounds the sample (<=0: run to completion).  Returns generated tokens and
// the wall time of the loop.
int ref_run_host_engine(int B, int S, int d, int V, const float* emb, const float* pos,
                        const float* wk, const float* wq, const float* wv, int n_req,
                        const int* prompt_offsets, const int* prompt_tokens, int max_steps,
                        long long* generated_tokens, long long* steps_done, double* seconds,
                        int* finished_ids, int* finished_offsets, int* finished_tokens,
                        int* n_finished) {
    return guarded([&] {
        ProcessingStorage ps;
        ItemStorage is;
        for (int i = 0; i < n_req; ++i)
            is.add_new_item(IdTokensPair(
                i, std::vector<int>(prompt_tokens + prompt_offsets[i],
                                    prompt_tokens + prompt_offsets[i + 1])));
        size_t b = B, s = S, dd = d, v = V;
        auto e = host_from<float>(emb, {v, dd});
        auto p = host_from<float>(pos, {s, dd});
        auto k = host_from<float>(wk, {dd, dd});
        auto q = host_from<float>(wq, {dd, dd});
        auto vv = host_from<float>(wv, {dd, dd});
        TensorFloat x({b, s, dd}, DeviceType::HOST), kt({b, dd, s}, DeviceType::HOST),
            vc({b, s, dd}, DeviceType::HOST), qo({b, dd}, DeviceType::HOST),
            qkt({b, s}, DeviceType::HOST), attn({b, dd}, DeviceType::HOST),
            score({b, v}, DeviceType::HOST);
        TensorInt inp_d({b, s}, DeviceType::DEVICE), inp_h({b, s}, DeviceType::HOST),
            len_d({b}, DeviceType::DEVICE), len_h({b}, DeviceType::HOST),
            idx_d({b}, DeviceType::DEVICE), idx_h({b}, DeviceType::HOST),
            dec_d({b}, DeviceType::DEVICE), dec_h({b}, DeviceType::HOST);
        TensorInt len_w({b}, DeviceType::HOST), dec_w({b}, DeviceType::HOST);
        cudaMemset(len_d.data(), 0, b * sizeof(int));
        std::vector<int> finished;
        for (int i = 0; i < B; ++i) finished.push_back(i);
        long long gen = 0, steps = 0;
        auto t0 = std::chrono::high_resolution_clock::now();
        int n_new = insert_new_items(finished, inp_d, inp_h, len_d, len_h, idx_d, idx_h, is, ps);
        while (!is_done(is, ps)) {
            if (max_steps > 0 && steps >= max_steps) break;
            // host forward on host tensors (inp_h / idx_h were filled by insert_new_items)
            len_w.copy_from(len_d);
            inference_optimized_encoder_host(e.data(), p.data(), inp_h.data(), x.data(),
                                             len_w.data(), idx_h.data(), B, S, d, n_new);
            self_attention_inference_host(x, len_w, k, q, vv, idx_h, kt, vc, qo, qkt, attn, n_new);
            decoder_host(attn, e, score, p, x, len_w, dec_w);
            len_d.copy_from(len_w);
            dec_d.copy_from(dec_w);
            int before = 0;
            for (int i = 0; i < B; ++i) before += (dec_w.data()[i] != EMPTY_ROW_TOKEN_ID);
            gen += before;
            finished = process_decoder_result(dec_d, dec_h, is, ps, S);
            n_new = insert_new_items(finished, inp_d, inp_h, len_d, len_h, idx_d, idx_h, is, ps);
            ++steps;
        }
        auto t1 = std::chrono::high_resolution_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        if (generated_tokens) *generated_tokens = gen;
        if (steps_done) *steps_done = steps;
        if (finished_ids)
            fill_finished(is.get_finished_items(), finished_ids, finished_offsets, finished_tokens,
                          n_finished);
    });
}

}  // extern "C"
