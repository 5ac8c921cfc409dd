// mli/compat.hpp -- C++17 mirror of the reference's public interface for the paged-attention decode
// path, implemented on top of the C ABI (include/mli_b200.h).  Source compatible with code written
// against xyg-coder/min_llm_inference: same global names, argument order, ownership and error
// behaviour, so a driver that includes "inferencer.h" / "inference_model.h" / "tensor.hpp" compiles
// unchanged against either tree (tests/dropin/ builds one driver source both ways).
//
// What sits behind each piece (reference file:line):
//   Tensor<T>, DeviceType, TensorDataType            include/tensor.hpp:11-120, :324-326
//   constants, ceil_div, cuda_check, error macros     include/constants.h:3-18, include/utils.h:5-103
//   Storage / ItemStorage / ProcessingStorage         include/item_storage.h:10-90
//   MemoryBlockManager / PagedAttentionsManager       include/paged_item_storage.h:10-57
//   layers and models                                 include/layers.h:54-154, include/inference_model.h:8-74
//   start_*_inference_engine                          include/inferencer.h:18-32
//   kernel-level launchers                            include/kernels/{paged_attention,decoder,encoder,
//                                                     self_attention_inference_optimized}.h
//   ThroughputCounter                                 include/throughput_counter.h:5-18
//
// Deliberate differences, all documented in INTEGRATION.md:
//   * paged_attention() runs ONE fused kernel instead of the three unfused launches (qkt / softmax /
//     softmax_v); it keeps its signature and fills qkt_output with the softmax probabilities exactly
//     as the reference leaves it.  The three launchers still exist (same signatures, the reference's
//     summation order) for code and tests written against them; launch_fused_decode_attention() is the
//     stand-alone fused call;
//   * the cuBLAS handle arguments are accepted and ignored (there is no cuBLAS in this build);
//   * the paged engines run on the device scheduler; by default they replay the reference's
//     stale-length behaviour so finished token lists are identical -- mli::set_fix_stale_lengths(true)
//     (or MLI_FIX_STALE_LENGTHS=1) selects the corrected behaviour.
#pragma once

#include <cuda_runtime.h>

#include <chrono>
#include <cstddef>
#include <cstdio>
#include <list>
#include <memory>
#include <optional>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>

#include "mli_b200.h"

// cuBLAS is not used; the type only has to exist for the reference's "cublas" signatures
#if __has_include(<cublas_v2.h>)
#include <cublas_v2.h>
#else
struct cublasContext;
typedef struct cublasContext* cublasHandle_t;
#endif

// ---- constants (include/constants.h) -------------------------------------------------------------
constexpr int TILE_SIZE = 16;
constexpr int WARP_SIZE = 32;
constexpr int TILE_SIZE_SQUARE = TILE_SIZE * TILE_SIZE;
constexpr int BLOCK_DIM = 256;
constexpr int EMPTY_ROW_TOKEN_ID = MLI_EMPTY_ROW_TOKEN_ID;
constexpr int EOF_TOKEN_ID = MLI_EOF_TOKEN_ID;
constexpr int PAGE_BLOCK_SIZE = MLI_PAGE_BLOCK_SIZE;
constexpr int DEFAULT_INIT_NUM_BLOCKS = MLI_DEFAULT_INIT_NUM_BLOCKS;
constexpr int INP_EMB_EMB_OFFSET = 0;
constexpr int K_CACHE_EMB_OFFSET = 1;
constexpr int V_CACHE_EMB_OFFSET = 2;

// ---- utils (include/utils.h) ------------------------------------------------------------------------
void cuda_check(cudaError_t error, const char* file, int line);  // printf + throw "Cuda Failure"
#define CUDA_CHECK_LAST() cuda_check(cudaGetLastError(), __FILE__, __LINE__)
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

class NonCopyableNonClonable {
protected:
    NonCopyableNonClonable() = default;
    ~NonCopyableNonClonable() = default;
    NonCopyableNonClonable(NonCopyableNonClonable&&) noexcept = default;
    NonCopyableNonClonable& operator=(NonCopyableNonClonable&&) noexcept = default;

public:
    NonCopyableNonClonable(const NonCopyableNonClonable&) = delete;
    NonCopyableNonClonable& operator=(const NonCopyableNonClonable&) = delete;
};

namespace mli {
// process-wide C-ABI context on the legacy default stream (the reference launches everything there)
mli_ctx* host_context();
// throws std::runtime_error("Cuda Failure") / the library's message when a C-ABI call fails
void check(int status);
void set_fix_stale_lengths(bool on);
bool fix_stale_lengths();
// request-sharded multi-GPU run of the paged engines inside one process (MLI_NUM_GPUS=n or
// set_num_gpus(n)): GPU g takes the g-th contiguous block of queued requests; the finished token
// lists come back through mli_comm_gather_tokens (NCCL all-gather)
void set_num_gpus(int n);
int num_gpus();
}  // namespace mli

// ---- Tensor (include/tensor.hpp) -----------------------------------------------------------------------
enum class DeviceType { HOST, DEVICE };
enum class TensorDataType { SYNC_ALLOCATE = 0, ASYNC_ALLOCATE = 1 };
#ifndef DEFAULT_ALLOC_METHOD
#define DEFAULT_ALLOC_METHOD 0
#endif
constexpr TensorDataType DEFAULT_TENSOR_DATA_TYPE = static_cast<TensorDataType>(DEFAULT_ALLOC_METHOD);

namespace mli {
// one allocation: device memory (cudaMalloc) or pinned host memory (cudaHostAlloc), freed with the
// last Tensor that aliases it.  Both allocation policies of the reference map onto it.
struct Buffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    DeviceType device = DeviceType::HOST;
    Buffer(size_t bytes, DeviceType device);
    ~Buffer();
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;
};
void copy_buffer(Buffer& dst, const Buffer& src);  // blocking copy, size must match
}  // namespace mli

template <typename T>
class Tensor {
public:
    Tensor(const std::vector<size_t>& shape, DeviceType device = DeviceType::HOST,
           TensorDataType = DEFAULT_TENSOR_DATA_TYPE)
        : shape_(shape), size_(1), device_(device) {
        for (size_t d : shape_) size_ *= d;
        buf_ = std::make_shared<mli::Buffer>(size_ * sizeof(T), device);
    }
    Tensor(const Tensor&) = default;  // aliases the same memory, like the reference
    Tensor& operator=(const Tensor&) = default;

    const std::vector<size_t>& shape() const { return shape_; }
    DeviceType device() const { return device_; }
    T* data() { return static_cast<T*>(buf_->ptr); }
    const T* data() const { return static_cast<const T*>(buf_->ptr); }
    size_t get_total_size() const { return size_; }
    void copy_from(const Tensor& other) {
        if (other.size_ != size_) throw std::runtime_error("Copy from: shape or device mismatch");
        mli::copy_buffer(*buf_, *other.buf_);
    }

private:
    std::vector<size_t> shape_;
    size_t size_;
    DeviceType device_;
    std::shared_ptr<mli::Buffer> buf_;
};

typedef Tensor<float> TensorFloat;
typedef Tensor<int> TensorInt;
typedef Tensor<float*> TensorFloatPoint;

// ---- request queues (include/item_storage.h) ----------------------------------------------------------
using IdTokensPair = std::pair<int, std::vector<int>>;

class Storage : public NonCopyableNonClonable {
public:
    Storage() = default;
    std::vector<IdTokensPair> pop_pairs(int size);
    void add(IdTokensPair&&);
    void add_to_front(IdTokensPair&&);
    int size() const;
    int head_length() const;
    const IdTokensPair& get_top() const;
    const std::list<IdTokensPair>& get_data() const;

private:
    std::list<IdTokensPair> data_;
};

class ItemStorage : public NonCopyableNonClonable {
public:
    ItemStorage() = default;
    std::vector<IdTokensPair> pop_finished_items(int size);
    std::vector<IdTokensPair> pop_new_items(int size);
    const IdTokensPair& get_top() const;
    void add_finished_item(IdTokensPair&&);
    void add_new_item(IdTokensPair&&);
    void add_new_item_to_head(IdTokensPair&&);
    int finish_count() const;
    int new_count() const;
    int head_length() const;
    const std::list<IdTokensPair>& get_finished_items() const;

private:
    Storage finished_items_;
    Storage new_items_;
};

class ProcessingStorage : public NonCopyableNonClonable {
public:
    ProcessingStorage() = default;
    void put(int batch_id, IdTokensPair&&);
    void remove(int batch_id);
    bool batch_id_processing(int batch_id);
    IdTokensPair& get_token(int batch_id);
    void move_to_new(int batch_id, ItemStorage& item_storage);
    int size() const;
    void move_to_finished(int batch_id, ItemStorage& item_storage);

private:
    std::unordered_map<int, IdTokensPair> batch_id_to_token_pairs_;
};

void append_token_to_id_string_pair(IdTokensPair& id_string_pair, int to_add);

std::vector<int> process_decoder_result(const TensorInt& decoder_result_device,
                                        TensorInt& decoder_result_host, ItemStorage& item_storage,
                                        ProcessingStorage& processing_storage, int n_sequence);

int insert_new_items(const std::vector<int>& finished_indices, TensorInt& inp_device,
                     TensorInt& inp_host, TensorInt& lengths_device, TensorInt& lengths_host,
                     TensorInt& new_items_indices_device, TensorInt& new_items_indices_host,
                     ItemStorage& item_storage, ProcessingStorage& processing_storage);

bool is_done(ItemStorage& item_storage, ProcessingStorage& processing_storage);

// ---- KV page manager (include/paged_item_storage.h) ---------------------------------------------------
class MemoryBlockManager {
public:
    MemoryBlockManager(int n_blocks, size_t each_block_size);
    int free_blocks_size() const;
    std::list<float*> pop_free_blocks(int size);  // throws "No enough block memories to return"
    void return_free_blocks(std::list<float*>&&);
    // extensions used by the device engine: the slab the blocks were carved from
    float* slab() { return block_memory_.data(); }
    int total_blocks() const { return n_blocks_; }
    size_t block_size() const { return each_block_size_; }

private:
    TensorFloat block_memory_;
    std::list<float*> free_blocks_;
    int n_blocks_;
    size_t each_block_size_;
};

using BatchIdMemoryBlocksPair = std::pair<int, std::list<float*>>;

class PagedAttentionsManager {
public:
    PagedAttentionsManager(size_t max_batches, size_t n_sequence, size_t emb_dim);
    std::list<BatchIdMemoryBlocksPair>& get_used_block_list();
    void maybe_flush_changes();
    void add_batch_block_pair(BatchIdMemoryBlocksPair&&);
    void set_block_pos(int batch_id, int i_block, float*);
    TensorFloatPoint& get_page_table_device();

private:
    TensorFloatPoint page_table_host;
    TensorFloatPoint page_table_device;
    std::list<BatchIdMemoryBlocksPair> used_blocks_;
    size_t width_;
    bool needs_sync_;
};

void allocate_memory_block(MemoryBlockManager&, PagedAttentionsManager&, BatchIdMemoryBlocksPair&);

void allocate_or_free_memory_blocks_if_needed(PagedAttentionsManager&, MemoryBlockManager&,
                                              ProcessingStorage&, ItemStorage&,
                                              const std::vector<int>& finished_indices,
                                              int n_forward_rounds);

std::vector<int> insert_new_items(TensorInt& inp_device, TensorInt& inp_host,
                                  TensorInt& lengths_device, TensorInt& lengths_host,
                                  TensorInt& new_items_indices_device,
                                  TensorInt& new_items_indices_host, ItemStorage& item_storage,
                                  ProcessingStorage& processing_storage,
                                  MemoryBlockManager& memory_block_manager,
                                  PagedAttentionsManager& paged_attention_manager,
                                  int n_forward_rounds);

// ---- kernel-level launchers (include/kernels/*.h) -----------------------------------------------------
void paged_attention(TensorFloatPoint& page_table, const TensorInt& lengths, const TensorFloat& wk,
                     const TensorFloat& wq, const TensorFloat& wv, const TensorInt& new_batch_idx,
                     TensorFloat& q_output, TensorFloat& qkt_output, TensorFloat& attention_result,
                     int n_new_items, int n_sequence);

void launch_fill_new_k_v_cache_paged_attention(TensorFloatPoint page_table,
                                               const TensorInt& new_batch_idx,
                                               const TensorInt& lengths, const TensorFloat& wk,
                                               const TensorFloat& wv, int n_new_items, int n_sequence);

void launch_get_latest_k_q_v_paged_attention(TensorFloatPoint& page_table, const TensorInt& lengths,
                                             const TensorFloat& wk, const TensorFloat& wq,
                                             const TensorFloat& wv, TensorFloat& q_output,
                                             int n_sequence);

// include/kernels/paged_attention.h:37-43, include/kernels/self_attention_inference_optimized.h:21-22
void launch_qkt_paged_attention(const TensorFloat& q_output, const TensorFloatPoint& page_table,
                                const TensorInt& lengths, TensorFloat& qkt_output);
void launch_softmax_in_place_with_lengths(TensorFloat& qkt_output, const TensorInt& lengths);
void launch_softmax_v_paged_attention(const TensorFloat& softmax_result, const TensorFloatPoint& page_table,
                                      TensorFloat& attention_result, const TensorInt& lengths);

// fused replacement of launch_qkt_paged_attention + launch_softmax_in_place_with_lengths +
// launch_softmax_v_paged_attention; softmax_result (may be nullptr) receives the probabilities
void launch_fused_decode_attention(const TensorFloat& q_output, const TensorFloatPoint& page_table,
                                   const TensorInt& lengths, TensorFloat& attention_result,
                                   TensorFloat* softmax_result, int n_sequence);

void paged_attention_with_cublas(TensorFloatPoint& page_table, const TensorInt& lengths,
                                 const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv,
                                 const TensorInt& new_batch_idx, TensorFloat& q_output,
                                 TensorFloat& qkt_output, TensorFloat& attention_result,
                                 TensorFloat& latest_emb, TensorFloat& temp_placeholder,
                                 int n_new_items, int n_sequence, cublasHandle_t& handle);

void launch_get_latest_k_q_v_paged_attention_cublas(TensorFloatPoint& page_table,
                                                    const TensorInt& lengths, TensorFloat& latest_emb,
                                                    const TensorFloat& wk, const TensorFloat& wq,
                                                    const TensorFloat& wv, TensorFloat& q_output,
                                                    TensorFloat& temp_placeholder,
                                                    cublasHandle_t& handle, int n_sequence);

void launch_fill_new_k_v_cache_paged_attention_warp_tiling(TensorFloatPoint page_table,
                                                           const TensorInt& new_batch_idx,
                                                           const TensorInt& lengths,
                                                           const TensorFloat& wk, const TensorFloat& wv,
                                                           int n_new_items, int n_sequence);

void launch_decoder(const TensorFloat& batch_result, const TensorFloat& emb_table,
                    TensorFloat& emb_score, const TensorFloat& wpe_table, TensorFloat& inp_embedding,
                    TensorInt& lengths, TensorInt& decoder_result);

void launch_paged_attention_decoder_multi_rounds(const TensorFloat& batch_result,
                                                 const TensorFloat& emb_table, TensorFloat& emb_score,
                                                 const TensorFloat& wpe_table,
                                                 TensorFloatPoint& page_table, TensorInt& lengths,
                                                 TensorInt& decoder_result, int i_decoder);

void launch_paged_attention_cublas_decoder_multi_rounds(
    const TensorFloat& batch_result, const TensorFloat& emb_table, TensorFloat& emb_score,
    const TensorFloat& wpe_table, TensorFloatPoint& page_table, TensorInt& lengths,
    TensorInt& decoder_result, int i_decoder, cublasHandle_t& handle);

void launch_inference_optimized_encoder_kernel(const float* emb_table, const float* wpe,
                                               const int* inp, float* inp_embedding,
                                               const int* lengths, const int* new_item_indices,
                                               int batch_size, int n_sequence, int embedding_dim,
                                               int n_new_items);

void launch_paged_attention_encoder_kernel(const float* emb_table, const float* wpe, const int* inp,
                                           float** page_table, const int* lengths,
                                           const int* new_item_indices, int batch_size,
                                           int n_sequence, int embedding_dim, int n_new_items);

void inference_self_attention(const TensorFloat& inp_embedding, const TensorInt& lengths,
                              const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv,
                              const TensorInt& new_batch_idx, TensorFloat& kt_cache,
                              TensorFloat& v_cache, TensorFloat& q_output, TensorFloat& qkt_output,
                              TensorFloat& attention_result, int n_new_items);

// ---- layers (include/layers.h) ---------------------------------------------------------------------------
class SelfAttentionLayer : public NonCopyableNonClonable {
public:
    SelfAttentionLayer(TensorFloat&& wk, TensorFloat&& wq, TensorFloat&& wv, size_t n_batch,
                       size_t input_dim, size_t n_sequence);
    void forward(const TensorFloat& inp_embedding, const TensorInt& lengths,
                 const TensorInt& new_batch_idx, TensorFloat& attention_result, int n_new_items);
    // extension: the device engine reads the owned weights
    const TensorFloat& wk() const { return wk_; }
    const TensorFloat& wq() const { return wq_; }
    const TensorFloat& wv() const { return wv_; }

private:
    TensorFloat wk_, wq_, wv_;
    TensorFloat kt_cache_, v_cache_, q_output_, qkt_output_;
};

class PagedAttentionLayer : public NonCopyableNonClonable {
public:
    PagedAttentionLayer(TensorFloat&& wk, TensorFloat&& wq, TensorFloat&& wv, size_t n_batch,
                        size_t emb_dim, size_t n_sequence);
    void forward(TensorFloatPoint& page_table, const TensorInt& lengths,
                 const TensorInt& new_batch_idx, TensorFloat& attention_result, int n_new_items);
    // extension: the device engine and the tcgen05 weight registration read the owned weights
    const TensorFloat& wk() const { return wk_; }
    const TensorFloat& wq() const { return wq_; }
    const TensorFloat& wv() const { return wv_; }

private:
    TensorFloat wk_, wq_, wv_;
    TensorFloat q_output_, qkt_output_;
};

class PagedAttentionCublasLayer : public NonCopyableNonClonable {
public:
    PagedAttentionCublasLayer(TensorFloat&& wk, TensorFloat&& wq, TensorFloat&& wv, size_t n_batch,
                              size_t emb_dim, size_t n_sequence);
    void forward(TensorFloatPoint& page_table, const TensorInt& lengths,
                 const TensorInt& new_batch_idx, TensorFloat& attention_result, int n_new_items,
                 cublasHandle_t& handle);
    const TensorFloat& wk() const { return wk_; }
    const TensorFloat& wq() const { return wq_; }
    const TensorFloat& wv() const { return wv_; }

private:
    TensorFloat wk_, wq_, wv_;
    TensorFloat q_output_, qkt_output_, latest_emb_, temp_placeholder_;
};

class EncoderLayer : public NonCopyableNonClonable {
public:
    void forward(const TensorFloat& emb_table, const TensorFloat& pos_emb, const TensorInt& inp,
                 TensorFloat& inp_embedding, const TensorInt& lengths,
                 const TensorInt& new_item_indices, int n_new_items);
};

class PagedEncoderLayer : public NonCopyableNonClonable {
public:
    void forward(const TensorFloat& emb_table, const TensorFloat& pos_emb, const TensorInt& inp,
                 TensorFloatPoint& page_table, const TensorInt& lengths,
                 const TensorInt& new_item_indices, int n_new_items);
};

class DecoderLayer : public NonCopyableNonClonable {
public:
    DecoderLayer(size_t n_batch, size_t n_vocab);
    void forward(const TensorFloat& batch_result, const TensorFloat& emb_table,
                 const TensorFloat& wpe_table, TensorFloat& inp_embedding, TensorInt& lengths,
                 TensorInt& decoder_result);

private:
    TensorFloat emb_score_;
};

class PagedDecoderLayer : public NonCopyableNonClonable {
public:
    PagedDecoderLayer(size_t n_batch, size_t n_vocab);
    void forward(const TensorFloat& batch_result, const TensorFloat& emb_table,
                 const TensorFloat& wpe_table, TensorFloatPoint& page_table, TensorInt& lengths,
                 TensorInt& decoder_result, int i_decoder_round);

private:
    TensorFloat emb_score_;
};

class PagedCublasDecoderLayer : public NonCopyableNonClonable {
public:
    PagedCublasDecoderLayer(size_t n_batch, size_t n_vocab);
    void forward(const TensorFloat& batch_result, const TensorFloat& emb_table,
                 const TensorFloat& wpe_table, TensorFloatPoint& page_table, TensorInt& lengths,
                 TensorInt& decoder_result, int i_decoder_round, cublasHandle_t& handle);

private:
    TensorFloat emb_score_;
};

// ---- models (include/inference_model.h) ---------------------------------------------------------------
class InferenceModel : public NonCopyableNonClonable {
public:
    InferenceModel(SelfAttentionLayer&&, EncoderLayer&&, DecoderLayer&&, size_t n_batch,
                   size_t n_sequence, size_t emb_dim);
    void forward(const TensorInt& inp, TensorInt& lengths, const TensorInt& new_item_indices,
                 TensorInt& decoder_result, int n_new_items, const TensorFloat& emb_table,
                 const TensorFloat& pos_emb_table);
    const SelfAttentionLayer& attention_layer() const { return attention_layer_; }  // extension
    size_t emb_dim() const { return emb_dim_; }

private:
    SelfAttentionLayer attention_layer_;
    EncoderLayer encoder_layer_;
    DecoderLayer decoder_layer_;
    size_t n_batch_, n_sequence_, emb_dim_;
    TensorFloat inp_embedding_, attention_result_;
};

class PagedAttentionInferenceModel : public NonCopyableNonClonable {
public:
    PagedAttentionInferenceModel(PagedAttentionLayer&&, PagedEncoderLayer&&, PagedDecoderLayer&&,
                                 size_t n_batch, size_t n_sequence, size_t emb_dim,
                                 int n_forward_rounds);
    void forward(const TensorInt& inp, TensorInt& lengths, const TensorInt& new_item_indices,
                 TensorInt& decoder_result, int n_new_items, const TensorFloat& emb_table,
                 const TensorFloat& pos_emb_table, TensorFloatPoint& page_table);
    const PagedAttentionLayer& attention_layer() const { return paged_attention_layer_; }  // extension
    size_t emb_dim() const { return emb_dim_; }

private:
    PagedAttentionLayer paged_attention_layer_;
    PagedEncoderLayer paged_encoder_layer_;
    PagedDecoderLayer paged_decoder_layer_;
    size_t n_batch_, n_sequence_, emb_dim_;
    TensorFloat attention_result_;
    int n_forward_rounds_;
};

class PagedAttentionCublasInferenceModel : public NonCopyableNonClonable {
public:
    PagedAttentionCublasInferenceModel(PagedAttentionCublasLayer&&, PagedEncoderLayer&&,
                                       PagedCublasDecoderLayer&&, size_t n_batch, size_t n_sequence,
                                       size_t emb_dim, int n_forward_rounds);
    void forward(const TensorInt& inp, TensorInt& lengths, const TensorInt& new_item_indices,
                 TensorInt& decoder_result, int n_new_items, const TensorFloat& emb_table,
                 const TensorFloat& pos_emb_table, TensorFloatPoint& page_table,
                 cublasHandle_t handle);
    const PagedAttentionCublasLayer& attention_layer() const { return paged_attention_layer_; }
    size_t emb_dim() const { return emb_dim_; }

private:
    PagedAttentionCublasLayer paged_attention_layer_;
    PagedEncoderLayer paged_encoder_layer_;
    PagedCublasDecoderLayer paged_decoder_layer_;
    size_t n_batch_, n_sequence_, emb_dim_;
    TensorFloat attention_result_;
    int n_forward_rounds_;
};

// ---- throughput counter (include/throughput_counter.h) ---------------------------------------------------
class ThroughputCounter {
public:
    ThroughputCounter();
    void print_throughput();
    void start_record();
    void add_record_if_recording(int new_tokens);
    // extensions: the device engine reports one job at a time
    void add_job(long long tokens, double seconds);
    long long tokens() const { return total_tokens_; }
    double seconds() const { return seconds_; }

private:
    long long total_tokens_;
    double seconds_;
    std::chrono::time_point<std::chrono::high_resolution_clock> last_timestamp_;
    bool in_recording_;
};

ThroughputCounter& get_global_throughput_counter();

// ---- engine loops (include/inferencer.h) --------------------------------------------------------------------
void start_inference_engine(const TensorFloat& emb_table, const TensorFloat& pos_table,
                            ItemStorage& item_storage, ProcessingStorage& processing_storage,
                            InferenceModel& inference_model, size_t n_batch_size, size_t n_sequence);

void start_paged_attention_inference_engine(
    const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
    ProcessingStorage& processing_storage, MemoryBlockManager& memory_block_manager,
    PagedAttentionsManager& paged_attention_manager, PagedAttentionInferenceModel& inference_model,
    size_t n_batch_size, size_t n_sequence, int n_forward_rounds);

void start_paged_attention_cublas_inference_engine(
    const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
    ProcessingStorage& processing_storage, MemoryBlockManager& memory_block_manager,
    PagedAttentionsManager& paged_attention_manager,
    PagedAttentionCublasInferenceModel& inference_model, size_t n_batch_size, size_t n_sequence,
    int n_forward_rounds);
