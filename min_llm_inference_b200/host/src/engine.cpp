// engine.cpp -- the three engine loops of include/inferencer.h.
//
// start_inference_engine (non-paged, config C1) keeps the reference's host loop
// (src/inferencer.cpp:11-41) over InferenceModel::forward -- it is launch-bound and tiny.
//
// The two paged engines (src/inferencer.cpp:43-133) hand the whole job to the on-device engine
// (mli_engine_*): requests are uploaded once, the scheduler / page manager run on the GPU, and the
// finished token lists come back in finish order.  The caller-visible post-state matches the
// reference: item_storage.finished holds every request in finish order, the new-item queue and
// processing_storage are empty, every page is back in memory_block_manager, and the global
// ThroughputCounter has the job's tokens and wall time (and prints them, as the reference does).
#include <chrono>

#include "mli/compat.hpp"

void start_inference_engine(const TensorFloat& emb_table, const TensorFloat& pos_table,
                            ItemStorage& item_storage, ProcessingStorage& processing_storage,
                            InferenceModel& inference_model, size_t n_batch_size, size_t n_sequence) {
    std::vector<int> free_rows(n_batch_size);
    for (size_t r = 0; r < n_batch_size; ++r) free_rows[r] = static_cast<int>(r);
    TensorInt inp_device({n_batch_size, n_sequence}, DeviceType::DEVICE);
    TensorInt inp_host({n_batch_size, n_sequence}, DeviceType::HOST);
    TensorInt lengths_device({n_batch_size}, DeviceType::DEVICE);
    TensorInt lengths_host({n_batch_size}, DeviceType::HOST);
    TensorInt idx_device({n_batch_size}, DeviceType::DEVICE);
    TensorInt idx_host({n_batch_size}, DeviceType::HOST);
    TensorInt result_device({n_batch_size}, DeviceType::DEVICE);
    TensorInt result_host({n_batch_size}, DeviceType::HOST);
    cuda_check(cudaMemset(lengths_device.data(), 0, n_batch_size * sizeof(int)), __FILE__, __LINE__);
    cuda_check(cudaMemset(inp_device.data(), 0, n_batch_size * n_sequence * sizeof(int)), __FILE__, __LINE__);

    int n_new = insert_new_items(free_rows, inp_device, inp_host, lengths_device, lengths_host,
                                 idx_device, idx_host, item_storage, processing_storage);
    while (!is_done(item_storage, processing_storage)) {
        inference_model.forward(inp_device, lengths_device, idx_device, result_device, n_new, emb_table,
                                pos_table);
        free_rows = process_decoder_result(result_device, result_host, item_storage, processing_storage,
                                           static_cast<int>(n_sequence));
        n_new = insert_new_items(free_rows, inp_device, inp_host, lengths_device, lengths_host,
                                 idx_device, idx_host, item_storage, processing_storage);
    }
}

namespace {

void run_paged_job(const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
                   ProcessingStorage& processing_storage, MemoryBlockManager& pool,
                   const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv, size_t n_batch,
                   size_t n_sequence, size_t emb_dim, int n_forward_rounds) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    // drain the queue: request k of the job is the k-th queued item (ids are the caller's)
    std::vector<IdTokensPair> reqs = item_storage.pop_new_items(item_storage.new_count());
    const int n_req = static_cast<int>(reqs.size());
    std::vector<int> offsets(n_req + 1, 0), tokens;
    for (int k = 0; k < n_req; ++k) {
        tokens.insert(tokens.end(), reqs[k].second.begin(), reqs[k].second.end());
        offsets[k + 1] = static_cast<int>(tokens.size());
    }
    if (pool.free_blocks_size() != pool.total_blocks())
        throw std::runtime_error("paged engine: memory_block_manager must start with every block free");

    mli_engine_cfg cfg{};
    cfg.n_batch = static_cast<int>(n_batch);
    cfg.n_sequence = static_cast<int>(n_sequence);
    cfg.emb_dim = static_cast<int>(emb_dim);
    cfg.n_vocab = static_cast<int>(emb_table.shape()[0]);
    cfg.n_blocks = pool.total_blocks();
    cfg.n_forward_rounds = n_forward_rounds;
    cfg.compat_stale_lengths = mli::fix_stale_lengths() ? 0 : 1;
    cfg.max_requests = n_req > 0 ? n_req : 1;
    cfg.page_pool = pool.slab();  // the engine carves its pages from the caller's slab

    mli_engine* engine = nullptr;
    mli::check(mli_engine_create(mli::host_context(), &cfg, emb_table.data(), pos_table.data(),
                                 wk.data(), wq.data(), wv.data(), &engine));
    std::vector<int> fin_ids(n_req > 0 ? n_req : 1), fin_offs(n_req + 1),
        fin_toks(static_cast<size_t>(n_req > 0 ? n_req : 1) * n_sequence);
    int n_fin = 0;
    mli_engine_stats stats{};
    try {
        mli::check(mli_engine_submit(engine, n_req, offsets.data(), tokens.data(), 0));
        mli::check(mli_engine_run(engine, 0, 0));
        mli::check(mli_engine_results(engine, fin_ids.data(), fin_offs.data(), fin_toks.data(), &n_fin));
        mli::check(mli_engine_get_stats(engine, &stats));
    } catch (...) {
        mli_engine_destroy(engine);
        throw;
    }
    mli_engine_destroy(engine);
    for (int k = 0; k < n_fin; ++k) {
        const int q = fin_ids[k];
        item_storage.add_finished_item(IdTokensPair(
            reqs[q].first, std::vector<int>(fin_toks.begin() + fin_offs[k], fin_toks.begin() + fin_offs[k + 1])));
    }
    (void)processing_storage;  // nothing is left processing when the job returns
    const auto t1 = std::chrono::high_resolution_clock::now();
    ThroughputCounter& counter = get_global_throughput_counter();
    counter.add_job(stats.generated_tokens, std::chrono::duration<double>(t1 - t0).count());
    counter.print_throughput();
}

}  // namespace

void start_paged_attention_inference_engine(
    const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
    ProcessingStorage& processing_storage, MemoryBlockManager& memory_block_manager,
    PagedAttentionsManager&, PagedAttentionInferenceModel& inference_model, size_t n_batch_size,
    size_t n_sequence, int n_forward_rounds) {
    const PagedAttentionLayer& layer = inference_model.attention_layer();
    run_paged_job(emb_table, pos_table, item_storage, processing_storage, memory_block_manager,
                  layer.wk(), layer.wq(), layer.wv(), n_batch_size, n_sequence,
                  inference_model.emb_dim(), n_forward_rounds);
}

void start_paged_attention_cublas_inference_engine(
    const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
    ProcessingStorage& processing_storage, MemoryBlockManager& memory_block_manager,
    PagedAttentionsManager&, PagedAttentionCublasInferenceModel& inference_model, size_t n_batch_size,
    size_t n_sequence, int n_forward_rounds) {
    const PagedAttentionCublasLayer& layer = inference_model.attention_layer();
    run_paged_job(emb_table, pos_table, item_storage, processing_storage, memory_block_manager,
                  layer.wk(), layer.wq(), layer.wv(), n_batch_size, n_sequence,
                  inference_model.emb_dim(), n_forward_rounds);
}
