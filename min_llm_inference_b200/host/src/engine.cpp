// engine.cpp -- the three engine loops of include/inferencer.h.
//
// All three (the non-paged start_inference_engine, src/inferencer.cpp:11-41, and the two paged
// engines, :43-133) hand the whole job to the on-device engine
// (mli_engine_*): requests are uploaded once, the scheduler / page manager run on the GPU, and the
// finished token lists come back in finish order.  The caller-visible post-state matches the
// reference: item_storage.finished holds every request in finish order, the new-item queue and
// processing_storage are empty, every page is back in memory_block_manager, and the global
// ThroughputCounter has the job's tokens and wall time (and prints them, as the reference does).
#include <chrono>
#include <cstdlib>
#include <thread>

#include "mli/compat.hpp"

namespace {
void run_device_job(const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
                    float* slab, int n_blocks, const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv,
                    size_t n_batch, size_t n_sequence, size_t emb_dim, int n_forward_rounds, bool compat);

// the reference's host loop (src/inferencer.cpp:11-41) over InferenceModel::forward, kept for
// MLI_DENSE_HOST_LOOP=1: one blocking D2H + up to three blocking H2D copies per generated token
void dense_host_loop(const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
                     ProcessingStorage& processing_storage, InferenceModel& inference_model, size_t n_batch_size,
                     size_t n_sequence) {
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
}  // namespace

// start_inference_engine (non-paged, BASELINE configs[0]; src/inferencer.cpp:11-41).  The reference
// keeps dense K^T / V caches of n_batch * n_sequence * emb_dim floats inside the model; they are not
// visible at this level, so the B200 build runs the job on the SAME on-device engine as the paged
// entry points, with a private pool that can never run dry (n_batch * n_sequence / 16 pages: every row
// can reach n_sequence) and corrected lengths -- the non-paged scheduler (item_storage.cpp:141-180) is the
// paged one without page pressure: queue heads go to free rows in row order, finished requests are
// reported in row order.  Tokens and finish order equal the reference's (tests/test_gpu_dropin.py,
// tests/test_gpu_forward_engine.py P2).  MLI_DENSE_HOST_LOOP=1 keeps the reference's host loop over the
// dense stage kernels instead.
void start_inference_engine(const TensorFloat& emb_table, const TensorFloat& pos_table,
                            ItemStorage& item_storage, ProcessingStorage& processing_storage,
                            InferenceModel& inference_model, size_t n_batch_size, size_t n_sequence) {
    const char* host_loop = std::getenv("MLI_DENSE_HOST_LOOP");
    if ((host_loop && host_loop[0] == '1') || n_sequence % PAGE_BLOCK_SIZE != 0) {
        dense_host_loop(emb_table, pos_table, item_storage, processing_storage, inference_model, n_batch_size,
                        n_sequence);
        return;
    }
    const SelfAttentionLayer& layer = inference_model.attention_layer();
    const int W = static_cast<int>(n_sequence) / PAGE_BLOCK_SIZE;
    const int n_blocks = static_cast<int>(n_batch_size) * std::max(W, DEFAULT_INIT_NUM_BLOCKS);
    run_device_job(emb_table, pos_table, item_storage, nullptr, n_blocks, layer.wk(), layer.wq(), layer.wv(),
                   n_batch_size, n_sequence, inference_model.emb_dim(), 1, false);
}

namespace mli {
static int g_num_gpus = -1;
void set_num_gpus(int n) { g_num_gpus = n < 1 ? 1 : n; }
int num_gpus() {
    if (g_num_gpus < 0) {
        const char* e = std::getenv("MLI_NUM_GPUS");
        g_num_gpus = (e && atoi(e) > 1) ? atoi(e) : 1;
    }
    return g_num_gpus;
}
}  // namespace mli

namespace {

// Request-sharded job on G GPUs of this process: the queue is cut into G contiguous
// blocks, GPU g runs its own device engine (rows, KV pages, scheduler) on block g, and ONE
// collective -- mli_comm_gather_tokens, an NCCL all-gather of the request tables -- brings every
// token list back.  Weights and tables are replicated from the caller's tensors (GPU 0); GPU 0 uses
// the caller's slab, the other GPUs allocate a pool of the same size.  Finished requests are
// reported rank by rank, each rank in its own finish order.
void run_paged_job_sharded(const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
                           MemoryBlockManager& pool, const TensorFloat& wk, const TensorFloat& wq,
                           const TensorFloat& wv, size_t n_batch, size_t n_sequence, size_t emb_dim,
                           int n_forward_rounds, int G) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    int n_dev = 0;
    cuda_check(cudaGetDeviceCount(&n_dev), __FILE__, __LINE__);
    if (G > n_dev) throw std::runtime_error("MLI_NUM_GPUS exceeds the visible devices");
    int dev0 = 0;
    cudaGetDevice(&dev0);
    if (dev0 != 0) throw std::runtime_error("sharded engine: the caller's tensors must live on device 0");
    std::vector<IdTokensPair> reqs = item_storage.pop_new_items(item_storage.new_count());
    const int n_req = static_cast<int>(reqs.size());
    const int per = (n_req + G - 1) / G > 0 ? (n_req + G - 1) / G : 1;
    const int rows = static_cast<int>((n_batch + G - 1) / G);
    const int S = static_cast<int>(n_sequence), d = static_cast<int>(emb_dim);
    const int V = static_cast<int>(emb_table.shape()[0]);

    struct Shard {
        mli_ctx* ctx = nullptr;
        mli_engine* engine = nullptr;
        mli_comm* comm = nullptr;
        std::vector<void*> dev_allocs;
        const float *emb = nullptr, *pos = nullptr, *wk = nullptr, *wq = nullptr, *wv = nullptr;
        int *all_tok = nullptr, *all_cnt = nullptr;
        std::vector<int> offsets, tokens, fin_ids, fin_offs, fin_toks;
        int n_local = 0, n_fin = 0;
        mli_engine_stats stats{};
        int rc = 0;
        std::string err;
    };
    std::vector<Shard> sh(G);
    auto cleanup = [&]() {
        for (int g = 0; g < G; ++g) {
            cudaSetDevice(g);
            if (sh[g].comm) mli_comm_destroy(sh[g].comm);
            if (sh[g].engine) mli_engine_destroy(sh[g].engine);
            for (void* p : sh[g].dev_allocs) cudaFree(p);
            if (g > 0 && sh[g].ctx) mli_ctx_destroy(sh[g].ctx);
        }
        cudaSetDevice(0);
    };
    try {
        for (int g = 0; g < G; ++g) {
            Shard& s = sh[g];
            cuda_check(cudaSetDevice(g), __FILE__, __LINE__);
            if (g == 0) {
                s.ctx = mli::host_context();
                s.emb = emb_table.data(); s.pos = pos_table.data();
                s.wk = wk.data(); s.wq = wq.data(); s.wv = wv.data();
            } else {
                mli::check(mli_ctx_create(&s.ctx, g, nullptr));
                int mode = 0;
                mli::check(mli_ctx_get_option(mli::host_context(), MLI_OPT_GEMM_MODE, &mode));
                mli::check(mli_ctx_set_option(s.ctx, MLI_OPT_GEMM_MODE, mode));
                auto replicate = [&](const float* src, size_t n) {
                    void* p = nullptr;
                    cuda_check(cudaMalloc(&p, n * sizeof(float)), __FILE__, __LINE__);
                    s.dev_allocs.push_back(p);
                    cuda_check(cudaMemcpyPeer(p, g, src, 0, n * sizeof(float)), __FILE__, __LINE__);
                    return static_cast<const float*>(p);
                };
                s.emb = replicate(emb_table.data(), (size_t)V * d);
                s.pos = replicate(pos_table.data(), (size_t)S * d);
                s.wk = replicate(wk.data(), (size_t)d * d);
                s.wq = replicate(wq.data(), (size_t)d * d);
                s.wv = replicate(wv.data(), (size_t)d * d);
            }
            const int lo = std::min(g * per, n_req), hi = std::min((g + 1) * per, n_req);
            s.n_local = hi - lo;
            s.offsets.assign(s.n_local + 1, 0);
            for (int k = lo; k < hi; ++k) {
                s.tokens.insert(s.tokens.end(), reqs[k].second.begin(), reqs[k].second.end());
                s.offsets[k - lo + 1] = static_cast<int>(s.tokens.size());
            }
            if (s.tokens.empty()) s.tokens.push_back(0);
            mli_engine_cfg cfg{};
            cfg.n_batch = rows;
            cfg.n_sequence = S;
            cfg.emb_dim = d;
            cfg.n_vocab = V;
            cfg.n_blocks = pool.total_blocks();
            cfg.n_forward_rounds = n_forward_rounds;
            cfg.compat_stale_lengths = mli::fix_stale_lengths() ? 0 : 1;
            cfg.max_requests = per;
            cfg.page_pool = (g == 0) ? pool.slab() : nullptr;
            mli::check(mli_engine_create(s.ctx, &cfg, s.emb, s.pos, s.wk, s.wq, s.wv, &s.engine));
            void* p = nullptr;
            cuda_check(cudaMalloc(&p, sizeof(int) * (size_t)G * per * S), __FILE__, __LINE__);
            s.dev_allocs.push_back(p);
            s.all_tok = static_cast<int*>(p);
            cuda_check(cudaMalloc(&p, sizeof(int) * (size_t)G * per), __FILE__, __LINE__);
            s.dev_allocs.push_back(p);
            s.all_cnt = static_cast<int*>(p);
        }
        std::vector<mli_ctx*> ctxs(G);
        std::vector<mli_comm*> comms(G, nullptr);
        for (int g = 0; g < G; ++g) ctxs[g] = sh[g].ctx;
        mli::check(mli_comm_init_all(ctxs.data(), G, comms.data()));
        for (int g = 0; g < G; ++g) sh[g].comm = comms[g];

        // one host thread per GPU: mli_engine_run blocks until its shard is done
        std::vector<std::thread> threads;
        for (int g = 0; g < G; ++g)
            threads.emplace_back([&, g]() {
                Shard& s = sh[g];
                cudaSetDevice(g);
                s.fin_ids.assign(std::max(s.n_local, 1), 0);
                s.fin_offs.assign(s.n_local + 1, 0);
                s.fin_toks.assign((size_t)std::max(s.n_local, 1) * S, 0);
                int rc = mli_engine_submit(s.engine, s.n_local, s.offsets.data(), s.tokens.data(), 0);
                if (!rc) rc = mli_engine_run(s.engine, 0, 0);
                if (!rc) rc = mli_engine_results(s.engine, s.fin_ids.data(), s.fin_offs.data(), s.fin_toks.data(), &s.n_fin);
                if (!rc) rc = mli_engine_get_stats(s.engine, &s.stats);
                if (rc) { s.rc = rc; s.err = mli_last_error(); }
            });
        for (auto& t : threads) t.join();
        for (int g = 0; g < G; ++g)
            if (sh[g].rc) {
                if (sh[g].rc == MLI_ERR_NO_BLOCKS) throw std::runtime_error("No enough block memories to return");
                printf("%s\n", sh[g].err.c_str());
                throw std::runtime_error(sh[g].rc == MLI_ERR_CUDA ? "Cuda Failure" : sh[g].err);
            }
        // the collective: every rank receives every rank's request table
        mli::check(mli_comm_group_start());
        for (int g = 0; g < G; ++g)
            mli::check(mli_comm_gather_tokens(sh[g].comm, sh[g].engine, per, sh[g].all_tok, sh[g].all_cnt));
        mli::check(mli_comm_group_end());
        cuda_check(cudaSetDevice(0), __FILE__, __LINE__);
        mli::check(mli_ctx_synchronize(sh[0].ctx));
        std::vector<int> all_tok((size_t)G * per * S), all_cnt((size_t)G * per);
        cuda_check(cudaMemcpy(all_tok.data(), sh[0].all_tok, all_tok.size() * sizeof(int), cudaMemcpyDeviceToHost), __FILE__, __LINE__);
        cuda_check(cudaMemcpy(all_cnt.data(), sh[0].all_cnt, all_cnt.size() * sizeof(int), cudaMemcpyDeviceToHost), __FILE__, __LINE__);
        long long generated = 0;
        for (int g = 0; g < G; ++g) {
            generated += sh[g].stats.generated_tokens;
            for (int k = 0; k < sh[g].n_fin; ++k) {
                const int slot = g * per + sh[g].fin_ids[k];     // row of the gathered table
                const int q = std::min(g * per, n_req) + sh[g].fin_ids[k];
                const int* t = all_tok.data() + (size_t)slot * S;
                item_storage.add_finished_item(IdTokensPair(reqs[q].first, std::vector<int>(t, t + all_cnt[slot])));
            }
        }
        cleanup();
        const auto t1 = std::chrono::high_resolution_clock::now();
        ThroughputCounter& counter = get_global_throughput_counter();
        counter.add_job(generated, std::chrono::duration<double>(t1 - t0).count());
        counter.print_throughput();
    } catch (...) {
        cleanup();
        throw;
    }
}

void run_paged_job(const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
                   ProcessingStorage& processing_storage, MemoryBlockManager& pool,
                   const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv, size_t n_batch,
                   size_t n_sequence, size_t emb_dim, int n_forward_rounds) {
    if (mli::num_gpus() > 1) {
        (void)processing_storage;
        run_paged_job_sharded(emb_table, pos_table, item_storage, pool, wk, wq, wv, n_batch, n_sequence, emb_dim,
                              n_forward_rounds, mli::num_gpus());
        return;
    }
    if (pool.free_blocks_size() != pool.total_blocks())
        throw std::runtime_error("paged engine: memory_block_manager must start with every block free");
    (void)processing_storage;  // nothing is left processing when the job returns
    // the engine carves its pages from the caller's slab
    run_device_job(emb_table, pos_table, item_storage, pool.slab(), pool.total_blocks(), wk, wq, wv, n_batch,
                   n_sequence, emb_dim, n_forward_rounds, !mli::fix_stale_lengths());
}

void run_device_job(const TensorFloat& emb_table, const TensorFloat& pos_table, ItemStorage& item_storage,
                    float* slab, int n_blocks, const TensorFloat& wk, const TensorFloat& wq, const TensorFloat& wv,
                    size_t n_batch, size_t n_sequence, size_t emb_dim, int n_forward_rounds, bool compat) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    // drain the queue: request k of the job is the k-th queued item (ids are the caller's)
    std::vector<IdTokensPair> reqs = item_storage.pop_new_items(item_storage.new_count());
    const int n_req = static_cast<int>(reqs.size());
    std::vector<int> offsets(n_req + 1, 0), tokens;
    for (int k = 0; k < n_req; ++k) {
        tokens.insert(tokens.end(), reqs[k].second.begin(), reqs[k].second.end());
        offsets[k + 1] = static_cast<int>(tokens.size());
    }

    mli_engine_cfg cfg{};
    cfg.n_batch = static_cast<int>(n_batch);
    cfg.n_sequence = static_cast<int>(n_sequence);
    cfg.emb_dim = static_cast<int>(emb_dim);
    cfg.n_vocab = static_cast<int>(emb_table.shape()[0]);
    cfg.n_blocks = n_blocks;
    cfg.n_forward_rounds = n_forward_rounds;
    cfg.compat_stale_lengths = compat ? 1 : 0;
    cfg.max_requests = n_req > 0 ? n_req : 1;
    cfg.page_pool = slab;  // nullptr: the engine allocates its own pool
    // opt-in scheduling policies of the device engine (off = the reference's behaviour), for callers of the
    // reference's entry points: MLI_PREFILL_CHUNK / MLI_MAX_PREFILL (positions per step)
    if (!compat) {
        if (const char* v = std::getenv("MLI_PREFILL_CHUNK")) cfg.prefill_chunk_positions = atoi(v);
    }
    if (const char* v = std::getenv("MLI_MAX_PREFILL")) cfg.max_prefill_positions = atoi(v);

    mli_engine* engine = nullptr;
    mli::check(mli_engine_create(mli::host_context(), &cfg, emb_table.data(), pos_table.data(),
                                 wk.data(), wq.data(), wv.data(), &engine));
    std::vector<int> fin_ids(n_req > 0 ? n_req : 1), fin_offs(n_req + 1),
        fin_toks(static_cast<size_t>(n_req > 0 ? n_req : 1) * n_sequence);
    int n_fin = 0;
    mli_engine_stats stats{};
    int run_rc = MLI_OK;
    try {
        mli::check(mli_engine_submit(engine, n_req, offsets.data(), tokens.data(), 0));
        // a job that cannot complete (a request outgrew the pool: the reference would spin) still hands back
        // what did finish before the exception below
        run_rc = mli_engine_run(engine, 0, 0);
        if (run_rc != MLI_ERR_NO_BLOCKS) mli::check(run_rc);
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
    if (run_rc != MLI_OK) throw std::runtime_error("No enough block memories to return");
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
