// engine_driver.cpp -- ONE driver source written purely against the reference's public C++ API
// (inferencer.h, inference_model.h, item_storage.h, paged_item_storage.h, tensor.hpp).  It is
// compiled twice by tests/dropin/Makefile:
//   * against /root/reference/include + the reference objects  -> oracle/_ref/dropin_driver_ref
//   * against min_llm_inference_b200/host/include + our libs   -> tests/dropin/_build/dropin_driver_mli
// and tests/test_gpu_dropin.py checks that both print the same finished token lists.  That is the
// drop-in claim, tested: no source change is needed to move a caller from the reference to this repo.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "constants.h"
#include "inference_model.h"
#include "inferencer.h"
#include "item_storage.h"
#include "paged_item_storage.h"
#include "tensor.hpp"

namespace {

struct Lcg {  // fixed, implementation-independent generator (the reference's own fixtures are unseeded)
    uint64_t s;
    explicit Lcg(uint64_t seed) : s(seed * 2862933555777941757ULL + 3037000493ULL) {}
    uint32_t next() {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        return static_cast<uint32_t>(s >> 33);
    }
    float unit() { return (next() & 0xFFFFFF) / 16777216.0f; }  // [0, 1)
};

TensorFloat random_tensor(const std::vector<size_t>& shape, Lcg& rng, float scale, float shift) {
    TensorFloat host(shape, DeviceType::HOST);
    float* p = host.data();
    for (size_t i = 0; i < host.get_total_size(); ++i) p[i] = (rng.unit() + shift) * scale;
    TensorFloat dev(shape, DeviceType::DEVICE);
    dev.copy_from(host);
    return dev;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 12) {
        fprintf(stderr, "usage: %s dense|paged|paged_cublas B S d V n_blocks n_req lo hi seed dist(R|Z) [rounds]\n", argv[0]);
        return 2;
    }
    const std::string kind = argv[1];
    const size_t B = atoi(argv[2]), S = atoi(argv[3]), d = atoi(argv[4]), V = atoi(argv[5]);
    const int n_blocks = atoi(argv[6]), n_req = atoi(argv[7]), lo = atoi(argv[8]), hi = atoi(argv[9]);
    const uint64_t seed = strtoull(argv[10], nullptr, 10);
    const bool dist_r = argv[11][0] == 'R';
    const int rounds = argc > 12 ? atoi(argv[12]) : 1;

    Lcg rng(seed);
    const float wscale = dist_r ? 1.0f : 1.5f * std::sqrt(12.0f / d);
    const float shift = dist_r ? 0.0f : -0.5f;
    TensorFloat wk = random_tensor({d, d}, rng, wscale, shift);
    TensorFloat wq = random_tensor({d, d}, rng, wscale, shift);
    TensorFloat wv = random_tensor({d, d}, rng, wscale, shift);
    TensorFloat emb_table = random_tensor({V, d}, rng, dist_r ? 1.0f : 2.0f, shift);
    TensorFloat pos_table = random_tensor({S, d}, rng, dist_r ? 1.0f : 0.5f, shift);

    ItemStorage item_storage;
    ProcessingStorage processing_storage;
    for (int i = 0; i < n_req; ++i) {
        const int len = lo + static_cast<int>(rng.next() % static_cast<uint32_t>(hi - lo + 1));
        std::vector<int> toks(len);
        for (int& t : toks) t = static_cast<int>(rng.next() % static_cast<uint32_t>(EOF_TOKEN_ID));
        item_storage.add_new_item(IdTokensPair(1000 + i, std::move(toks)));
    }

    if (kind == "dense") {
        InferenceModel model(SelfAttentionLayer(std::move(wk), std::move(wq), std::move(wv), B, d, S),
                             EncoderLayer(), DecoderLayer(B, V), B, S, d);
        start_inference_engine(emb_table, pos_table, item_storage, processing_storage, model, B, S);
    } else {
        PagedAttentionsManager paged_attention_manager(B, S, d);
        MemoryBlockManager memory_block_manager(n_blocks, PAGE_BLOCK_SIZE * 3 * d);
        if (kind == "paged") {
            PagedAttentionInferenceModel model(
                PagedAttentionLayer(std::move(wk), std::move(wq), std::move(wv), B, d, S),
                PagedEncoderLayer(), PagedDecoderLayer(B, V), B, S, d, rounds);
            start_paged_attention_inference_engine(emb_table, pos_table, item_storage, processing_storage,
                                                   memory_block_manager, paged_attention_manager, model,
                                                   B, S, rounds);
        } else {
            PagedAttentionCublasInferenceModel model(
                PagedAttentionCublasLayer(std::move(wk), std::move(wq), std::move(wv), B, d, S),
                PagedEncoderLayer(), PagedCublasDecoderLayer(B, V), B, S, d, rounds);
            start_paged_attention_cublas_inference_engine(emb_table, pos_table, item_storage,
                                                          processing_storage, memory_block_manager,
                                                          paged_attention_manager, model, B, S, rounds);
        }
        if (memory_block_manager.free_blocks_size() != n_blocks) {
            fprintf(stderr, "pages leaked: %d of %d free\n", memory_block_manager.free_blocks_size(), n_blocks);
            return 3;
        }
    }
    if (item_storage.finish_count() != n_req || item_storage.new_count() != 0 || processing_storage.size() != 0) {
        fprintf(stderr, "engine did not drain: finished %d new %d processing %d\n", item_storage.finish_count(),
                item_storage.new_count(), processing_storage.size());
        return 4;
    }
    // one line per finished request, in finish order: "RESULT id : tokens..."
    for (const IdTokensPair& p : item_storage.get_finished_items()) {
        printf("RESULT %d :", p.first);
        for (int t : p.second) printf(" %d", t);
        printf("\n");
    }
    return 0;
}
