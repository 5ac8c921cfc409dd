// stepwise_driver.cpp -- the STEP-level part of the drop-in claim: a caller that writes the engine loop itself, one
// forward() at a time, against the reference's scheduling API
//   include/paged_item_storage.h:10-57 (MemoryBlockManager, PagedAttentionsManager, paged insert_new_items,
//   allocate_or_free_memory_blocks_if_needed), include/item_storage.h:64-90 (process_decoder_result, non-paged
//   insert_new_items, is_done), include/inference_model.h:8-74 (forward)
// -- the sequence of calls of src/inferencer.cpp:11-41 and :43-85, written out by the user.  After every iteration it
// prints what the scheduler decided (rows that received requests, rows that finished, free pages, the page list of
// every resident row as slab indices, device lengths, tokens).  tests/dropin/Makefile compiles it against the
// reference tree and against this repo's host mirror; tests/test_gpu_dropin.py requires identical output, i.e.
// identical decisions at every step (including the physical order of the free list and the reference's stale
// lengths), not only identical final token lists.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "constants.h"
#include "inference_model.h"
#include "item_storage.h"
#include "paged_item_storage.h"
#include "tensor.hpp"

namespace {

struct Lcg {
    uint64_t s;
    explicit Lcg(uint64_t seed) : s(seed * 2862933555777941757ULL + 3037000493ULL) {}
    uint32_t next() {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        return static_cast<uint32_t>(s >> 33);
    }
    float unit() { return (next() & 0xFFFFFF) / 16777216.0f; }
};

TensorFloat random_tensor(const std::vector<size_t>& shape, Lcg& rng, float scale, float shift) {
    TensorFloat host(shape, DeviceType::HOST);
    float* p = host.data();
    for (size_t i = 0; i < host.get_total_size(); ++i) p[i] = (rng.unit() + shift) * scale;
    TensorFloat dev(shape, DeviceType::DEVICE);
    dev.copy_from(host);
    return dev;
}

std::vector<int> ints_to_host(const TensorInt& t) {
    std::vector<int> v(t.get_total_size());
    cudaDeviceSynchronize();
    cudaMemcpy(v.data(), t.data(), v.size() * sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

void print_list(const char* tag, const std::vector<int>& v) {
    printf(" %s=[", tag);
    for (size_t i = 0; i < v.size(); ++i) printf(i ? ",%d" : "%d", v[i]);
    printf("]");
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 12) {
        fprintf(stderr, "usage: %s dense|paged B S d V n_blocks n_req lo hi seed dist(R|Z) [rounds]\n", argv[0]);
        return 2;
    }
    const std::string kind = argv[1];
    const size_t B = atoi(argv[2]), S = atoi(argv[3]), d = atoi(argv[4]), V = atoi(argv[5]);
    const int n_blocks = atoi(argv[6]), n_req = atoi(argv[7]), lo = atoi(argv[8]), hi = atoi(argv[9]);
    const uint64_t seed = strtoull(argv[10], nullptr, 10);
    const bool dist_r = argv[11][0] == 'R';
    const int rounds = argc > 12 ? atoi(argv[12]) : 1;
    const int max_printed = 80;

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
        item_storage.add_new_item(IdTokensPair(500 + i, std::move(toks)));
    }

    TensorInt inp_device({B, S}, DeviceType::DEVICE), inp_host({B, S}, DeviceType::HOST);
    TensorInt lengths_device({B}, DeviceType::DEVICE), lengths_host({B}, DeviceType::HOST);
    TensorInt new_idx_device({B}, DeviceType::DEVICE), new_idx_host({B}, DeviceType::HOST);
    cudaMemset(lengths_device.data(), 0, B * sizeof(int));
    for (size_t i = 0; i < B; ++i) lengths_host.data()[i] = 0;
    int iteration = 0;

    if (kind == "dense") {
        InferenceModel model(SelfAttentionLayer(std::move(wk), std::move(wq), std::move(wv), B, d, S), EncoderLayer(),
                             DecoderLayer(B, V), B, S, d);
        TensorInt decoder_device({B}, DeviceType::DEVICE), decoder_host({B}, DeviceType::HOST);
        std::vector<int> free_rows(B);
        for (size_t i = 0; i < B; ++i) free_rows[i] = static_cast<int>(i);
        int n_new = insert_new_items(free_rows, inp_device, inp_host, lengths_device, lengths_host, new_idx_device,
                                     new_idx_host, item_storage, processing_storage);
        while (!is_done(item_storage, processing_storage) && iteration < 100000) {
            model.forward(inp_device, lengths_device, new_idx_device, decoder_device, n_new, emb_table, pos_table);
            const std::vector<int> finished =
                process_decoder_result(decoder_device, decoder_host, item_storage, processing_storage, static_cast<int>(S));
            if (iteration < max_printed) {
                printf("STEP %d n_new=%d", iteration, n_new);
                print_list("finished", finished);
                print_list("tokens", ints_to_host(decoder_device));
                print_list("lengths", ints_to_host(lengths_device));
                printf("\n");
            }
            n_new = insert_new_items(finished, inp_device, inp_host, lengths_device, lengths_host, new_idx_device,
                                     new_idx_host, item_storage, processing_storage);
            ++iteration;
        }
    } else {
        PagedAttentionsManager manager(B, S, d);
        MemoryBlockManager blocks(n_blocks, PAGE_BLOCK_SIZE * 3 * d);
        // slab base and page size in floats, to print pages as slab indices: take everything once and give it back
        const float* base = nullptr;
        {
            std::list<float*> all = blocks.pop_free_blocks(n_blocks);
            base = *std::min_element(all.begin(), all.end());
            blocks.return_free_blocks(std::move(all));
        }
        const size_t page_floats = PAGE_BLOCK_SIZE * 3 * d;
        PagedAttentionInferenceModel model(PagedAttentionLayer(std::move(wk), std::move(wq), std::move(wv), B, d, S),
                                           PagedEncoderLayer(), PagedDecoderLayer(B, V), B, S, d, rounds);
        TensorInt decoder_device({B, static_cast<size_t>(rounds)}, DeviceType::DEVICE),
            decoder_host({B, static_cast<size_t>(rounds)}, DeviceType::HOST);
        std::vector<int> new_rows = insert_new_items(inp_device, inp_host, lengths_device, lengths_host, new_idx_device,
                                                     new_idx_host, item_storage, processing_storage, blocks, manager,
                                                     rounds);
        while (!is_done(item_storage, processing_storage) && iteration < 100000) {
            model.forward(inp_device, lengths_device, new_idx_device, decoder_device, static_cast<int>(new_rows.size()),
                          emb_table, pos_table, manager.get_page_table_device());
            const std::vector<int> tokens = ints_to_host(decoder_device);
            const std::vector<int> finished =
                process_decoder_result(decoder_device, decoder_host, item_storage, processing_storage, static_cast<int>(S));
            allocate_or_free_memory_blocks_if_needed(manager, blocks, processing_storage, item_storage, finished, rounds);
            if (iteration < max_printed) {
                printf("STEP %d", iteration);
                print_list("new_rows", new_rows);
                print_list("finished", finished);
                print_list("tokens", tokens);
                print_list("lengths", ints_to_host(lengths_device));
                printf(" free=%d pages={", blocks.free_blocks_size());
                for (const BatchIdMemoryBlocksPair& p : manager.get_used_block_list()) {
                    printf(" %d:", p.first);
                    for (const float* q : p.second) printf("%d.", static_cast<int>((q - base) / page_floats));
                }
                printf(" }\n");
            }
            new_rows = insert_new_items(inp_device, inp_host, lengths_device, lengths_host, new_idx_device, new_idx_host,
                                        item_storage, processing_storage, blocks, manager, rounds);
            ++iteration;
        }
        if (blocks.free_blocks_size() != n_blocks) {
            fprintf(stderr, "pages leaked: %d of %d free\n", blocks.free_blocks_size(), n_blocks);
            return 3;
        }
    }
    printf("ITERATIONS %d\n", iteration);
    if (item_storage.finish_count() != n_req) {
        fprintf(stderr, "did not drain: %d of %d finished\n", item_storage.finish_count(), n_req);
        return 4;
    }
    for (const IdTokensPair& p : item_storage.get_finished_items()) {
        printf("RESULT %d :", p.first);
        for (int t : p.second) printf(" %d", t);
        printf("\n");
    }
    return 0;
}
