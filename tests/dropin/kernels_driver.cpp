// kernels_driver.cpp -- the KERNEL-level half of the drop-in claim (engine_driver.cpp is the engine-level half).
// ONE source written purely against the reference's kernel launchers and page-table classes
//   include/kernels/encoder.h:5-25, include/kernels/paged_attention.h:17-67,
//   include/kernels/self_attention_inference_optimized.h:21-50, include/kernels/decoder.h:19-31,
//   include/paged_item_storage.h:10-46, include/tensor.hpp
// i.e. the calls the reference's tests/paged_attention_kernels_test.cpp, encoder_test.cpp and decoder_test.cpp
// make (those need gtest, which is not installed, and unseeded fixtures; this driver feeds fixed tensors).
// tests/dropin/Makefile compiles it twice, against /root/reference/include + the reference objects and against
// min_llm_inference_b200/host/include + our libraries; tests/test_gpu_dropin.py compares what the two print:
//   HASH <stage> <fnv1a of the float / int bits>   stages our exact mode reproduces bit for bit
//   VALS <stage> v0 v1 ...                         the fused attention block (summation order differs: rel 1e-4)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "constants.h"
#include "kernels/decoder.h"
#include "kernels/encoder.h"
#include "kernels/paged_attention.h"
#include "kernels/self_attention_inference_optimized.h"
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

TensorInt int_tensor(const std::vector<int>& v, const std::vector<size_t>& shape) {
    TensorInt host(shape, DeviceType::HOST);
    for (size_t i = 0; i < v.size(); ++i) host.data()[i] = v[i];
    TensorInt dev(shape, DeviceType::DEVICE);
    dev.copy_from(host);
    return dev;
}

uint64_t fnv(const void* data, size_t bytes, uint64_t h = 1469598103934665603ULL) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < bytes; ++i) {
        h ^= p[i];
        h *= 1099511628211ULL;
    }
    return h;
}

template <typename T>
std::vector<T> to_host(const T* dev, size_t n) {
    std::vector<T> v(n);
    cudaDeviceSynchronize();
    if (cudaMemcpy(v.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) {
        fprintf(stderr, "copy to host failed\n");
        exit(5);
    }
    return v;
}

// sub-row `which` (0 embedding, 1 K, 2 V) of the first lengths[r] positions of every row, through the page pointers
uint64_t hash_pages(const std::vector<std::vector<float*>>& pages, const std::vector<int>& len, size_t d, int which,
                    int extra) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t r = 0; r < pages.size(); ++r)
        for (int j = 0; j < len[r] + (len[r] > 0 ? extra : 0); ++j) {
            const float* src = pages[r][j / PAGE_BLOCK_SIZE] + static_cast<size_t>(j % PAGE_BLOCK_SIZE) * 3 * d + which * d;
            const std::vector<float> row = to_host(src, d);
            h = fnv(row.data(), d * sizeof(float), h);
        }
    return h;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s B S d V seed dist(R|Z) [dense]\n", argv[0]);
        return 2;
    }
    const size_t B = atoi(argv[1]), S = atoi(argv[2]), d = atoi(argv[3]), V = atoi(argv[4]);
    const uint64_t seed = strtoull(argv[5], nullptr, 10);
    const bool dist_r = argv[6][0] == 'R';
    const size_t W = S / PAGE_BLOCK_SIZE;

    Lcg rng(seed);
    const float wscale = dist_r ? 1.0f : 1.5f * std::sqrt(12.0f / d);
    const float shift = dist_r ? 0.0f : -0.5f;
    TensorFloat wk = random_tensor({d, d}, rng, wscale, shift);
    TensorFloat wq = random_tensor({d, d}, rng, wscale, shift);
    TensorFloat wv = random_tensor({d, d}, rng, wscale, shift);
    TensorFloat emb_table = random_tensor({V, d}, rng, dist_r ? 1.0f : 2.0f, shift);
    TensorFloat pos_table = random_tensor({S, d}, rng, dist_r ? 1.0f : 0.5f, shift);

    // prompts: ragged lengths in [0, S-2], every fifth row empty; all non-empty rows are "new"
    std::vector<int> len(B), inp(B * S, 0), new_idx(B, 0);
    int n_new = 0;
    for (size_t r = 0; r < B; ++r) {
        len[r] = (r % 5 == 3) ? 0 : 1 + static_cast<int>(rng.next() % static_cast<uint32_t>(S - 2));
        for (int j = 0; j < len[r]; ++j) inp[r * S + j] = static_cast<int>(rng.next() % static_cast<uint32_t>(EOF_TOKEN_ID));
        if (len[r] > 0) new_idx[n_new++] = static_cast<int>(r);
    }
    TensorInt lengths = int_tensor(len, {B});
    TensorInt inp_dev = int_tensor(inp, {B, S});
    TensorInt new_items = int_tensor(new_idx, {B});

    if (argc > 7 && std::string(argv[7]) == "dense") {
        // ---- the non-paged path (configs[0]): encoder.h:16-20, self_attention_inference_optimized.h:28-50,
        //      decoder.h:19-25 on the dense layouts inp_embedding[B,S,d], kt_cache[B,d,S], v_cache[B,S,d] ----
        TensorFloat inp_embedding({B, S, d}, DeviceType::DEVICE), kt_cache({B, d, S}, DeviceType::DEVICE),
            v_cache({B, S, d}, DeviceType::DEVICE);
        cudaMemset(inp_embedding.data(), 0, B * S * d * sizeof(float));
        cudaMemset(kt_cache.data(), 0, B * S * d * sizeof(float));
        cudaMemset(v_cache.data(), 0, B * S * d * sizeof(float));
        launch_inference_optimized_encoder_kernel(emb_table.data(), pos_table.data(), inp_dev.data(),
                                                  inp_embedding.data(), lengths.data(), new_items.data(),
                                                  static_cast<int>(B), static_cast<int>(S), static_cast<int>(d), n_new);
        TensorFloat q_output({B, d}, DeviceType::DEVICE), qkt({B, S}, DeviceType::DEVICE),
            attention({B, d}, DeviceType::DEVICE);
        cudaMemset(q_output.data(), 0, B * d * sizeof(float));
        cudaMemset(attention.data(), 0, B * d * sizeof(float));
        inference_self_attention(inp_embedding, lengths, wk, wq, wv, new_items, kt_cache, v_cache, q_output, qkt,
                                 attention, n_new);
        {
            const std::vector<float> e = to_host(inp_embedding.data(), B * S * d), kt = to_host(kt_cache.data(), B * S * d),
                                     v = to_host(v_cache.data(), B * S * d), q = to_host(q_output.data(), B * d);
            uint64_t he = 1469598103934665603ULL, hk = he, hv = he, hq = he;
            for (size_t r = 0; r < B; ++r) {
                he = fnv(e.data() + r * S * d, len[r] * d * sizeof(float), he);
                hv = fnv(v.data() + r * S * d, len[r] * d * sizeof(float), hv);
                for (size_t k = 0; k < d; ++k) hk = fnv(kt.data() + (r * d + k) * S, len[r] * sizeof(float), hk);
                if (len[r] > 0) hq = fnv(q.data() + r * d, d * sizeof(float), hq);
            }
            printf("HASH encoder %016llx\nHASH k_cache %016llx\nHASH v_cache %016llx\nHASH q %016llx\n",
                   (unsigned long long)he, (unsigned long long)hk, (unsigned long long)hv, (unsigned long long)hq);
            const std::vector<float> a = to_host(attention.data(), B * d);
            printf("VALS attention");
            for (size_t r = 0; r < B; ++r)
                for (size_t c = 0; c < d; c += d / 4) printf(" %.9g", a[r * d + c]);
            printf("\n");
        }
        // decoder on a GIVEN activation (the two attention blocks agree to 1e-4, not to the bit)
        TensorFloat given = random_tensor({B, d}, rng, 1.0f, shift);
        TensorFloat emb_score({B, V}, DeviceType::DEVICE);
        TensorInt decoder_result = int_tensor(std::vector<int>(B, -7), {B});
        launch_decoder(given, emb_table, emb_score, pos_table, inp_embedding, lengths, decoder_result);
        const std::vector<float> sc = to_host(emb_score.data(), B * V);
        uint64_t h = 1469598103934665603ULL;
        for (size_t r = 0; r < B; ++r)
            if (len[r] > 0) h = fnv(sc.data() + r * V, V * sizeof(float), h);
        printf("HASH logits %016llx\n", (unsigned long long)h);
        const std::vector<int> tok = to_host(decoder_result.data(), B), new_len = to_host(lengths.data(), B);
        printf("TOKENS");
        for (int t : tok) printf(" %d", t);
        printf("\nLENGTHS");
        for (int l : new_len) printf(" %d", l);
        printf("\n");
        const std::vector<float> e2 = to_host(inp_embedding.data(), B * S * d);
        uint64_t h2 = 1469598103934665603ULL;
        for (size_t r = 0; r < B; ++r)
            if (new_len[r] > 0) h2 = fnv(e2.data() + (r * S + len[r]) * d, d * sizeof(float), h2);
        printf("HASH next_embedding %016llx\n", (unsigned long long)h2);
        return 0;
    }

    // pages: every row gets a full table row, in an order that is not the slab's
    MemoryBlockManager blocks(static_cast<int>(B * W), PAGE_BLOCK_SIZE * 3 * d);
    PagedAttentionsManager manager(B, S, d);
    std::vector<std::vector<float*>> pages(B);
    for (size_t r = 0; r < B; ++r) {
        std::list<float*> got = blocks.pop_free_blocks(static_cast<int>(W));
        for (size_t k = 0; k < r % W; ++k) {   // rotate
            got.push_back(got.front());
            got.pop_front();
        }
        if (r % 2) got.reverse();
        pages[r].assign(got.begin(), got.end());
        manager.add_batch_block_pair(BatchIdMemoryBlocksPair(static_cast<int>(r), std::move(got)));
    }
    manager.maybe_flush_changes();
    TensorFloatPoint& page_table = manager.get_page_table_device();

    if (argc > 7 && std::string(argv[7]) == "cublas") {
        // ---- the warp-tiling + cuBLAS twins (paged_attention.h:46-67, decoder.h:34-37): the reference's fast build.
        //      Their summation order is cuBLAS's, so nothing here is compared by bits: VALS lines, rel 1e-4 ----
        cublasHandle_t handle;
        if (cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS) return 6;
        launch_paged_attention_encoder_kernel(emb_table.data(), pos_table.data(), inp_dev.data(), page_table.data(),
                                              lengths.data(), new_items.data(), static_cast<int>(B),
                                              static_cast<int>(S), static_cast<int>(d), n_new);
        TensorFloat q_output({B, d}, DeviceType::DEVICE), qkt({B, S}, DeviceType::DEVICE),
            attention({B, d}, DeviceType::DEVICE), latest_emb({B, d}, DeviceType::DEVICE),
            placeholder({B, d}, DeviceType::DEVICE);
        cudaMemset(q_output.data(), 0, B * d * sizeof(float));
        cudaMemset(attention.data(), 0, B * d * sizeof(float));
        cudaMemset(latest_emb.data(), 0, B * d * sizeof(float));
        paged_attention_with_cublas(page_table, lengths, wk, wq, wv, new_items, q_output, qkt, attention, latest_emb,
                                    placeholder, n_new, static_cast<int>(S), handle);
        auto sample_pages = [&](const char* name, int which) {
            printf("VALS %s", name);
            for (size_t r = 0; r < B; ++r)
                for (int j = 0; j < len[r]; j += 5) {
                    const std::vector<float> row = to_host(
                        pages[r][j / PAGE_BLOCK_SIZE] + static_cast<size_t>(j % PAGE_BLOCK_SIZE) * 3 * d + which * d, d);
                    for (size_t c = 0; c < d; c += d / 4) printf(" %.9g", row[c]);
                }
            printf("\n");
        };
        sample_pages("k_cache", 1);
        sample_pages("v_cache", 2);
        const std::vector<float> q = to_host(q_output.data(), B * d), a = to_host(attention.data(), B * d);
        printf("VALS q");
        for (size_t r = 0; r < B; ++r)
            if (len[r] > 0)
                for (size_t c = 0; c < d; c += d / 8) printf(" %.9g", q[r * d + c]);
        printf("\nVALS attention");
        for (size_t r = 0; r < B; ++r)
            for (size_t c = 0; c < d; c += d / 8) printf(" %.9g", a[r * d + c]);
        printf("\n");
        TensorFloat emb_score({B, V}, DeviceType::DEVICE);
        TensorInt decoder_result = int_tensor(std::vector<int>(B, -7), {B, 1});
        launch_paged_attention_cublas_decoder_multi_rounds(attention, emb_table, emb_score, pos_table, page_table,
                                                           lengths, decoder_result, 0, handle);
        const std::vector<float> sc = to_host(emb_score.data(), B * V);
        printf("VALS logits");
        for (size_t r = 0; r < B; ++r)
            if (len[r] > 0)
                for (size_t c = 0; c < V; c += V / 8) printf(" %.9g", sc[r * V + c]);
        printf("\n");
        const std::vector<int> tok = to_host(decoder_result.data(), B), new_len = to_host(lengths.data(), B);
        printf("TOKENS");
        for (int t : tok) printf(" %d", t);
        printf("\nLENGTHS");
        for (int l : new_len) printf(" %d", l);
        printf("\n");
        cublasDestroy(handle);
        return 0;
    }

    // ---- stage by stage (the reference's unfused chain) ----
    launch_paged_attention_encoder_kernel(emb_table.data(), pos_table.data(), inp_dev.data(), page_table.data(),
                                          lengths.data(), new_items.data(), static_cast<int>(B), static_cast<int>(S),
                                          static_cast<int>(d), n_new);
    printf("HASH encoder %016llx\n", (unsigned long long)hash_pages(pages, len, d, 0, 0));

    launch_fill_new_k_v_cache_paged_attention(page_table, new_items, lengths, wk, wv, n_new, static_cast<int>(S));
    TensorFloat q_output({B, d}, DeviceType::DEVICE);
    cudaMemset(q_output.data(), 0, B * d * sizeof(float));
    launch_get_latest_k_q_v_paged_attention(page_table, lengths, wk, wq, wv, q_output, static_cast<int>(S));
    printf("HASH k_cache %016llx\n", (unsigned long long)hash_pages(pages, len, d, 1, 0));
    printf("HASH v_cache %016llx\n", (unsigned long long)hash_pages(pages, len, d, 2, 0));
    {
        const std::vector<float> q = to_host(q_output.data(), B * d);
        uint64_t h = 1469598103934665603ULL;
        for (size_t r = 0; r < B; ++r)
            if (len[r] > 0) h = fnv(q.data() + r * d, d * sizeof(float), h);
        printf("HASH q %016llx\n", (unsigned long long)h);
    }

    TensorFloat qkt({B, S}, DeviceType::DEVICE);
    cudaMemset(qkt.data(), 0, B * S * sizeof(float));
    launch_qkt_paged_attention(q_output, page_table, lengths, qkt);
    {
        const std::vector<float> s = to_host(qkt.data(), B * S);
        uint64_t h = 1469598103934665603ULL;
        for (size_t r = 0; r < B; ++r) h = fnv(s.data() + r * S, len[r] * sizeof(float), h);   // defined up to L only
        printf("HASH qkt %016llx\n", (unsigned long long)h);
    }
    launch_softmax_in_place_with_lengths(qkt, lengths);
    {
        const std::vector<float> s = to_host(qkt.data(), B * S);
        printf("HASH softmax %016llx\n", (unsigned long long)fnv(s.data(), s.size() * sizeof(float)));
    }
    TensorFloat attention({B, d}, DeviceType::DEVICE);
    cudaMemset(attention.data(), 0, B * d * sizeof(float));
    launch_softmax_v_paged_attention(qkt, page_table, attention, lengths);
    const std::vector<float> unfused = to_host(attention.data(), B * d);
    printf("HASH softmax_v %016llx\n", (unsigned long long)fnv(unfused.data(), unfused.size() * sizeof(float)));

    // ---- the attention block in one call (paged_attention.h:17-26) on the same pages: K, V, q are rewritten with
    //      the same values, the result must agree with the chain above (the product fuses the three kernels)
    TensorFloat q2({B, d}, DeviceType::DEVICE), qkt2({B, S}, DeviceType::DEVICE), attention2({B, d}, DeviceType::DEVICE);
    cudaMemset(attention2.data(), 0, B * d * sizeof(float));
    paged_attention(page_table, lengths, wk, wq, wv, new_items, q2, qkt2, attention2, n_new, static_cast<int>(S));
    {
        const std::vector<float> a = to_host(attention2.data(), B * d);
        printf("VALS paged_attention");
        for (size_t r = 0; r < B; ++r)
            for (size_t c = 0; c < d; c += d / 4) printf(" %.9g", a[r * d + c]);
        printf("\n");
        printf("VALS chain");
        for (size_t r = 0; r < B; ++r)
            for (size_t c = 0; c < d; c += d / 4) printf(" %.9g", unfused[r * d + c]);
        printf("\n");
    }

    // ---- decoder on the chain's result: logits, tokens, lengths, next embedding (decoder.h:27-31) ----
    TensorFloat emb_score({B, V}, DeviceType::DEVICE);
    TensorInt decoder_result = int_tensor(std::vector<int>(B, -7), {B, 1});
    launch_paged_attention_decoder_multi_rounds(attention, emb_table, emb_score, pos_table, page_table, lengths,
                                                decoder_result, 0);
    {
        const std::vector<float> s = to_host(emb_score.data(), B * V);
        uint64_t h = 1469598103934665603ULL;
        for (size_t r = 0; r < B; ++r)
            if (len[r] > 0) h = fnv(s.data() + r * V, V * sizeof(float), h);
        printf("HASH logits %016llx\n", (unsigned long long)h);
        const std::vector<int> tok = to_host(decoder_result.data(), B);
        const std::vector<int> new_len = to_host(lengths.data(), B);
        printf("TOKENS");
        for (int t : tok) printf(" %d", t);
        printf("\nLENGTHS");
        for (int l : new_len) printf(" %d", l);
        printf("\n");
        // the embedding of the new token sits at position L of every row that goes on
        std::vector<int> went_on(B);
        for (size_t r = 0; r < B; ++r) went_on[r] = (new_len[r] > 0) ? len[r] : -1;
        uint64_t h2 = 1469598103934665603ULL;
        for (size_t r = 0; r < B; ++r)
            if (went_on[r] >= 0) {
                const int j = went_on[r];
                const std::vector<float> row =
                    to_host(pages[r][j / PAGE_BLOCK_SIZE] + static_cast<size_t>(j % PAGE_BLOCK_SIZE) * 3 * d, d);
                h2 = fnv(row.data(), d * sizeof(float), h2);
            }
        printf("HASH next_embedding %016llx\n", (unsigned long long)h2);
    }
    return 0;
}
