// comm.cu -- the one collective of the path behind the C ABI (mli_comm_*): the final token gather.
//
// Requests shard by GPU: every rank runs a full engine on its own requests, pages and
// scheduler; nothing is exchanged during a decode step.  When a job is over the per-rank request
// tables (token lists + counts) are all-gathered with NCCL over NVLink / NVSwitch so that every rank
// (and the caller of the reference's start_paged_attention_inference_engine, src/inferencer.cpp:43-85)
// holds the whole job's finished token lists.
//
// NCCL is loaded with dlopen at the first mli_comm_* call, so libmli_b200.so itself has no NCCL
// link dependency (single-GPU users never need it) and, inside a PyTorch process, the libnccl.so.2
// torch already mapped is the one that gets used.  <nccl.h> supplies the types only.
#include "common.cuh"
#include "kernels.h"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

struct mli_comm {
    mli_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    cudaEvent_t ev = nullptr;
};

namespace mli {
namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.handle, "nccl" #name))
        SYM(GetUniqueId); SYM(CommInitRank); SYM(CommInitAll); SYM(CommDestroy); SYM(AllGather);
        SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString); SYM(GetVersion);
#undef SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    if (!api.ok) {
        set_error("mli_comm: libnccl.so.2 could not be loaded (NCCL is only needed for multi-GPU jobs)");
        return nullptr;
    }
    return &api;
}

int nccl_fail(NcclApi* api, ncclResult_t r, int line) {
    char buf[256];
    snprintf(buf, sizeof(buf), "[NCCL ERROR] comm.cu:%d: %s", line, api->GetErrorString(r));
    set_error(buf);
    return MLI_ERR_CUDA;
}
#define MLI_NCCL(api, expr)                                      \
    do {                                                         \
        ncclResult_t _r = (expr);                                \
        if (_r != ncclSuccess) return nccl_fail(api, _r, __LINE__); \
    } while (0)

}  // namespace
}  // namespace mli

using namespace mli;

extern "C" {

int mli_comm_get_unique_id(void* id_out, size_t id_bytes) {
    MLI_REQUIRE(id_out && id_bytes >= sizeof(ncclUniqueId), "unique id buffer must hold MLI_COMM_ID_BYTES");
    NcclApi* api = nccl_api();
    if (!api) return MLI_ERR_UNSUPPORTED;
    ncclUniqueId id;
    MLI_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return MLI_OK;
}

int mli_comm_init_rank(mli_ctx* ctx, int world_size, int rank, const void* unique_id, mli_comm** out) {
    MLI_ENTER(ctx, "null ctx");
    MLI_REQUIRE(unique_id && out && world_size >= 1 && rank >= 0 && rank < world_size, "bad communicator arguments");
    NcclApi* api = nccl_api();
    if (!api) return MLI_ERR_UNSUPPORTED;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    mli_comm* c = new mli_comm();
    c->ctx = ctx;
    c->world = world_size;
    c->rank = rank;
    ncclResult_t r = api->CommInitRank(&c->comm, world_size, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return nccl_fail(api, r, __LINE__);
    }
    cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming);
    *out = c;
    return MLI_OK;
}

int mli_comm_init_all(mli_ctx* const* ctxs, int n, mli_comm** comms_out) {
    MLI_REQUIRE(ctxs && comms_out && n >= 1 && n <= 64, "bad communicator arguments");
    NcclApi* api = nccl_api();
    if (!api) return MLI_ERR_UNSUPPORTED;
    int devs[64];
    ncclComm_t comms[64];
    for (int i = 0; i < n; ++i) {
        MLI_REQUIRE(ctxs[i] != nullptr, "null ctx");
        devs[i] = ctxs[i]->device;
    }
    MLI_NCCL(api, api->CommInitAll(comms, n, devs));
    for (int i = 0; i < n; ++i) {
        mli_comm* c = new mli_comm();
        c->ctx = ctxs[i];
        c->comm = comms[i];
        c->world = n;
        c->rank = i;
        cudaSetDevice(devs[i]);
        cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming);
        comms_out[i] = c;
    }
    return MLI_OK;
}

int mli_comm_group_start(void) {
    NcclApi* api = nccl_api();
    if (!api) return MLI_ERR_UNSUPPORTED;
    MLI_NCCL(api, api->GroupStart());
    return MLI_OK;
}

int mli_comm_group_end(void) {
    NcclApi* api = nccl_api();
    if (!api) return MLI_ERR_UNSUPPORTED;
    MLI_NCCL(api, api->GroupEnd());
    return MLI_OK;
}

int mli_comm_gather_tokens(mli_comm* c, mli_engine* e, int per_rank, int* all_tokens_dev, int* all_counts_dev) {
    MLI_REQUIRE(c && e && all_tokens_dev && all_counts_dev, "null argument");
    MLI_ENTER(c->ctx, "null ctx");
    NcclApi* api = nccl_api();
    if (!api) return MLI_ERR_UNSUPPORTED;
    const int *tok = nullptr, *cnt = nullptr;
    int cap = 0, S = 0;
    cudaStream_t es = nullptr;
    mli_ctx* ectx = nullptr;
    int rc = engine_token_table(e, &tok, &cnt, &cap, &S, &es, &ectx);
    if (rc) return rc;
    MLI_REQUIRE(ectx == c->ctx, "engine and communicator belong to different contexts");
    MLI_REQUIRE(per_rank >= 1 && per_rank <= cap,
                "per_rank must not exceed the engine's max_requests (the request table is sent as is)");
    // the table is final once the engine's stream has drained; the gather itself runs on the
    // context's stream (stream-ordered, no host synchronisation)
    MLI_CUDA(cudaEventRecord(c->ev, es));
    MLI_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->ev, 0));
    // (inside a caller's mli_comm_group_start/_end these nest; otherwise they form their own group)
    MLI_NCCL(api, api->GroupStart());
    ncclResult_t r1 = api->AllGather(tok, all_tokens_dev, (size_t)per_rank * S, ncclInt32, c->comm, c->ctx->stream);
    ncclResult_t r2 = api->AllGather(cnt, all_counts_dev, (size_t)per_rank, ncclInt32, c->comm, c->ctx->stream);
    ncclResult_t r3 = api->GroupEnd();
    if (r1 != ncclSuccess) return nccl_fail(api, r1, __LINE__);
    if (r2 != ncclSuccess) return nccl_fail(api, r2, __LINE__);
    if (r3 != ncclSuccess) return nccl_fail(api, r3, __LINE__);
    return MLI_OK;
}

int mli_comm_info(mli_comm* c, int* world_size, int* rank, int* nccl_version) {
    MLI_REQUIRE(c, "null communicator");
    if (world_size) *world_size = c->world;
    if (rank) *rank = c->rank;
    if (nccl_version) {
        NcclApi* api = nccl_api();
        int v = 0;
        if (api && api->GetVersion) api->GetVersion(&v);
        *nccl_version = v;
    }
    return MLI_OK;
}

int mli_comm_destroy(mli_comm* c) {
    if (!c) return MLI_OK;
    NcclApi* api = nccl_api();
    if (c->ctx) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
    }
    if (api && c->comm) api->CommDestroy(c->comm);
    if (c->ev) cudaEventDestroy(c->ev);
    delete c;
    return MLI_OK;
}

}  // extern "C"
