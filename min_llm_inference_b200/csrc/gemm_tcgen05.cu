// gemm_tcgen05.cu -- placeholder until the tcgen05 path lands (see next commit)
#include "common.cuh"
#include "kernels.h"
namespace mli {
bool tcgen05_supported(mli_ctx*) { return false; }
int launch_prefill_kv_paged_tc(mli_ctx*, float* const*, const TileDesc*, const int*, int, const int*,
                               const float*, const float*, int, int) {
    set_error("tcgen05 path not built");
    return MLI_ERR_UNSUPPORTED;
}
int launch_qkv_latest_paged_tc(mli_ctx*, float* const*, const int*, const float*, const float*,
                               const float*, float*, int, int, int) {
    set_error("tcgen05 path not built");
    return MLI_ERR_UNSUPPORTED;
}
int launch_logits_tc(mli_ctx*, const float*, const float*, float*, int, int, int) {
    set_error("tcgen05 path not built");
    return MLI_ERR_UNSUPPORTED;
}
}  // namespace mli
