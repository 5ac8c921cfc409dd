"""CPU-only: the C++ mirror of the reference API builds and exports the reference's entry points."""
import subprocess
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
LIB = REPO / "min_llm_inference_b200" / "libmli_b200_host.so"

WANTED = [
    "start_inference_engine(", "start_paged_attention_inference_engine(",
    "start_paged_attention_cublas_inference_engine(", "InferenceModel::forward(",
    "PagedAttentionInferenceModel::forward(", "PagedAttentionCublasInferenceModel::forward(",
    "PagedAttentionLayer::forward(", "PagedEncoderLayer::forward(", "PagedDecoderLayer::forward(",
    "MemoryBlockManager::pop_free_blocks", "MemoryBlockManager::return_free_blocks(",
    "PagedAttentionsManager::add_batch_block_pair(", "PagedAttentionsManager::get_page_table_device(",
    "allocate_or_free_memory_blocks_if_needed(", "process_decoder_result(", "is_done(",
    "insert_new_items(", "paged_attention(", "paged_attention_with_cublas(",
    "launch_get_latest_k_q_v_paged_attention(", "launch_fill_new_k_v_cache_paged_attention(",
    "launch_paged_attention_decoder_multi_rounds(", "launch_paged_attention_encoder_kernel(",
    "inference_self_attention(", "launch_decoder(", "get_global_throughput_counter(", "cuda_check(",
]


def test_host_mirror_exports_reference_api():
    subprocess.run(["make", "-C", str(REPO / "min_llm_inference_b200" / "csrc"), "-j8"], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", str(REPO / "min_llm_inference_b200" / "host"), "-j8"], check=True,
                   stdout=subprocess.DEVNULL)
    syms = subprocess.run(["nm", "-D", "--demangle", "--defined-only", str(LIB)], check=True,
                          capture_output=True, text=True).stdout
    missing = [w for w in WANTED if w not in syms]
    assert not missing, f"reference entry points missing from the host mirror: {missing}"


def test_drop_in_driver_compiles_against_the_mirror():
    subprocess.run(["make", "-C", str(REPO / "tests" / "dropin"), "mli"], check=True,
                   stdout=subprocess.DEVNULL)
    for name in ("dropin_driver_mli", "dropin_stepwise_mli", "dropin_kernels_mli"):
        assert (REPO / "tests" / "dropin" / "_build" / name).exists(), name
