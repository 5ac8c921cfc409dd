"""B200-native (sm_100a) paged-attention decode path for min_llm_inference.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/mli_b200.h) -> libmli_b200.so
  host/      C++ mirror of the reference's classes on top of the C ABI (drop-in headers)
  capi.py    ctypes binding used by tests/ and bench.py (PyTorch only supplies device memory)
"""
from .capi import (Context, Engine, Comm, EngineCfg, EngineStats, MliError, build_library, load_library,
                   LIB_PATH, OPT_GEMM_MODE, OPT_ATTN_CHUNK_PAGES, OPT_ATTN_CTAS_PER_SM, OPT_PDL, OPT_KV_FORMAT,
                   OPT_ATTN_KERNEL, OPT_ATTN_MIN_DYN,
                   GEMM_TCGEN05, GEMM_SIMT_EXACT, PAGE_BLOCK_SIZE, EOF_TOKEN_ID,
                   EMPTY_ROW_TOKEN_ID, DEFAULT_INIT_NUM_BLOCKS)

__all__ = [n for n in dir() if not n.startswith("_")]
