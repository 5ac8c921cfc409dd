"""ctypes binding of the C ABI in include/mli_b200.h (libmli_b200.so).

PyTorch is used by callers only for device memory and streams; every compute call goes through the
C ABI into hand-written sm_100a kernels.  There is no CPU or PyTorch fallback: if the library is
missing or no CUDA device is present, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_PATH = Path(os.environ.get("MLI_B200_LIB", PKG_DIR / "libmli_b200.so"))   # override: A/B builds

OPT_GEMM_MODE = 1
OPT_ATTN_CHUNK_PAGES = 2
OPT_ATTN_CTAS_PER_SM = 3
OPT_PDL = 4
OPT_KV_FORMAT = 5
OPT_ATTN_KERNEL = 6
OPT_ATTN_MIN_DYN = 7
GEMM_TCGEN05 = 0
GEMM_SIMT_EXACT = 1

PAGE_BLOCK_SIZE = 16
EOF_TOKEN_ID = 1023
EMPTY_ROW_TOKEN_ID = -1
DEFAULT_INIT_NUM_BLOCKS = 4

_P, _I, _LL = C.c_void_p, C.c_int, C.c_longlong


class EngineCfg(C.Structure):
    _fields_ = [
        ("n_batch", _I), ("n_sequence", _I), ("emb_dim", _I), ("n_vocab", _I),
        ("n_blocks", _I), ("n_forward_rounds", _I), ("compat_stale_lengths", _I),
        ("max_requests", _I), ("page_pool", _P),
        ("max_new_tokens", _I), ("max_prefill_positions", _I), ("prefill_chunk_positions", _I),
    ]


class EngineStats(C.Structure):
    _fields_ = [
        ("steps", _LL), ("generated_tokens", _LL), ("preemptions", _LL), ("admitted", _LL),
        ("n_finished", _I), ("gpu_ms", C.c_float), ("attn_ms", C.c_float),
        ("attn_bytes", C.c_double), ("attn_launches", _LL),
        ("gemm_ms", C.c_float), ("gemm_flops", C.c_double), ("gemm_launches", _LL),
        ("gemm_max_flops", C.c_double), ("gemm_max_ms", C.c_float),
        ("peak_resident_rows", _I), ("min_free_pages", _I),
    ]


# name -> (restype, argtypes); mirrors include/mli_b200.h one to one
SIGNATURES = {
    "mli_ctx_create": (_I, [C.POINTER(_P), _I, _P]),
    "mli_ctx_destroy": (_I, [_P]),
    "mli_ctx_set_stream": (_I, [_P, _P]),
    "mli_ctx_set_option": (_I, [_P, _I, _I]),
    "mli_ctx_get_option": (_I, [_P, _I, C.POINTER(_I)]),
    "mli_ctx_synchronize": (_I, [_P]),
    "mli_ctx_register_weights": (_I, [_P, _P, _P, _P, _P, _I, _I]),
    "mli_ctx_unregister_weights": (_I, [_P]),
    "mli_last_error": (C.c_char_p, []),
    "mli_version": (C.c_char_p, []),
    "mli_kernel_launch_count": (_LL, []),
    "mli_debug_set_gemm_stamps": (_I, [_P, _P]),
    "mli_debug_last_gemm_plan": (_I, [_P, _I, _P]),
    "mli_debug_set_step_trace": (_I, [_P, _P]),
    "mli_paged_encoder": (_I, [_P] * 7 + [_I] * 4),
    "mli_prefill_kv_paged": (_I, [_P] * 6 + [_I] * 4),
    "mli_qkv_latest_paged": (_I, [_P] * 7 + [_I] * 3),
    "mli_decode_attention_paged": (_I, [_P] * 6 + [_I] * 3),
    "mli_qkt_paged": (_I, [_P] * 5 + [_I] * 3),
    "mli_softmax_in_place_with_lengths": (_I, [_P] * 3 + [_I] * 2),
    "mli_softmax_v_paged": (_I, [_P] * 5 + [_I] * 3),
    "mli_paged_attention": (_I, [_P] * 10 + [_I] * 4),
    "mli_paged_decoder": (_I, [_P] * 8 + [_I] * 6),
    "mli_paged_forward": (_I, [_P] * 5 + [_I] + [_P] * 8 + [_I] * 5),
    "mli_dense_encoder": (_I, [_P] * 7 + [_I] * 4),
    "mli_self_attention": (_I, [_P] * 12 + [_I] * 5),
    "mli_dense_decoder": (_I, [_P] * 8 + [_I] * 4),
    "mli_dense_forward": (_I, [_P] * 5 + [_I] + [_P] * 10 + [_I] * 4),
    "mli_engine_create": (_I, [_P, C.POINTER(EngineCfg)] + [_P] * 5 + [C.POINTER(_P)]),
    "mli_engine_destroy": (_I, [_P]),
    "mli_engine_submit": (_I, [_P, _I, _P, _P, _I]),
    "mli_engine_enqueue": (_I, [_P, _I, _P, _P, _I, C.POINTER(_I)]),
    "mli_engine_poll_finished": (_I, [_P, _I, _P, _P, _P, _LL, C.POINTER(_I)]),
    "mli_engine_run": (_I, [_P, _LL, _I]),
    "mli_engine_results": (_I, [_P, _P, _P, _P, C.POINTER(_I)]),
    "mli_engine_copy_tokens": (_I, [_P, _P, _P]),
    "mli_engine_get_stats": (_I, [_P, C.POINTER(EngineStats)]),
    "mli_comm_get_unique_id": (_I, [_P, C.c_size_t]),
    "mli_comm_init_rank": (_I, [_P, _I, _I, _P, C.POINTER(_P)]),
    "mli_comm_init_all": (_I, [C.POINTER(_P), _I, C.POINTER(_P)]),
    "mli_comm_group_start": (_I, []),
    "mli_comm_group_end": (_I, []),
    "mli_comm_gather_tokens": (_I, [_P, _P, _I, _P, _P]),
    "mli_comm_info": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "mli_comm_destroy": (_I, [_P]),
}

COMM_ID_BYTES = 128


class MliError(RuntimeError):
    pass


def build_library(force: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into libmli_b200.so (nvcc cross-compiles without a GPU)."""
    if force or not LIB_PATH.exists():
        subprocess.run(["make", "-C", str(PKG_DIR / "csrc"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def load_library() -> C.CDLL:
    """dlopen the product library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise MliError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU / PyTorch fallback)")
        lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(x):
    """device/host pointer of a torch tensor, numpy array, int or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


class Context:
    """RAII wrapper of mli_ctx; methods mirror the C entry points (same argument order)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = _P()
        rc = self.lib.mli_ctx_create(C.byref(h), device, stream)
        if rc != 0:
            raise MliError(self.lib.mli_last_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mli_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise MliError(f"mli error {rc}: {self.lib.mli_last_error().decode()}")

    def call(self, name, *args):
        fn = getattr(self.lib, name)
        conv = []
        for a, t in zip(args, fn.argtypes[1:]):
            conv.append(_ptr(a) if t is _P else a)
        self._check(fn(self.h, *conv))

    def set_option(self, opt, value):
        self._check(self.lib.mli_ctx_set_option(self.h, opt, value))

    def get_option(self, opt) -> int:
        v = _I()
        self._check(self.lib.mli_ctx_get_option(self.h, opt, C.byref(v)))
        return v.value

    def synchronize(self):
        self._check(self.lib.mli_ctx_synchronize(self.h))

    def register_weights(self, wk, wq, wv, emb, emb_dim, n_vocab):
        self._check(self.lib.mli_ctx_register_weights(self.h, _ptr(wk), _ptr(wq), _ptr(wv), _ptr(emb),
                                                      emb_dim, n_vocab))

    def unregister_weights(self):
        self._check(self.lib.mli_ctx_unregister_weights(self.h))

    def set_stream(self, stream):
        self._check(self.lib.mli_ctx_set_stream(self.h, stream))

    def launch_count(self) -> int:
        return self.lib.mli_kernel_launch_count()


class Comm:
    """NCCL communicator behind the C ABI (mli_comm_*): the final token gather of a request-sharded job."""

    def __init__(self, ctx: "Context", world: int, rank: int, unique_id: bytes):
        self.ctx, self.lib = ctx, ctx.lib
        self.world, self.rank = world, rank
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        h = _P()
        ctx._check(self.lib.mli_comm_init_rank(ctx.h, world, rank, buf, C.byref(h)))
        self.h = h

    @staticmethod
    def unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(COMM_ID_BYTES)
        if lib.mli_comm_get_unique_id(buf, COMM_ID_BYTES) != 0:
            raise MliError(lib.mli_last_error().decode())
        return buf.raw

    def gather_tokens(self, engine: "Engine", per_rank: int, all_tokens_dev, all_counts_dev):
        self.ctx._check(self.lib.mli_comm_gather_tokens(self.h, engine.h, per_rank, _ptr(all_tokens_dev),
                                                        _ptr(all_counts_dev)))

    def nccl_version(self) -> int:
        v = _I()
        self.ctx._check(self.lib.mli_comm_info(self.h, None, None, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.mli_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """On-device continuous-batching engine (mli_engine_*)."""

    def __init__(self, ctx: Context, cfg: EngineCfg, emb, pos, wk, wq, wv):
        self.ctx = ctx
        self.lib = ctx.lib
        self.cfg = cfg
        self._keep = (emb, pos, wk, wq, wv)
        h = _P()
        ctx._check(self.lib.mli_engine_create(ctx.h, C.byref(cfg), _ptr(emb), _ptr(pos), _ptr(wk),
                                              _ptr(wq), _ptr(wv), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mli_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, prompt_offsets, prompt_tokens, is_device=False):
        n_req = len(prompt_offsets) - 1
        self.n_req = n_req
        self.ctx._check(self.lib.mli_engine_submit(self.h, n_req, _ptr(prompt_offsets),
                                                   _ptr(prompt_tokens), 1 if is_device else 0))

    def enqueue(self, prompt_offsets, prompt_tokens, is_device=False) -> int:
        """append requests to the live engine (no reset); returns the id of the first new request"""
        n_req = len(prompt_offsets) - 1
        first = _I()
        self.ctx._check(self.lib.mli_engine_enqueue(self.h, n_req, _ptr(prompt_offsets),
                                                    _ptr(prompt_tokens), 1 if is_device else 0,
                                                    C.byref(first)))
        self.n_req = getattr(self, "n_req", 0) + n_req
        return first.value

    def poll_finished(self, max_out=None):
        """requests finished since the last poll (non-blocking): ({id: tokens}, ids in finish order)"""
        import numpy as np
        n, S = (max_out or max(self.n_req, 1)), self.cfg.n_sequence
        ids = np.zeros(n, np.int32)
        offs = np.zeros(n + 1, np.int32)
        toks = np.zeros(n * S, np.int32)
        k = _I()
        self.ctx._check(self.lib.mli_engine_poll_finished(self.h, n, ids.ctypes.data, offs.ctypes.data,
                                                          toks.ctypes.data, n * S, C.byref(k)))
        k = k.value
        return {int(ids[i]): toks[offs[i]:offs[i + 1]].copy() for i in range(k)}, ids[:k].copy()

    def run(self, max_steps=0, profile_attention=False):
        self.ctx._check(self.lib.mli_engine_run(self.h, max_steps, 1 if profile_attention else 0))

    def results(self):
        import numpy as np
        n, S = self.n_req, self.cfg.n_sequence
        ids = np.zeros(n, np.int32)
        offs = np.zeros(n + 1, np.int32)
        toks = np.empty(n * S, np.int32)     # only [0, offs[n_finished]) is written and read
        nf = _I()
        self.ctx._check(self.lib.mli_engine_results(self.h, ids.ctypes.data, offs.ctypes.data,
                                                    toks.ctypes.data, C.byref(nf)))
        k = nf.value
        return {int(ids[i]): toks[offs[i]:offs[i + 1]].copy() for i in range(k)}, ids[:k].copy()

    def copy_tokens(self, tokens_dev, counts_dev):
        self.ctx._check(self.lib.mli_engine_copy_tokens(self.h, _ptr(tokens_dev), _ptr(counts_dev)))

    def stats(self) -> EngineStats:
        st = EngineStats()
        self.ctx._check(self.lib.mli_engine_get_stats(self.h, C.byref(st)))
        return st
