"""CPU-only: the C-ABI library builds, loads and exports every symbol include/mli_b200.h declares;
compute entry points fail loudly without a GPU (there is no CPU fallback in the product path)."""
import ctypes as C
import re
from pathlib import Path

import pytest

import min_llm_inference_b200 as mli
from min_llm_inference_b200 import capi

REPO = Path(__file__).resolve().parent.parent
HEADER = (REPO / "include" / "mli_b200.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"^(?:int|const char\*|long long)\s+(mli_\w+)\s*\(", HEADER, re.M)))


def test_library_builds_and_exports_every_declared_symbol():
    capi.build_library()
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mli_b200.h but not exported"
    assert set(names) == set(capi.SIGNATURES), "capi.py and the header disagree"


def test_header_cites_the_reference_interface_for_every_stage():
    for name in ("mli_paged_encoder", "mli_prefill_kv_paged", "mli_qkv_latest_paged",
                 "mli_decode_attention_paged", "mli_paged_attention", "mli_paged_decoder",
                 "mli_paged_forward", "mli_self_attention", "mli_dense_forward"):
        decl = HEADER.index(name + "(")
        comment = HEADER[HEADER.rindex("/*", 0, decl):decl]
        assert re.search(r"\.(cu|cpp|h):\d+", comment), f"{name}: no reference file:line cited"
    engine = HEADER[HEADER.index("on-device continuous-batching engine"):HEADER.index("} mli_engine_cfg;")]
    assert "src/inferencer.cpp:43-133" in engine and "src/paged_item_storage.cpp:14-203" in engine


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mli.MliError, match="no CUDA device"):
        mli.Context(0)


def test_product_sources_never_touch_the_oracle():
    """the oracle is test infrastructure: nothing under the package or include/ may name it"""
    for path in list((REPO / "min_llm_inference_b200").rglob("*")) + list((REPO / "include").rglob("*")):
        if path.is_file() and path.suffix in {".cu", ".cuh", ".h", ".hpp", ".cpp", ".py"}:
            txt = path.read_text()
            assert "oracle/" not in txt and "liboracle" not in txt and "libmli_ref" not in txt, path


def test_option_constants_match_the_header():
    """capi.py's OPT_* values are the header's MLI_OPT_* enum"""
    enum = dict((k, int(v)) for k, v in re.findall(r"\b(MLI_OPT_\w+)\s*=\s*(\d+)", HEADER))
    assert len(enum) >= 7
    for name, value in enum.items():
        py = "OPT_" + name[len("MLI_OPT_"):]
        assert getattr(capi, py) == value, f"{py} != {name}"
        assert getattr(mli, py) == value, f"{py} is not exported by the package"


def test_comm_and_engine_argument_errors_need_no_gpu():
    """contract violations are reported before anything touches a device; the multi-GPU entry points cite why
    the reference has nothing for them to replace"""
    lib = capi.load_library()
    small = C.create_string_buffer(8)
    assert lib.mli_comm_get_unique_id(small, 8) == -2          # MLI_ERR_ARG: buffer shorter than MLI_COMM_ID_BYTES
    assert b"MLI_COMM_ID_BYTES" in lib.mli_last_error()
    out = C.c_void_p()
    assert lib.mli_comm_init_rank(None, 2, 0, small, C.byref(out)) == -2      # null context
    assert lib.mli_comm_gather_tokens(None, None, 1, None, None) == -2
    assert lib.mli_engine_submit(None, 0, None, None, 0) == -2
    assert lib.mli_engine_enqueue(None, 0, None, None, 0, None) == -2
    n = C.c_int()
    assert lib.mli_engine_poll_finished(None, 0, None, None, None, 0, C.byref(n)) == -2
    comm = HEADER[HEADER.index("multi-GPU: request sharding"):HEADER.index("#define MLI_COMM_ID_BYTES")]
    assert "include/inferencer.h:23-32" in comm and "single-GPU" in comm
    for name in ("mli_engine_enqueue", "mli_engine_poll_finished"):
        decl = HEADER.index("int " + name + "(")
        comment = HEADER[HEADER.rindex("/*", 0, decl):decl]
        assert re.search(r"\.(cu|cpp|h):\d+", comment), f"{name}: no reference file:line cited"
