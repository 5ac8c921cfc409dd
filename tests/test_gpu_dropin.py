"""Drop-in proof: tests/dropin/engine_driver.cpp is ONE source written against the reference's
public C++ API; it is compiled against the reference tree (oracle/_ref/dropin_driver_ref) and against
this repo's host mirror (tests/dropin/_build/dropin_driver_mli).  Both binaries must print the same
finished token lists, in the same finish order, for the non-paged engine (C1 shape) and for the two
paged engines (where our default replays the reference's stale-length quirk)."""
import os
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
MLI = REPO / "tests" / "dropin" / "_build" / "dropin_driver_mli"
REF = REPO / "oracle" / "_ref" / "dropin_driver_ref"


def run(binary, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([str(binary)] + [str(a) for a in args], capture_output=True, text=True, env=e,
                         timeout=600)
    assert out.returncode == 0, f"{binary.name} failed: {out.stderr[-2000:]}"
    return [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]


@pytest.fixture(scope="module")
def binaries(torch_cuda):
    if not MLI.exists() or not REF.exists():
        pytest.skip("drop-in drivers not built (python -c 'import __graft_entry__ as g; g.build()')")
    return MLI, REF


# kind, B, S, d, V, n_blocks, n_req, lo, hi, seed
CASES = [
    ("dense", 32, 256, 256, 1024, 0, 72, 1, 128, 5),          # BASELINE config C1
    ("paged", 16, 128, 256, 1024, 64, 40, 1, 64, 6),
    ("paged_cublas", 16, 128, 256, 1024, 64, 40, 1, 64, 7),
    ("paged", 8, 128, 128, 1024, 36, 24, 20, 64, 8),          # pool pressure: pre-emption
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-B{c[1]}-d{c[3]}")
@pytest.mark.parametrize("dist", ["R", "Z"])
@pytest.mark.parametrize("gemm_mode", ["1", "0"], ids=["simt_exact", "tcgen05"])
def test_same_source_same_tokens(binaries, case, dist, gemm_mode):
    mli, ref = binaries
    args = list(case) + [dist]
    theirs = run(ref, args)
    mine = run(mli, args, env={"MLI_GEMM_MODE": gemm_mode})
    assert len(theirs) == case[6]
    assert mine == theirs, "finished token lists differ between the reference build and ours"


def test_corrected_lengths_switch(binaries):
    """MLI_FIX_STALE_LENGTHS=1: the paged engine decodes properly instead of replaying quirk Q1, and
    then agrees with the reference's NON-paged engine on the same requests"""
    mli, ref = binaries
    paged = ["paged", 8, 128, 128, 1024, 64, 12, 5, 40, 9, "Z"]
    dense = ["dense", 8, 128, 128, 1024, 0, 12, 5, 40, 9, "Z"]
    fixed = run(mli, paged, env={"MLI_FIX_STALE_LENGTHS": "1", "MLI_GEMM_MODE": "1"})
    want = run(ref, dense)
    assert sorted(fixed) == sorted(want)
    quirk = run(mli, paged, env={"MLI_GEMM_MODE": "1"})
    assert sorted(quirk) != sorted(want), "with 12 requests on 8 rows the quirk must show"


def test_chunked_prefill_through_the_reference_entry_points(binaries):
    """MLI_PREFILL_CHUNK: the drop-in's paged engine prefills prompts in chunks (device engine policy); the token
    list of every request is the one the unchunked run produces (and the reference's non-paged engine)"""
    mli, ref = binaries
    paged = ["paged", 8, 128, 128, 1024, 64, 14, 30, 100, 11, "Z"]
    env = {"MLI_FIX_STALE_LENGTHS": "1", "MLI_GEMM_MODE": "1"}
    plain = run(mli, paged, env=env)
    chunked = run(mli, paged, env=dict(env, MLI_PREFILL_CHUNK="32"))
    assert sorted(plain) == sorted(chunked)
    want = run(ref, ["dense", 8, 128, 128, 1024, 0, 14, 30, 100, 11, "Z"])
    assert sorted(chunked) == sorted(want)


# ---- kernel level: tests/dropin/kernels_driver.cpp (the reference's launchers, stage by stage) ----------------
KMLI = REPO / "tests" / "dropin" / "_build" / "dropin_kernels_mli"
KREF = REPO / "oracle" / "_ref" / "dropin_kernels_ref"


def run_kernels(binary, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([str(binary)] + [str(a) for a in args], capture_output=True, text=True, env=e,
                         timeout=600)
    assert out.returncode == 0, f"{binary.name} failed: {out.stderr[-2000:]}"
    res = {}
    for line in out.stdout.splitlines():
        f = line.split()
        if f and f[0] == "HASH":
            res[f[1]] = f[2]
        elif f and f[0] == "VALS":
            res[f[1]] = [float(x) for x in f[2:]]
        elif f and f[0] in ("TOKENS", "LENGTHS"):
            res[f[0]] = [int(x) for x in f[1:]]
    return res


# B, S, d, V, seed
KERNEL_CASES = [(12, 64, 128, 1024, 3), (20, 128, 256, 1024, 4), (7, 256, 512, 1024, 5)]


@pytest.mark.parametrize("case", KERNEL_CASES, ids=lambda c: f"B{c[0]}-S{c[1]}-d{c[2]}")
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_same_source_same_kernel_results(torch_cuda, case, dist):
    """the reference's kernel launchers called one by one from ONE source (encoder, prefill, latest QKV, qkt,
    masked softmax, softmax.V, decoder; then paged_attention() as a block): compiled against the reference and
    against this repo, exact-order mode.  Every stage of the chain is BIT-identical (hashes of the float bits);
    the fused block agrees with the reference's block at rel 1e-4; tokens and lengths are equal"""
    if not KMLI.exists() or not KREF.exists():
        pytest.skip("kernel-level drop-in drivers not built")
    import numpy as np
    args = list(case) + [dist]
    theirs = run_kernels(KREF, args)
    mine = run_kernels(KMLI, args, env={"MLI_GEMM_MODE": "1"})
    stages = ["encoder", "k_cache", "v_cache", "q", "qkt", "softmax", "softmax_v", "logits", "next_embedding"]
    assert set(stages) <= set(theirs), theirs.keys()
    for s in stages:
        assert mine[s] == theirs[s], f"stage {s}: bits differ from the reference's kernel"
    assert mine["TOKENS"] == theirs["TOKENS"] and mine["LENGTHS"] == theirs["LENGTHS"]
    assert any(t >= 0 for t in theirs["TOKENS"]) and any(t == -1 for t in theirs["TOKENS"])   # empty rows emit -1
    a, b = np.asarray(mine["paged_attention"]), np.asarray(theirs["paged_attention"])
    scale = np.abs(b).max()
    assert np.abs(a - b).max() <= 1e-4 * scale, "paged_attention(): fused block differs from the reference's block"
    # tensor-core mode: tokens equal on these seeds, block within 1e-4
    tc = run_kernels(KMLI, args, env={"MLI_GEMM_MODE": "0"})
    assert np.abs(np.asarray(tc["paged_attention"]) - b).max() <= 1e-4 * scale
    assert tc["LENGTHS"] == theirs["LENGTHS"]


@pytest.mark.parametrize("case", [(32, 256, 256, 1024, 6), (9, 64, 128, 1024, 7)], ids=lambda c: f"B{c[0]}-S{c[1]}-d{c[2]}")
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_same_source_same_dense_kernel_results(torch_cuda, case, dist):
    """the non-paged launchers (configs[0] shape first): dense encoder, inference_self_attention, launch_decoder --
    embeddings, K^T / V caches, q, logits, next embedding bit-identical; attention at rel 1e-4; tokens, lengths equal"""
    if not KMLI.exists() or not KREF.exists():
        pytest.skip("kernel-level drop-in drivers not built")
    import numpy as np
    args = list(case) + [dist, "dense"]
    theirs = run_kernels(KREF, args)
    mine = run_kernels(KMLI, args, env={"MLI_GEMM_MODE": "1"})
    for s in ["encoder", "k_cache", "v_cache", "q", "logits", "next_embedding"]:
        assert mine[s] == theirs[s], f"dense stage {s}: bits differ from the reference's kernel"
    assert mine["TOKENS"] == theirs["TOKENS"] and mine["LENGTHS"] == theirs["LENGTHS"]
    a, b = np.asarray(mine["attention"]), np.asarray(theirs["attention"])
    assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max()


@pytest.mark.parametrize("case", KERNEL_CASES, ids=lambda c: f"B{c[0]}-S{c[1]}-d{c[2]}")
@pytest.mark.parametrize("dist", ["R", "Z"])
@pytest.mark.parametrize("gemm_mode", ["1", "0"], ids=["simt_exact", "tcgen05"])
def test_same_source_cublas_twins(torch_cuda, case, dist, gemm_mode):
    """the reference's fast build (warp-tiling prefill + cuBLAS latest-QKV / logits, paged_attention_with_cublas,
    the cuBLAS decoder) from the same source: cuBLAS's summation order is its own, so K, V, q, attention and logits
    are compared at rel 1e-4 (the reference's own tests use abs 1e-3 there), tokens and lengths exactly"""
    if not KMLI.exists() or not KREF.exists():
        pytest.skip("kernel-level drop-in drivers not built")
    import numpy as np
    args = list(case) + [dist, "cublas"]
    theirs = run_kernels(KREF, args)
    mine = run_kernels(KMLI, args, env={"MLI_GEMM_MODE": gemm_mode})
    for s in ["k_cache", "v_cache", "q", "attention", "logits"]:
        a, b = np.asarray(mine[s]), np.asarray(theirs[s])
        assert a.shape == b.shape and a.size > 0
        assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max(), f"{s}: rel err {np.abs(a - b).max() / np.abs(b).max():.2e}"
    assert mine["LENGTHS"] == theirs["LENGTHS"]
    assert mine["TOKENS"] == theirs["TOKENS"]


# ---- step level: tests/dropin/stepwise_driver.cpp (the caller writes the engine loop itself) ------------------
SMLI = REPO / "tests" / "dropin" / "_build" / "dropin_stepwise_mli"
SREF = REPO / "oracle" / "_ref" / "dropin_stepwise_ref"

# kind, B, S, d, V, n_blocks, n_req, lo, hi, seed, dist, rounds
STEP_CASES = [
    ("dense", 8, 64, 128, 1024, 0, 20, 1, 40, 21, "Z", 1),
    ("paged", 8, 128, 128, 1024, 64, 20, 1, 60, 22, "Z", 1),
    ("paged", 8, 128, 128, 1024, 30, 22, 20, 64, 23, "Z", 1),     # pool pressure: growth, tail pre-emption
    ("paged", 6, 64, 128, 1024, 40, 14, 1, 30, 24, "R", 3),       # three decode rounds per scheduler step
    ("paged", 12, 128, 256, 1024, 50, 30, 10, 64, 25, "Z", 2),
]


@pytest.mark.parametrize("case", STEP_CASES, ids=lambda c: f"{c[0]}-B{c[1]}-blocks{c[5]}-R{c[11]}")
def test_same_source_same_decisions_step_by_step(torch_cuda, case):
    """a user-written engine loop over the reference's scheduling API (insert_new_items, forward,
    process_decoder_result, allocate_or_free_memory_blocks_if_needed): after EVERY iteration the rows admitted, rows
    finished, tokens, device lengths (stale-length quirk included), free-page count and each resident row's page list
    (slab indices, i.e. the physical order of the free list) are identical to the reference build's, and so are the
    final token lists"""
    if not SMLI.exists() or not SREF.exists():
        pytest.skip("step-level drop-in drivers not built")
    e = dict(os.environ, MLI_GEMM_MODE="1")
    outs = []
    for binary in (SREF, SMLI):
        out = subprocess.run([str(binary)] + [str(a) for a in case], capture_output=True, text=True, env=e, timeout=600)
        assert out.returncode == 0, f"{binary.name} failed: {out.stderr[-2000:]}"
        outs.append([l for l in out.stdout.splitlines() if l.split(" ")[0] in ("STEP", "ITERATIONS", "RESULT")])
    theirs, mine = outs
    assert any(l.startswith("STEP") for l in theirs) and sum(l.startswith("RESULT") for l in theirs) == case[6]
    for a, b in zip(mine, theirs):
        assert a == b, f"first difference:\nours: {a[:400]}\nref:  {b[:400]}"
    assert len(mine) == len(theirs)
