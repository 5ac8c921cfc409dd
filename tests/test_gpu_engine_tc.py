"""The production engine path (tensor-core GEMM mode): device scheduler with work lists -> encoder
over 16-position granules -> ONE merged latest-QKV + prefill GEMM -> fused attention -> split-K
logits (partials summed by the decoder) -> decoder, replayed as a multi-step CUDA graph.

Token parity in this mode: 3xTF32 logits differ from the exact-order fp32 chain by ~1e-6 relative,
so a token may legitimately differ where the two best logits are closer than that (the reference's
own naive and cuBLAS builds disagree in the same places).  Every mismatch is therefore classified
with a float64 replay of that request (harness.classify_token_mismatches): it must be a top-2 tie
below 2e-5 of the logit scale, at most 1 % of the requests may contain one, and everything the
scheduler decides (steps, generated tokens, pre-emptions, finish order) must be identical.  The
exact-order mode (MLI_OPT_GEMM_MODE=1) is the bit-exact one and is tested against the reference
in test_gpu_forward_engine.py.

Checked against the CPU oracle (oracle/oracle.c: tokens per request, finish order, and the
scheduler's decision counters) and against the same engine in exact-order SIMT mode, at sizes that
force every branch: more active rows than one 256-row GEMM tile, pool pressure (pre-emption and
re-prefill), several forward rounds per step, the reference's stale-lengths quirk (compat), and the
bench workload itself.  3xTF32 logits differ from the exact chain by ~1e-6 relative; on these fixed
seeds no argmax flips, so tokens are compared exactly."""
import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli
from test_gpu_forward_engine import run_mli_engine

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc(ctx):
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    yield ctx
    ctx.unregister_weights()
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


CASES = [
    # more rows than one GEMM tile (n_valid > 256), roomy pool
    dict(B=300, S=64, d=128, V=1024, n_blocks=1400, n_req=420, lo=1, hi=30, R=1),
    # pool pressure: growth, tail pre-emption, re-prefill of pre-empted requests
    dict(B=64, S=128, d=128, V=1024, n_blocks=260, n_req=150, lo=1, hi=64, R=1),
    # several decode rounds per scheduler step
    dict(B=24, S=128, d=256, V=1024, n_blocks=120, n_req=60, lo=1, hi=48, R=3),
    # emb_dim the tensor-core kernels do not cover: the engine must route to the SIMT kernels
    dict(B=12, S=64, d=96, V=1000, n_blocks=48, n_req=30, lo=1, hi=30, R=1),
    # emb_dim 2048 / 4096 (configs[2] / configs[3]): longer K, more feature tiles than SMs allow splits for
    dict(B=8, S=64, d=2048, V=1024, n_blocks=40, n_req=14, lo=1, hi=40, R=1),
    dict(B=6, S=64, d=4096, V=1024, n_blocks=28, n_req=9, lo=1, hi=40, R=1),
    # BASELINE configs[4] at one GPU: 8192 rows in one engine (scheduler loops over 1024-thread
    # blocks of rows, 32 GEMM tiles per step, 8192-row attention prefix)
    dict(B=8192, S=64, d=128, V=1024, n_blocks=8192 * 4 + 512, n_req=8192 + 600, lo=1, hi=30, R=1),
    # emb_dim 1024 with 1024 rows: the merged GEMM runs its bulk plan (no K split) as a PERSISTENT launch with
    # dynamic (feature tile, activation tile) items; step 0 admits every row at once (a prefill burst of
    # ~130 activation tiles), later steps mix active rows and re-admissions
    dict(B=1024, S=64, d=1024, V=1024, n_blocks=1024 * 4 + 64, n_req=1400, lo=1, hi=40, R=1, max_new=6),
    dict(B=1024, S=64, d=1024, V=1024, n_blocks=1024 * 4 + 64, n_req=1100, lo=10, hi=40, R=2, max_new=4),
]


@pytest.mark.parametrize("case", CASES, ids=[f"B{c['B']}-d{c['d']}-R{c['R']}" for c in CASES])
@pytest.mark.parametrize("compat", [0, 1])
def test_tc_engine_matches_cpu_oracle(torch_cuda, tc, case, compat):
    torch = torch_cuda
    if case["B"] >= 1024 and compat == 1:
        pytest.skip("with 8800 requests a 3xTF32 tie flip is likely, and the float64 tie classifier "
                    "replays the corrected decoding, not the stale-lengths quirk")
    w = H.make_weights(41, case["d"], case["V"], case["S"], "Z")
    offs, toks = H.make_prompts(43, case["n_req"], case["lo"], case["hi"], V=case["V"])
    mine, order, st = run_mli_engine(tc, torch, case, w, offs, toks, compat=compat)
    import os
    rc, want, oorder, ost = H.run_oracle_engine("paged", case, w, offs, toks, fix=1 - compat,
                                                threads=min(16, os.cpu_count() or 8))
    assert rc == 0
    assert st.n_finished == case["n_req"]
    assert (st.steps, st.generated_tokens, st.preemptions) == (ost.steps, ost.generated_tokens,
                                                               ost.preemptions)
    assert order.tolist() == oorder.tolist(), "finish order differs from the oracle"
    ties, errors = H.classify_token_mismatches(w, mine, want)
    assert not errors, f"(request, position, margin) differ from the oracle beyond a numerical tie: {errors[:4]}"
    assert len(ties) <= max(1, case["n_req"] // 100), f"too many tie flips: {ties}"


def test_bench_workload_tc_equals_exact_mode(torch_cuda, tc):
    """the bench job (BASELINE configs[1]) in tensor-core mode produces the token lists of the
    exact-order mode, request by request, with identical scheduler counters"""
    torch = torch_cuda
    from bench import WORKLOADS
    wl = WORKLOADS["c2a"]
    cfg = dict(B=wl["B"], S=wl["S"], d=wl["d"], V=wl["V"], n_blocks=wl["n_blocks"], R=wl["R"])
    w = H.make_weights(1001, wl["d"], wl["V"], wl["S"], "Z")
    offs, toks = H.make_prompts(2002, wl["n_req"], wl["lo"], wl["hi"])
    a, order_a, st_a = run_mli_engine(tc, torch, cfg, w, offs, toks, compat=0)
    tc.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    b, order_b, st_b = run_mli_engine(tc, torch, cfg, w, offs, toks, compat=0)
    assert (st_a.steps, st_a.generated_tokens, st_a.preemptions) == (st_b.steps, st_b.generated_tokens,
                                                                     st_b.preemptions)
    assert order_a.tolist() == order_b.tolist()
    ties, errors = H.classify_token_mismatches(w, a, b)
    assert not errors, f"(request, position, margin) differ beyond a numerical tie: {errors[:4]}"
    assert len(ties) <= wl["n_req"] // 100, f"too many tie flips: {ties}"
    print("tie flips (request, position, relative top-2 margin):", ties)
    # and the job is reproducible run to run (fixed reduction orders everywhere)
    tc.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    c, order_c, _ = run_mli_engine(tc, torch, cfg, w, offs, toks, compat=0)
    assert order_c.tolist() == order_a.tolist()
    assert all(np.array_equal(a[i], c[i]) for i in range(wl["n_req"]))


def test_bounded_runs_resume(torch_cuda, tc):
    """mli_engine_run(max_steps) uses the one-step graph and can be called repeatedly; the result
    equals an unbounded run (which replays the multi-step graph)"""
    torch = torch_cuda
    case = CASES[1]
    w = H.make_weights(41, case["d"], case["V"], case["S"], "Z")
    offs, toks = H.make_prompts(43, case["n_req"], case["lo"], case["hi"])
    whole, order, st = run_mli_engine(tc, torch, case, w, offs, toks, compat=0)
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(case["B"], case["S"], case["d"], case["V"], case["n_blocks"], 1, 0, case["n_req"], None)
    eng = mli.Engine(tc, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    eng.submit(offs, toks)
    for _ in range(2000):
        eng.run(max_steps=7)
        if eng.stats().n_finished == case["n_req"]:
            break
    res, order2 = eng.results()
    eng.close()
    assert order2.tolist() == order.tolist()
    assert all(np.array_equal(res[i], whole[i]) for i in range(case["n_req"]))
