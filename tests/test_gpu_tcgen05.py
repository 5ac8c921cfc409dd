"""tcgen05 3xTF32 GEMM mode (MLI_OPT_GEMM_MODE=0) against the exact-order SIMT mode and the
reference CUDA kernels.  Tolerance: rel 1e-5 on K/V/q/logits up to emb_dim 2048 and 2e-5 at emb_dim
4096 (the products are exact to ~5e-7; what is left is the tensor core's truncating fp32 accumulation
over chains of up to 1024 products, worst on the reference's all-positive U(0,1] data; the north_star
bound is 1e-4), tokens exact on these seeds.  Shapes the tensor-core path does not cover (emb_dim % 128,
n_vocab % 128) must silently use the SIMT kernels and stay bit-exact."""
import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture()
def tc(ctx):
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    yield ctx
    ctx.unregister_weights()
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


@pytest.mark.parametrize("B,V,d", [(8, 128, 128), (64, 1024, 256), (256, 1024, 1024), (77, 1024, 2048),
                                   (300, 2048, 512), (128, 1024, 4096)])   # last: the configs[3] shape
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_logits(torch_cuda, tc, ref, B, V, d, dist):
    torch = torch_cuda
    S = 64
    rng = np.random.default_rng(B + V + d)
    w = H.make_weights(5, d, V, S, dist)
    attn = H.uniform01(rng, (B, d)) if dist == "R" else (rng.random((B, d), dtype=np.float32) - 0.5)
    L = np.ones(B, np.int32) * 3
    case = H.PagedCase(1, B, S, d, L, dist)
    pool, tab = case.device(torch)
    demb, dpos, dattn = dev(torch, w["emb"]), dev(torch, w["pos"]), dev(torch, attn)
    outs = []
    for mode in (mli.GEMM_TCGEN05, mli.GEMM_SIMT_EXACT):
        tc.set_option(mli.OPT_GEMM_MODE, mode)
        score = torch.zeros((B, V), device="cuda")
        dec = torch.zeros((B, 1), dtype=torch.int32, device="cuda")
        dL = dev(torch, L)
        tc.call("mli_paged_decoder", dattn, demb, score, dpos, tab, dL, dec, B, V, S, d, 1, 0)
        tc.synchronize()
        outs.append((score.cpu().numpy(), dec.cpu().numpy()))
    err = H.rel_err(outs[0][0], outs[1][0])
    assert err < (TOL if d <= 2048 else 2 * TOL), f"logits rel err {err:.2e}"
    assert np.array_equal(outs[0][1], outs[1][1]), "tokens differ between tcgen05 and exact mode"


@pytest.mark.parametrize("B,S,d", [(8, 64, 128), (33, 128, 256), (256, 128, 1024), (40, 256, 2048),
                                   (48, 1024, 128),    # more activation tiles than the grid cap
                                   (24, 64, 4096)])    # emb_dim of BASELINE configs[3]
@pytest.mark.parametrize("dist", ["R", "Z"])
@pytest.mark.parametrize("registered", [False, True])
def test_latest_and_prefill(torch_cuda, tc, ref, B, S, d, dist, registered):
    torch = torch_cuda
    V = 1024
    rng = np.random.default_rng(B + S + d)
    L = rng.integers(1, S, size=B).astype(np.int32)
    L[rng.random(B) < 0.2] = 0
    L[0] = S - 1
    case = H.PagedCase(7, B, S, d, L, dist)
    w = H.make_weights(11, d, V, S, dist)
    cand = np.flatnonzero(L > 0)
    n_new = max(1, len(cand) // 2) if S < 1024 else len(cand)
    new_idx = np.zeros(B, np.int32)
    new_idx[:n_new] = rng.permutation(cand)[:n_new]
    dw = {k: dev(torch, v) for k, v in w.items()}
    dL, dnew = dev(torch, L), dev(torch, new_idx)
    if registered:
        tc.register_weights(dw["wk"], dw["wq"], dw["wv"], dw["emb"], d, V)
    res = []
    for mode in (mli.GEMM_TCGEN05, mli.GEMM_SIMT_EXACT):
        tc.set_option(mli.OPT_GEMM_MODE, mode)
        pool, tab = case.device(torch)
        q = torch.full((B, d), 3.0, device="cuda")
        tc.call("mli_prefill_kv_paged", tab, dnew, dL, dw["wk"], dw["wv"], n_new, B, S, d)
        tc.call("mli_qkv_latest_paged", tab, dL, dw["wk"], dw["wq"], dw["wv"], q, B, S, d)
        tc.synchronize()
        res.append((pool.cpu().numpy(), q.cpu().numpy()))
    # the reference's naive kernels as a third witness
    pool_r, tab_r = case.device(torch)
    qr = torch.full((B, d), 3.0, device="cuda")
    H.check_ref(ref.ref_prefill_kv_paged(H.p(tab_r), H.p(dnew), H.p(dL), H.p(dw["wk"]), H.p(dw["wv"]),
                                         n_new, B, S, d, 0))
    H.check_ref(ref.ref_qkv_latest_paged(H.p(tab_r), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]), H.p(dw["wv"]),
                                         H.p(qr), B, S, d, 0))
    assert np.array_equal(res[1][0], pool_r.cpu().numpy()) and np.array_equal(res[1][1], qr.cpu().numpy())
    e_pool, e_q = H.rel_err(res[0][0], res[1][0]), H.rel_err(res[0][1], res[1][1])
    tol = TOL if d <= 2048 else 2 * TOL
    assert e_pool < tol and e_q < tol, f"K/V rel err {e_pool:.2e}, q rel err {e_q:.2e}"
    # untouched regions (other sub-rows, rows with L == 0, q rows of empty rows) stay bit-identical
    untouched = res[1][0] == case.pool
    assert np.array_equal(res[0][0][untouched], case.pool[untouched])
    assert np.all(res[0][1][L == 0] == 3.0)


def test_uncovered_shapes_fall_back_to_exact_simt(torch_cuda, tc, ref):
    torch = torch_cuda
    B, S, d, V = 5, 64, 132, 1000          # d % 128 != 0, V % 128 != 0
    rng = np.random.default_rng(3)
    L = rng.integers(1, S, size=B).astype(np.int32)
    case = H.PagedCase(7, B, S, d, L, "Z")
    w = H.make_weights(11, d, V, S, "Z")
    dw = {k: dev(torch, v) for k, v in w.items()}
    dL = dev(torch, L)
    pa, ta = case.device(torch)
    pb, tb = case.device(torch)
    qa, qb = torch.zeros((B, d), device="cuda"), torch.zeros((B, d), device="cuda")
    tc.call("mli_qkv_latest_paged", ta, dL, dw["wk"], dw["wq"], dw["wv"], qa, B, S, d)
    tc.synchronize()
    H.check_ref(ref.ref_qkv_latest_paged(H.p(tb), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]), H.p(dw["wv"]),
                                         H.p(qb), B, S, d, 0))
    assert torch.equal(pa, pb) and torch.equal(qa, qb)


@pytest.mark.parametrize("dist", ["R", "Z"])
def test_engine_tcgen05_tokens_match_reference_dense_engine(torch_cuda, tc, ref, dist):
    """end to end in tensor-core mode: tokens of every request equal the reference's non-paged
    engine (P2) on these seeds"""
    from test_gpu_forward_engine import run_mli_engine, run_ref_engine
    torch = torch_cuda
    case = dict(B=16, S=128, d=256, V=1024, n_blocks=64, n_req=40, lo=1, hi=64)
    w = H.make_weights(31, case["d"], case["V"], case["S"], dist)
    offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
    tc.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    mine, order, st = run_mli_engine(tc, torch, case, w, offs, toks, compat=0)
    theirs, _, _ = run_ref_engine(ref, "dense", case, w, offs, toks)
    bad = [i for i in range(case["n_req"]) if not np.array_equal(mine[i], theirs[i])]
    assert not bad, f"requests {bad} differ"


def gemm_plan(ctx, kind):
    """kind: 0 latest-token stage, 1 prefill stage, 2 logits, 3 the engine's merged step launch"""
    import ctypes as C
    plan = (C.c_int * 5)()
    ctx._check(ctx.lib.mli_debug_last_gemm_plan(ctx.h, kind, plan))
    return list(plan)


@pytest.mark.parametrize("d", [1024, 2048, 4096])
def test_three_launch_shapes_of_the_bulk_gemm_agree_bitwise(torch_cuda, tc, d, monkeypatch):
    """launches of a thousand rows and more run the tcgen05 GEMM as a static grid, as a persistent launch with
    dynamic tiles or on CTA pairs (cta_group::2).  Every output element sees the same k order, the same two
    accumulators and the same operand rounding (cvt.rna vs its integer form) in all three, so K, V, q, the
    logits and the tokens must be BIT-identical -- and the plan query proves each variant really ran.  The
    engine job adds the merged launch with prefill granules (step 0 admits 1024 prompts at once)."""
    torch = torch_cuda
    from test_gpu_forward_engine import run_mli_engine
    B, S, V = 1536, 16, 1024     # (the logits launch drops its K split, i.e. goes bulk, above ~1200 rows)
    rng = np.random.default_rng(d)
    L = rng.integers(1, S - 1, size=B).astype(np.int32)
    L[5] = 0
    case = H.PagedCase(1, B, S, d, L, "Z")
    w = H.make_weights(11, d, V, S, "Z")
    dw = {k: dev(torch, v) for k, v in w.items()}
    attn = dev(torch, rng.random((B, d), dtype=np.float32) - 0.5)
    job = dict(B=1024, S=64, d=d, V=V, n_blocks=1024 * 4 + 64, n_req=1100, lo=4, hi=40, R=1, max_new=3)
    offs, toks = H.make_prompts(43, job["n_req"], job["lo"], job["hi"])
    wj = H.make_weights(11, d, V, job["S"], "Z")      # (the job's own position table: 64 rows)
    out = {}
    for name, env, want in (("pairs", {}, 2), ("persistent", {"MLI_TC_NO_PAIR": "1"}, 1),
                            ("static", {"MLI_TC_STATIC_TILES": "1"}, 0)):
        for k in ("MLI_TC_NO_PAIR", "MLI_TC_STATIC_TILES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        pool, tab = case.device(torch)
        dL = dev(torch, L)
        q = torch.zeros((B, d), device="cuda")
        tc.call("mli_qkv_latest_paged", tab, dL, dw["wk"], dw["wq"], dw["wv"], q, B, S, d)
        scores = torch.zeros((B, V), device="cuda")
        dec = torch.zeros((B, 1), dtype=torch.int32, device="cuda")
        tc.call("mli_paged_decoder", attn, dw["emb"], scores, dw["pos"], tab, dL, dec, B, V, S, d, 1, 0)
        plan_logits = gemm_plan(tc, 2)
        tc.synchronize()
        mine, order, st = run_mli_engine(tc, torch, job, wj, offs, toks, compat=0)
        plan_step = gemm_plan(tc, 3)
        if d <= 2048:
            assert plan_logits[0] == want, (name, plan_logits)
            assert plan_step[0] == want and plan_step[1] == 1 and plan_step[4] == 1, (name, plan_step)
        else:
            # emb_dim 4096: chains of at most 1024 products need K cut in four.  The pair kernel walks K in two
            # passes of two accumulators; without it the static grid splits K over clusters of two CTAs.  Both add
            # (a0 + a1) + (a2 + a3), so the results must still be bit-identical
            assert plan_step[0] == (2 if name == "pairs" else 0), (name, plan_step)
            assert (plan_step[1], plan_step[4]) == ((1, 2) if name == "pairs" else (2, 1)), (name, plan_step)
        flat = np.concatenate([mine[i] for i in range(job["n_req"])])
        out[name] = (pool.cpu().numpy(), q.cpu().numpy(), scores.cpu().numpy(), dec.cpu().numpy(), flat,
                     np.asarray(order))
    for name in ("persistent", "static"):
        for a, b, what in zip(out["pairs"], out[name], ("pages", "q", "logits", "tokens", "engine tokens",
                                                        "engine finish order")):
            assert np.array_equal(a, b), f"{what}: the CTA-pair launch and the {name} launch differ"
