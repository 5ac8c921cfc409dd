"""P1-P3 of the parity ladder.

P1  mli_paged_forward vs the reference's PagedAttentionInferenceModel::forward on identical
    (inp, lengths, new rows, page table, pages): tokens and lengths exact, pages within 1e-4
    (attention feeds the next embedding only through the token, so pages are in fact bit-exact
    in exact-GEMM mode).
P2  device engine (corrected lengths) vs the reference's NON-paged engine start_inference_engine,
    per request id -- the clean end-to-end pin (the non-paged engine has no Q1 quirk).
P3  device engine in compat mode vs start_paged_attention{,_cublas}_inference_engine.
Mismatching tokens are only tolerated when the reference's own top-2 logit margin is inside fp32
re-association noise; with the fixed seeds below there are none.
"""
import ctypes as C

import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli

pytestmark = pytest.mark.gpu


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("gemm_mode", [mli.GEMM_SIMT_EXACT, mli.GEMM_TCGEN05])
@pytest.mark.parametrize("B,S,d,V,R", [(8, 64, 64, 1024, 1), (24, 128, 256, 1024, 1),
                                       (12, 128, 1024, 1024, 3), (6, 256, 2048, 1024, 2)])
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_paged_forward_matches_reference(torch_cuda, ctx, ref, B, S, d, V, R, dist, gemm_mode):
    torch = torch_cuda
    if gemm_mode == mli.GEMM_TCGEN05:
        try:
            ctx.set_option(mli.OPT_GEMM_MODE, gemm_mode)
        except mli.MliError:
            pytest.skip("tcgen05 path not available")
    else:
        ctx.set_option(mli.OPT_GEMM_MODE, gemm_mode)
    rng = np.random.default_rng(400 + B + d + R)
    L = rng.integers(1, S - R - 1, size=B).astype(np.int32)
    L[rng.random(B) < 0.2] = 0
    # rows need pages for positions up to L+R
    case = H.PagedCase(21, B, S, d, np.minimum(L + R, S - 1) * (L > 0), dist)
    case.lengths = L
    w = H.make_weights(23, d, V, S, dist, eof_ratio=1.001)
    cand = np.flatnonzero(L > 0)
    n_new = max(1, len(cand) // 2)
    new_idx = np.zeros(B, np.int32)
    new_idx[:n_new] = rng.permutation(cand)[:n_new]
    inp = rng.integers(0, 1023, size=(B, S)).astype(np.int32)
    dw = {k: dev(torch, v) for k, v in w.items()}
    dnew, dinp = dev(torch, new_idx), dev(torch, inp)

    # rows that are not "new" must already hold embeddings and K/V: prefill them on both sides with
    # the reference so the starting state is identical
    pools, tabs, lens, decs = [], [], [], []
    for _ in range(2):
        pool, tab = case.device(torch)
        dL = dev(torch, L)
        all_idx = dev(torch, np.arange(B, dtype=np.int32))
        H.check_ref(ref.ref_paged_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(dinp), H.p(tab),
                                          H.p(dL), H.p(all_idx), B, S, d, B))
        H.check_ref(ref.ref_prefill_kv_paged(H.p(tab), H.p(all_idx), H.p(dL), H.p(dw["wk"]),
                                             H.p(dw["wv"]), B, B, S, d, 0))
        pools.append(pool); tabs.append(tab); lens.append(dL)
        decs.append(torch.full((B, R), -7, dtype=torch.int32, device="cuda"))
    ctx.call("mli_paged_forward", dinp, lens[0], dnew, decs[0], n_new, dw["emb"], dw["pos"], tabs[0],
             dw["wk"], dw["wq"], dw["wv"], None, None, B, S, d, V, R)
    ctx.synchronize()
    H.check_ref(ref.ref_paged_forward(H.p(dinp), H.p(lens[1]), H.p(dnew), H.p(decs[1]), n_new,
                                      H.p(dw["emb"]), H.p(dw["pos"]), H.p(tabs[1]), H.p(dw["wk"]),
                                      H.p(dw["wq"]), H.p(dw["wv"]), B, S, d, V, R, 0))
    assert torch.equal(decs[0], decs[1]), "tokens differ from the reference forward"
    assert torch.equal(lens[0], lens[1]), "lengths differ from the reference forward"
    err = H.rel_err(pools[0].cpu().numpy(), pools[1].cpu().numpy())
    if gemm_mode == mli.GEMM_SIMT_EXACT:
        assert torch.equal(pools[0], pools[1]), f"pages differ (rel {err:.2e})"
    else:
        assert err < 1e-4
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


def run_mli_engine(ctx, torch, cfg, w, offs, toks, compat):
    dw = {k: dev(torch, v) for k, v in w.items()}
    ec = mli.EngineCfg(cfg["B"], cfg["S"], cfg["d"], cfg["V"], cfg["n_blocks"], cfg.get("R", 1),
                       compat, len(offs) - 1, None, cfg.get("max_new", 0), cfg.get("max_prefill", 0),
                       cfg.get("chunk", 0))
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    eng.submit(offs, toks)
    eng.run()
    res, order = eng.results()
    st = eng.stats()
    eng.close()
    return res, order, st


def run_ref_engine(ref, kind, cfg, w, offs, toks, variant=0):
    n_req = len(offs) - 1
    S = cfg["S"]
    ids = np.zeros(n_req, np.int32)
    fo = np.zeros(n_req + 1, np.int32)
    ft = np.zeros(n_req * S, np.int32)
    nf = C.c_int(0)
    sec = C.c_double(0)
    if kind == "paged":
        H.check_ref(ref.ref_run_paged_engine(variant, cfg["B"], S, cfg["d"], cfg["V"], cfg["n_blocks"],
                                             cfg.get("R", 1), H.p(w["emb"]), H.p(w["pos"]),
                                             H.p(w["wk"]), H.p(w["wq"]), H.p(w["wv"]), n_req,
                                             H.p(offs), H.p(toks), H.p(ids), H.p(fo), H.p(ft),
                                             C.byref(nf), C.byref(sec)))
    else:
        H.check_ref(ref.ref_run_dense_engine(cfg["B"], S, cfg["d"], cfg["V"], H.p(w["emb"]),
                                             H.p(w["pos"]), H.p(w["wk"]), H.p(w["wq"]), H.p(w["wv"]),
                                             n_req, H.p(offs), H.p(toks), H.p(ids), H.p(fo), H.p(ft),
                                             C.byref(nf), C.byref(sec)))
    k = nf.value
    return {int(ids[i]): ft[fo[i]:fo[i + 1]].copy() for i in range(k)}, ids[:k].copy(), sec.value


ENGINE_CASES = [
    # B, S, d, V, n_blocks, n_req, prompt lo/hi
    dict(B=4, S=64, d=64, V=1024, n_blocks=16, n_req=10, lo=1, hi=40),
    dict(B=16, S=128, d=256, V=1024, n_blocks=64, n_req=40, lo=1, hi=64),
    dict(B=8, S=128, d=128, V=1024, n_blocks=36, n_req=24, lo=20, hi=64),   # pool pressure: pre-emption
]


@pytest.mark.parametrize("case", ENGINE_CASES)
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_engine_vs_reference_dense_engine(torch_cuda, ctx, ref, case, dist):
    """P2: corrected device engine == the reference's non-paged engine, per request id"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    w = H.make_weights(31, case["d"], case["V"], case["S"], dist)
    offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
    mine, order, st = run_mli_engine(ctx, torch, case, w, offs, toks, compat=0)
    theirs, _, _ = run_ref_engine(ref, "dense", case, w, offs, toks)
    assert st.n_finished == case["n_req"] == len(theirs)
    for i in range(case["n_req"]):
        assert np.array_equal(mine[i], theirs[i]), f"request {i}: tokens differ from the reference"


@pytest.mark.parametrize("case", [c for c in ENGINE_CASES if c["d"] in (128, 256)],
                         ids=lambda c: f"B{c['B']}-d{c['d']}")
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_engine_with_warp_per_position_attention(torch_cuda, ctx, ref, case, dist):
    """P2 again with the attention's other consumer design forced (MLI_OPT_ATTN_KERNEL = 2; auto picks it
    only for long launches): tokens of every request equal the reference's non-paged engine"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    ctx.set_option(mli.OPT_ATTN_KERNEL, 2)
    try:
        w = H.make_weights(31, case["d"], case["V"], case["S"], dist)
        offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
        mine, order, st = run_mli_engine(ctx, torch, case, w, offs, toks, compat=0)
    finally:
        ctx.set_option(mli.OPT_ATTN_KERNEL, 0)
    theirs, _, _ = run_ref_engine(ref, "dense", case, w, offs, toks)
    assert st.n_finished == case["n_req"] == len(theirs)
    bad = [i for i in range(case["n_req"]) if not np.array_equal(mine[i], theirs[i])]
    assert not bad, f"requests {bad}: tokens differ from the reference"


@pytest.mark.parametrize("case", ENGINE_CASES)
@pytest.mark.parametrize("dist", ["R", "Z"])
@pytest.mark.parametrize("variant", [0, 1])
def test_engine_compat_vs_reference_paged_engine(torch_cuda, ctx, ref, case, dist, variant):
    """P3: compat (Q1-replaying) device engine == the reference's paged engines, per request id,
    and in the same finish order"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    w = H.make_weights(31, case["d"], case["V"], case["S"], dist)
    offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
    mine, order, st = run_mli_engine(ctx, torch, case, w, offs, toks, compat=1)
    theirs, their_order, _ = run_ref_engine(ref, "paged", case, w, offs, toks, variant)
    assert st.n_finished == case["n_req"] == len(theirs)
    assert order.tolist() == their_order.tolist(), "finish order differs"
    bad = [i for i in range(case["n_req"]) if not np.array_equal(mine[i], theirs[i])]
    assert not bad, f"requests {bad}: tokens differ from the reference paged engine"


@pytest.mark.parametrize("case", ENGINE_CASES)
def test_engine_matches_oracle_stats(torch_cuda, ctx, case):
    """device scheduler makes the same decisions as the CPU restatement: steps, pre-emptions"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    w = H.make_weights(31, case["d"], case["V"], case["S"], "Z")
    offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
    for compat in (0, 1):
        mine, order, st = run_mli_engine(ctx, torch, case, w, offs, toks, compat=compat)
        rc, theirs, oorder, ost = H.run_oracle_engine("paged", case, w, offs, toks, fix=1 - compat)
        assert rc == 0
        assert (st.steps, st.generated_tokens, st.preemptions) == (ost.steps, ost.generated_tokens,
                                                                   ost.preemptions)
        assert order.tolist() == oorder.tolist()
        for i in range(case["n_req"]):
            assert np.array_equal(mine[i], theirs[i])
