"""P0 parity ladder: every C-ABI stage against the reference's own CUDA launcher
(oracle/_ref, compiled from /root/reference) on identical seeded tensors.

GEMM stages run in MLI_OPT_GEMM_MODE=1 (exact-order SIMT) here and must be BIT-EXACT with the
reference's naive kernels; the tcgen05 mode is covered in test_gpu_tcgen05.py with a tolerance.
The fused decode attention is compared with the reference's qkt -> softmax -> softmax_v chain at
rel 1e-4 (north_star tolerance) on both input distributions.
"""
import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli

pytestmark = pytest.mark.gpu

SHAPES = [
    # B, S, d, V
    (8, 64, 64, 1024),
    (33, 128, 256, 1024),
    (16, 256, 1024, 1024),
    (5, 128, 1028, 1000),   # d % 64 != 0, V % 4 == 0 but not a multiple of the tile
    (7, 512, 2048, 1024),
]


def make_lengths(rng, B, S, zero_frac=0.2):
    L = rng.integers(1, S, size=B).astype(np.int32)      # 1..S-1
    L[rng.random(B) < zero_frac] = 0
    if B > 2:
        L[0] = S - 1        # maximum
        L[1] = 1            # minimum
    return L


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(autouse=True)
def exact_mode(ctx):
    prev = ctx.get_option(mli.OPT_GEMM_MODE)
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    yield
    ctx.set_option(mli.OPT_GEMM_MODE, prev)


@pytest.mark.parametrize("B,S,d,V", SHAPES)
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_encoder_prefill_latest_bit_exact(torch_cuda, ctx, ref, B, S, d, V, dist):
    torch = torch_cuda
    rng = np.random.default_rng(100 + B + d)
    L = make_lengths(rng, B, S)
    case = H.PagedCase(7, B, S, d, L, dist)
    w = H.make_weights(11, d, V, S, dist)
    cand = np.flatnonzero(L > 0)
    n_new = max(1, len(cand) // 2)
    new_idx = np.zeros(B, np.int32)
    new_idx[:n_new] = rng.permutation(cand)[:n_new]
    inp = rng.integers(0, min(V, 1023), size=(B, S)).astype(np.int32)
    pool_a, tab_a = case.device(torch)
    pool_b, tab_b = case.device(torch)
    dw = {k: dev(torch, v) for k, v in w.items()}
    dL, dnew, dinp = dev(torch, L), dev(torch, new_idx), dev(torch, inp)

    # encoder: launch_paged_attention_encoder_kernel (encoder.cu:134-147)
    ctx.call("mli_paged_encoder", dw["emb"], dw["pos"], dinp, tab_a, dL, dnew, B, S, d, n_new)
    H.check_ref(ref.ref_paged_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(dinp), H.p(tab_b), H.p(dL),
                                      H.p(dnew), B, S, d, n_new))
    ctx.synchronize()
    assert torch.equal(pool_a, pool_b), "encoder output differs from the reference"

    # prefill: launch_fill_new_k_v_cache_paged_attention (paged_attention.cu:96-115)
    ctx.call("mli_prefill_kv_paged", tab_a, dnew, dL, dw["wk"], dw["wv"], n_new, B, S, d)
    H.check_ref(ref.ref_prefill_kv_paged(H.p(tab_b), H.p(dnew), H.p(dL), H.p(dw["wk"]), H.p(dw["wv"]),
                                         n_new, B, S, d, 0))
    ctx.synchronize()
    assert torch.equal(pool_a, pool_b), "prefill K/V differ from the reference (must be bit-exact)"

    # latest-token QKV: launch_get_latest_k_q_v_paged_attention (paged_attention.cu:188-199)
    q0 = rng.random((B, d), dtype=np.float32)
    qa, qb = dev(torch, q0), dev(torch, q0)
    ctx.call("mli_qkv_latest_paged", tab_a, dL, dw["wk"], dw["wq"], dw["wv"], qa, B, S, d)
    H.check_ref(ref.ref_qkv_latest_paged(H.p(tab_b), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]),
                                         H.p(dw["wv"]), H.p(qb), B, S, d, 0))
    ctx.synchronize()
    assert torch.equal(pool_a, pool_b), "latest K/V differ from the reference"
    assert torch.equal(qa, qb), "q_output differs from the reference (rows with L == 0 untouched)"

    # the reference's warp-tiling / cuBLAS build only has to agree within tolerance (opaque order)
    pool_c, tab_c = case.device(torch)
    H.check_ref(ref.ref_paged_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(dinp), H.p(tab_c), H.p(dL),
                                      H.p(dnew), B, S, d, n_new))
    H.check_ref(ref.ref_prefill_kv_paged(H.p(tab_c), H.p(dnew), H.p(dL), H.p(dw["wk"]), H.p(dw["wv"]),
                                         n_new, B, S, d, 1))
    qc = dev(torch, q0)
    H.check_ref(ref.ref_qkv_latest_paged(H.p(tab_c), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]),
                                         H.p(dw["wv"]), H.p(qc), B, S, d, 1))
    assert H.rel_err(pool_a.cpu().numpy(), pool_c.cpu().numpy()) < 1e-4
    # the cuBLAS build multiplies an uninitialised latest_emb row for empty rows
    # (paged_attention_cublas.cu:23-25 skips them), so only rows with L > 0 are comparable
    live = L > 0
    assert H.rel_err(qa.cpu().numpy()[live], qc.cpu().numpy()[live]) < 1e-4


@pytest.mark.parametrize("B,S,d,V", SHAPES + [(4, 2048, 1024, 1024), (3, 1024, 4096, 1024)])
@pytest.mark.parametrize("dist", ["R", "Z"])
@pytest.mark.parametrize("chunk_pages", [0, 1, 3])
def test_fused_decode_attention(torch_cuda, ctx, ref, B, S, d, V, dist, chunk_pages):
    torch = torch_cuda
    rng = np.random.default_rng(200 + B + d + S)
    L = make_lengths(rng, B, S)
    case = H.PagedCase(9, B, S, d, L, dist)
    pool, tab = case.device(torch)
    if dist == "R":
        q = H.uniform01(rng, (B, d))
    else:
        q = ((rng.random((B, d), dtype=np.float32) - 0.5) * 2.0 * np.sqrt(12.0 / d)).astype(np.float32)
    dq, dL = dev(torch, q), dev(torch, L)
    out = torch.full((B, d), 7.0, device="cuda")
    probs = torch.full((B, S), 7.0, device="cuda")
    ctx.set_option(mli.OPT_ATTN_CHUNK_PAGES, chunk_pages)
    try:
        ctx.call("mli_decode_attention_paged", dq, tab, dL, out, probs, B, S, d)
        ctx.synchronize()
    finally:
        ctx.set_option(mli.OPT_ATTN_CHUNK_PAGES, 0)
    # reference chain: qkt (paged_attention.cu:270) -> softmax (:360) -> softmax_v (:333)
    qkt = torch.zeros((B, S), device="cuda")
    ref_out = torch.full((B, d), 7.0, device="cuda")
    H.check_ref(ref.ref_qkt_paged(H.p(dq), H.p(tab), H.p(dL), H.p(qkt), B, S, d))
    H.check_ref(ref.ref_softmax_in_place_with_lengths(H.p(qkt), H.p(dL), B, S))
    H.check_ref(ref.ref_softmax_v_paged(H.p(qkt), H.p(tab), H.p(ref_out), H.p(dL), B, S, d))
    a, b = out.cpu().numpy(), ref_out.cpu().numpy()
    assert np.all(a[L == 0] == 0.0), "empty rows must produce zeros"
    assert H.rel_err(a, b) < 1e-4, f"attention rel err {H.rel_err(a, b):.3e}"
    pa, pb = probs.cpu().numpy(), qkt.cpu().numpy()
    assert np.abs(pa - pb).max() < 1e-4, "softmax probabilities differ"
    for r in range(B):
        assert np.all(pa[r, L[r]:] == 0.0), "probabilities past L must be zero-filled"


@pytest.mark.parametrize("B,S,d,V", SHAPES + [(4, 1024, 1024, 1024)])
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_unfused_stages_bit_exact(torch_cuda, ctx, ref, B, S, d, V, dist):
    """mli_qkt_paged / mli_softmax_in_place_with_lengths / mli_softmax_v_paged (the one-to-one mirrors of
    the reference's three unfused launchers, tests/paged_attention_kernels_test.cpp:115-169): raw scores,
    probabilities and P.V all bit-exact (same arithmetic order in every stage, the warp reduction of the
    softmax included), untouched entries kept"""
    torch = torch_cuda
    rng = np.random.default_rng(250 + B + d + S)
    L = make_lengths(rng, B, S)
    case = H.PagedCase(9, B, S, d, L, dist)
    pool, tab = case.device(torch)
    q = H.uniform01(rng, (B, d)) if dist == "R" else \
        ((rng.random((B, d), dtype=np.float32) - 0.5) * 2.0 * np.sqrt(12.0 / d)).astype(np.float32)
    dq, dL = dev(torch, q), dev(torch, L)
    mine = torch.full((B, S), 5.0, device="cuda")
    theirs = torch.full((B, S), 5.0, device="cuda")
    ctx.call("mli_qkt_paged", dq, tab, dL, mine, B, S, d)
    ctx.synchronize()
    H.check_ref(ref.ref_qkt_paged(H.p(dq), H.p(tab), H.p(dL), H.p(theirs), B, S, d))
    assert torch.equal(mine, theirs), "raw scores differ from launch_qkt_paged_attention"
    ctx.call("mli_softmax_in_place_with_lengths", mine, dL, B, S)
    ctx.synchronize()
    H.check_ref(ref.ref_softmax_in_place_with_lengths(H.p(theirs), H.p(dL), B, S))
    pa, pb = mine.cpu().numpy(), theirs.cpu().numpy()
    assert torch.equal(mine, theirs), f"probabilities differ from launch_softmax_in_place_with_lengths by {np.abs(pa - pb).max():.2e}"
    for r in range(B):
        assert np.all(pa[r, L[r]:] == 0.0)
    out_a = torch.full((B, d), 7.0, device="cuda")
    out_b = torch.full((B, d), 7.0, device="cuda")
    ctx.call("mli_softmax_v_paged", mine, tab, out_a, dL, B, S, d)
    ctx.synchronize()
    H.check_ref(ref.ref_softmax_v_paged(H.p(theirs), H.p(tab), H.p(out_b), H.p(dL), B, S, d))
    assert torch.equal(out_a, out_b), "P.V differs from launch_softmax_v_paged_attention"


@pytest.mark.parametrize("B,S,d,V", SHAPES)
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_decoder_bit_exact(torch_cuda, ctx, ref, B, S, d, V, dist):
    torch = torch_cuda
    rng = np.random.default_rng(300 + B + d)
    L = make_lengths(rng, B, S)
    L[-1] = S - 1   # hits the L+1 >= S rule (decoder.cu:176)
    case = H.PagedCase(13, B, S, d, L, dist)
    w = H.make_weights(17, d, V, S, dist, eof_ratio=1.3)
    attn = H.uniform01(rng, (B, d)) if dist == "R" else (rng.random((B, d), dtype=np.float32) - 0.5)
    pool_a, tab_a = case.device(torch)
    pool_b, tab_b = case.device(torch)
    dw = {k: dev(torch, v) for k, v in w.items()}
    dattn = dev(torch, attn)
    La, Lb = dev(torch, L), dev(torch, L)
    n_dec, i_dec = 3, 1
    da = torch.full((B, n_dec), -5, dtype=torch.int32, device="cuda")
    db = torch.full((B, n_dec), -5, dtype=torch.int32, device="cuda")
    sa = torch.zeros((B, V), device="cuda")
    sb = torch.zeros((B, V), device="cuda")
    ctx.call("mli_paged_decoder", dattn, dw["emb"], sa, dw["pos"], tab_a, La, da, B, V, S, d, n_dec, i_dec)
    H.check_ref(ref.ref_paged_decoder(H.p(dattn), H.p(dw["emb"]), H.p(sb), H.p(dw["pos"]), H.p(tab_b),
                                      H.p(Lb), H.p(db), B, V, S, d, n_dec, i_dec, 0))
    ctx.synchronize()
    assert torch.equal(sa, sb), "logits differ from gemm_transpose_kernel (must be bit-exact)"
    assert torch.equal(da, db), "tokens differ"
    assert torch.equal(La, Lb), "lengths differ"
    assert torch.equal(pool_a, pool_b), "next-token embedding differs"
    # cuBLAS decoder: tolerance on logits, exact on tokens (tests/decoder_test.cpp:270)
    Lc = dev(torch, L)
    dc = torch.full((B, n_dec), -5, dtype=torch.int32, device="cuda")
    sc = torch.zeros((B, V), device="cuda")
    pool_c, tab_c = case.device(torch)
    H.check_ref(ref.ref_paged_decoder(H.p(dattn), H.p(dw["emb"]), H.p(sc), H.p(dw["pos"]), H.p(tab_c),
                                      H.p(Lc), H.p(dc), B, V, S, d, n_dec, i_dec, 1))
    assert H.rel_err(sa.cpu().numpy(), sc.cpu().numpy()) < 1e-4


def test_argmax_tie_rule(torch_cuda, ctx):
    """on exact ties the device tree keeps min (bitreverse8(index % 256), index); see
    test_cpu_oracle.py::test_argmax_device_rule"""
    torch = torch_cuda
    B, S, d, V = 2, 64, 64, 1024
    L = np.array([3, 5], np.int32)
    case = H.PagedCase(1, B, S, d, L, "R")
    pool, tab = case.device(torch)
    # emb rows 300 and 513 identical and maximal: threads 44 and 1 differ in bit 0 -> 300 wins
    emb = np.zeros((V, d), np.float32)
    emb[300] = 1.0
    emb[513] = 1.0
    attn = np.ones((B, d), np.float32)
    pos = np.zeros((S, d), np.float32)
    dL = dev(torch, L)
    dec = torch.zeros((B, 1), dtype=torch.int32, device="cuda")
    ctx.call("mli_paged_decoder", dev(torch, attn), dev(torch, emb), None, dev(torch, pos), tab, dL, dec,
             B, V, S, d, 1, 0)
    ctx.synchronize()
    assert dec.cpu().numpy().ravel().tolist() == [300, 300]
    score = emb @ attn[0]
    assert H.load_oracle().orc_argmax_device_rule(H.p(score.astype(np.float32)), V) == 300
