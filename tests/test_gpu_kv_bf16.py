"""Opt-in compact KV page format (MLI_OPT_KV_FORMAT = 1): a position is
[inp fp32 x d | K bf16 x d | V bf16 x d] = 8*d bytes instead of 12*d.  It is NOT the reference's
layout (include/utils.h:32-60), so there is no reference build to compare with; parity is stated as

  * projection: the K / V rows the tcgen05 GEMM stores are the round-to-nearest-even bf16 of what the
    exact-order fp32 path computes (one bf16 ulp of slack for values that sit on a rounding boundary:
    3xTF32 vs the exact chain differ by ~1e-6), q stays fp32 within 1e-5; embeddings untouched,
  * attention: on pages that already hold bf16 K / V the single-launch kernel equals a float64
    evaluation of the same (rounded) inputs at rel 1e-4 -- storage is the only difference,
  * engine: every request finishes, the scheduler invariants hold, and the first generated token of
    a request agrees with the fp32-format engine for >= 90 % of the requests (K, V carry a 2^-9
    relative rounding, so near-ties may flip; later tokens then legitimately diverge).
"""
import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli
from test_gpu_forward_engine import run_mli_engine

pytestmark = pytest.mark.gpu
PAGE = 16


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def to_bf16_bits(x):
    """round-to-nearest-even fp32 -> bf16 bit patterns (uint16)"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b):
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


@pytest.fixture()
def compact(ctx):
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
        ctx.set_option(mli.OPT_KV_FORMAT, 1)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    yield ctx
    ctx.set_option(mli.OPT_KV_FORMAT, 0)
    ctx.unregister_weights()
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


class CompactCase:
    """pages in the compact layout: float32 words [16][2*d]: inp (d), K bf16 (d/2 words), V bf16 (d/2)"""

    def __init__(self, seed, B, S, d, lengths):
        rng = np.random.default_rng(seed)
        self.B, self.S, self.d, self.W = B, S, d, S // PAGE
        self.lengths = np.asarray(lengths, np.int32)
        need = [(-(-min(int(L) + 1, S) // PAGE) if L > 0 else 0) for L in self.lengths]
        self.n_pages = max(1, int(sum(need)))
        self.page_words = PAGE * 2 * d
        order = rng.permutation(self.n_pages)
        self.page_ids = -np.ones((B, self.W), np.int64)
        k = 0
        for r in range(B):
            for j in range(need[r]):
                self.page_ids[r, j] = order[k]
                k += 1
        self.inp = ((rng.random((self.n_pages, PAGE, d), dtype=np.float32) - 0.5) * 2.0).astype(np.float32)
        self.K = ((rng.random((self.n_pages, PAGE, d), dtype=np.float32) - 0.5) * 2.0).astype(np.float32)
        self.V = ((rng.random((self.n_pages, PAGE, d), dtype=np.float32) - 0.5) * 2.0).astype(np.float32)

    def pool(self):
        p = np.zeros((self.n_pages, PAGE, 2 * self.d), np.float32)
        p[:, :, :self.d] = self.inp
        kv = np.concatenate([to_bf16_bits(self.K), to_bf16_bits(self.V)], axis=2)      # uint16 [.., 2d]
        p[:, :, self.d:] = np.ascontiguousarray(kv).view(np.float32)
        return p.reshape(self.n_pages, self.page_words)

    def device(self, torch):
        pool = torch.from_numpy(self.pool()).cuda()
        t = np.zeros((self.B, self.W), np.uint64)
        m = self.page_ids >= 0
        t[m] = np.uint64(pool.data_ptr()) + self.page_ids[m].astype(np.uint64) * np.uint64(self.page_words * 4)
        return pool, torch.from_numpy(t.view(np.int64)).cuda()

    def rows(self, pool_np, r, L):
        """(inp [L,d] f32, K [L,d] f32 from bf16, V) of row r out of a pool dump"""
        d = self.d
        inp, K, V = [], [], []
        for j in range(L):
            pos = pool_np[int(self.page_ids[r, j // PAGE])].reshape(PAGE, 2 * d)[j % PAGE]
            kv = np.ascontiguousarray(pos[d:]).view(np.uint16)
            inp.append(pos[:d])
            K.append(bf16_bits_to_f32(kv[:d]))
            V.append(bf16_bits_to_f32(kv[d:]))
        return np.array(inp), np.array(K), np.array(V)


@pytest.mark.parametrize("kernel", [1, 2], ids=["column-split", "warp-per-position"])
@pytest.mark.parametrize("B,S,d", [(8, 64, 128), (40, 128, 1024), (12, 256, 2048), (6, 64, 4096)])
def test_attention_on_compact_pages(torch_cuda, compact, B, S, d, kernel):
    compact.set_option(mli.OPT_ATTN_KERNEL, kernel)
    try:
        _attention_on_compact_pages(torch_cuda, compact, B, S, d)
    finally:
        compact.set_option(mli.OPT_ATTN_KERNEL, 0)


def _attention_on_compact_pages(torch_cuda, compact, B, S, d):
    torch = torch_cuda
    rng = np.random.default_rng(50 + B + d)
    L = rng.integers(1, S, size=B).astype(np.int32)
    L[rng.random(B) < 0.2] = 0
    L[0] = S - 1
    case = CompactCase(3, B, S, d, L)
    pool, tab = case.device(torch)
    q = ((rng.random((B, d), dtype=np.float32) - 0.5) * 2.0 * np.sqrt(12.0 / d)).astype(np.float32)
    out = torch.full((B, d), 7.0, device="cuda")
    compact.call("mli_decode_attention_paged", dev(torch, q), tab, dev(torch, L), out, None, B, S, d)
    compact.synchronize()
    a = out.cpu().numpy()
    pool_np = pool.cpu().numpy()
    want = np.zeros((B, d), np.float64)
    for r in range(B):
        if L[r] == 0:
            continue
        _, K, V = case.rows(pool_np, r, int(L[r]))
        s = K.astype(np.float64) @ q[r].astype(np.float64) / np.sqrt(np.float64(d))
        p = np.exp(s - s.max())
        p /= p.sum()
        want[r] = p @ V.astype(np.float64)
    assert np.all(a[L == 0] == 0.0)
    assert H.rel_err(a, want) < 1e-4, f"attention on bf16 pages: rel err {H.rel_err(a, want):.2e}"
    # the [B,S] probabilities are not offered in this format
    probs = torch.zeros((B, S), device="cuda")
    with pytest.raises(mli.MliError):
        compact.call("mli_decode_attention_paged", dev(torch, q), tab, dev(torch, L), out, probs, B, S, d)


@pytest.mark.parametrize("B,S,d", [(8, 64, 128), (33, 128, 256), (64, 128, 1024), (10, 64, 2048)])
def test_projection_stores_rounded_rows(torch_cuda, compact, B, S, d):
    torch = torch_cuda
    V = 1024
    rng = np.random.default_rng(70 + B + d)
    L = rng.integers(2, S, size=B).astype(np.int32)
    L[rng.random(B) < 0.2] = 0
    case = CompactCase(5, B, S, d, L)
    w = H.make_weights(11, d, V, S, "Z")
    dw = {k: dev(torch, v) for k, v in w.items()}
    cand = np.flatnonzero(L > 0)
    n_new = max(1, len(cand) // 2)
    new_idx = np.zeros(B, np.int32)
    new_idx[:n_new] = rng.permutation(cand)[:n_new]
    pool, tab = case.device(torch)
    before = pool.cpu().numpy().copy()
    qd = torch.full((B, d), 3.0, device="cuda")
    dL, dnew = dev(torch, L), dev(torch, new_idx)
    compact.call("mli_prefill_kv_paged", tab, dnew, dL, dw["wk"], dw["wv"], n_new, B, S, d)
    compact.call("mli_qkv_latest_paged", tab, dL, dw["wk"], dw["wq"], dw["wv"], qd, B, S, d)
    compact.synchronize()
    after = pool.cpu().numpy()
    qh = qd.cpu().numpy()
    new_rows = set(new_idx[:n_new].tolist())
    worst = 0.0
    for r in range(B):
        if L[r] == 0:
            assert np.all(qh[r] == 3.0)
            continue
        inp, K, Vv = case.rows(after, r, int(L[r]))
        inp0, K0, V0 = case.rows(before, r, int(L[r]))
        assert np.array_equal(inp, inp0), "input embeddings must not be touched"
        x = inp.astype(np.float64)
        wantK, wantV = x @ w["wk"].astype(np.float64), x @ w["wv"].astype(np.float64)
        written = range(int(L[r])) if r in new_rows else [int(L[r]) - 1]
        for j in range(int(L[r])):
            if j in written:
                for got, want in ((K[j], wantK[j]), (Vv[j], wantV[j])):
                    # a bf16 value within one ulp of the exact result (2^-8 relative), plus the fp32
                    # error of the projection itself, which is relative to the row's scale (1e-5)
                    err = np.abs(got.astype(np.float64) - want)
                    scale = float(np.max(np.abs(want)))
                    bound = np.abs(want) * 2.0 ** -8 + 1e-5 * scale
                    assert np.all(err <= bound), f"row {r} pos {j}: stored bf16 is not the rounded projection"
                    worst = max(worst, float(np.max(err)) / scale)
            else:
                assert np.array_equal(K[j], K0[j]) and np.array_equal(Vv[j], V0[j]), "untouched position changed"
        wantq = x[-1] @ w["wq"].astype(np.float64)
        assert H.rel_err(qh[r], wantq) < 1e-5
    assert worst <= 2.0 ** -8 + 1e-5


def test_compact_engine_runs_and_agrees_on_first_tokens(torch_cuda, compact):
    torch = torch_cuda
    case = dict(B=64, S=128, d=256, V=1024, n_blocks=300, n_req=160, lo=1, hi=64, R=1)
    w = H.make_weights(41, case["d"], case["V"], case["S"], "Z")
    offs, toks = H.make_prompts(43, case["n_req"], case["lo"], case["hi"])
    a, order_a, st_a = run_mli_engine(compact, torch, case, w, offs, toks, compat=0)
    compact.set_option(mli.OPT_KV_FORMAT, 0)
    b, order_b, st_b = run_mli_engine(compact, torch, case, w, offs, toks, compat=0)
    compact.set_option(mli.OPT_KV_FORMAT, 1)
    assert st_a.n_finished == st_b.n_finished == case["n_req"]
    agree = 0
    for i in range(case["n_req"]):
        n0 = offs[i + 1] - offs[i]
        assert np.array_equal(a[i][:n0], toks[offs[i]:offs[i + 1]]), "prompt must be preserved"
        assert n0 < len(a[i]) <= case["S"]
        agree += int(a[i][n0] == b[i][n0])
    assert agree >= 0.9 * case["n_req"], f"first generated token agrees for only {agree}/{case['n_req']} requests"


def test_compact_format_needs_the_tensor_core_mode(torch_cuda, compact):
    torch = torch_cuda
    compact.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    case = dict(B=4, S=64, d=128, V=1024, n_blocks=16, n_req=4, lo=1, hi=20, R=1)
    w = H.make_weights(41, case["d"], case["V"], case["S"], "Z")
    offs, toks = H.make_prompts(43, case["n_req"], case["lo"], case["hi"])
    with pytest.raises(mli.MliError):
        run_mli_engine(compact, torch, case, w, offs, toks, compat=0)
    compact.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
