"""CPU-only checks of the oracle (oracle/oracle.c): against an independent float64 numpy model,
against the reference's scheduler semantics (tests/paged_item_storage_test.cpp,
tests/item_storage_test.cpp), and against the golden vectors generated from the reference's CUDA
path on a B200 (tests/golden/make_golden.py)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import harness as H

orc = H.load_oracle()
GOLDEN = Path(__file__).resolve().parent / "golden"


def np_attention(case, pool, q, L):
    B, S, d = case.B, case.S, case.d
    out = np.zeros((B, d))
    K = case.gather(pool, 1).astype(np.float64)
    V = case.gather(pool, 2).astype(np.float64)
    for r in range(B):
        if L[r] == 0:
            continue
        s = K[r, :L[r]] @ q[r].astype(np.float64) / np.sqrt(d)
        p = np.exp(s - s.max())
        p /= p.sum()
        out[r] = p @ V[r, :L[r]]
    return out


@pytest.mark.parametrize("dist", ["R", "Z"])
def test_stage_functions_against_float64_model(dist):
    B, S, d, V = 6, 64, 64, 1024
    rng = np.random.default_rng(5)
    L = np.array([0, 1, 17, 33, 62, 63], np.int32)
    case = H.PagedCase(3, B, S, d, L, dist)
    w = H.make_weights(4, d, V, S, dist)
    pool, tab = case.host()
    inp = rng.integers(0, 1023, size=(B, S)).astype(np.int32)
    idx = np.arange(B, dtype=np.int32)
    orc.orc_paged_encoder(H.p(w["emb"]), H.p(w["pos"]), H.p(inp), H.p(tab), H.p(L), H.p(idx), B, S, d, B)
    x = case.gather(pool, 0)
    for r in range(B):
        for j in range(L[r]):
            assert np.array_equal(x[r, j], w["emb"][inp[r, j]] + w["pos"][j])
    q = np.zeros((B, d), np.float32)
    qkt = np.zeros((B, S), np.float32)
    attn = np.full((B, d), 9.0, np.float32)
    orc.orc_paged_attention(H.p(tab), H.p(L), H.p(w["wk"]), H.p(w["wq"]), H.p(w["wv"]), H.p(idx),
                            H.p(q), H.p(qkt), H.p(attn), B, B, S, d)
    K, Vv = case.gather(pool, 1), case.gather(pool, 2)
    for r in range(B):
        for j in range(L[r]):
            assert H.rel_err(K[r, j], x[r, j].astype(np.float64) @ w["wk"].astype(np.float64)) < 1e-5
            assert H.rel_err(Vv[r, j], x[r, j].astype(np.float64) @ w["wv"].astype(np.float64)) < 1e-5
        if L[r]:
            assert H.rel_err(q[r], x[r, L[r] - 1].astype(np.float64) @ w["wq"].astype(np.float64)) < 1e-5
    if dist == "Z":     # with U(0,1] inputs softmax is one-hot and fp32 scores cannot match fp64 ones
        assert H.rel_err(attn, np_attention(case, pool, q, L)) < 1e-4
    assert np.all(attn[0] == 0) and np.all(qkt[0] == 0)
    for r in range(B):
        assert np.all(qkt[r, L[r]:] == 0)
        if L[r]:
            assert abs(qkt[r].sum() - 1) < 1e-5


def bitrev8(x):
    return int(f"{x:08b}"[::-1], 2)


def test_argmax_device_rule():
    """decoder.cu:146-172: the shared-memory tree keeps the LOWER thread on ties at every level
    (gap 128, 64, ... 1), so two equal maxima held by threads ta, tb are decided by the lowest bit
    in which ta and tb differ: the winner is min over maxima of (bitreverse8(index % 256), index).
    (SURVEY App. A Q4 says "min index mod 256"; simulating the kernel shows the bit-reversed order.)"""
    s = np.zeros(1024, np.float32)
    s[[300, 513]] = 1.0        # threads 44 (0b00101100) and 1 (0b00000001): bit 0 decides -> 44
    assert orc.orc_argmax_device_rule(H.p(s), 1024) == 300
    s[5] = 1.0                 # thread 5 has bit 0 set: still thread 44
    assert orc.orc_argmax_device_rule(H.p(s), 1024) == 300
    s[556] = 1.0               # 556 % 256 = 44 as well, same thread keeps its FIRST maximum (300)
    assert orc.orc_argmax_device_rule(H.p(s), 1024) == 300
    s[256] = 1.0               # thread 0 beats everyone
    assert orc.orc_argmax_device_rule(H.p(s), 1024) == 256
    s[700] = 2.0
    assert orc.orc_argmax_device_rule(H.p(s), 1024) == 700
    rng = np.random.default_rng(0)
    for _ in range(50):        # closed form == literal simulation on random tie sets
        s = np.zeros(1000, np.float32)
        ties = rng.choice(1000, size=rng.integers(2, 9), replace=False)
        s[ties] = 3.0
        want = min(ties.tolist(), key=lambda i: (bitrev8(i % 256), i))
        assert orc.orc_argmax_device_rule(H.p(s), 1000) == want


def test_decoder_length_and_eof_rules():
    """decoder.cu:173-188 / tests/decoder_test.cpp MaxLengthTest"""
    B, S, d, V = 4, 64, 64, 1024
    L = np.array([0, 10, 62, 63], np.int32)
    case = H.PagedCase(3, B, S, d, np.array([0, 11, 63, 63], np.int32), "R")
    pool, tab = case.host()
    w = H.make_weights(4, d, V, S, "R")
    score = np.zeros((B, V), np.float32)
    score[1, 7] = 1.0
    score[2, 9] = 1.0
    score[3, 1023] = 1.0
    dec = np.zeros((B, 1), np.int32)
    LL = L.copy()
    before = pool.copy()
    orc.orc_paged_decoder(H.p(score), H.p(dec), H.p(LL), H.p(tab), H.p(w["pos"]), H.p(w["emb"]),
                          B, V, S, d, 1, 0)
    assert dec.ravel().tolist() == [-1, 7, 9, 1023]
    assert LL.tolist() == [0, 11, 63, 0]         # 62+1 = 63 < 64 continues; EOF resets
    assert np.array_equal(case.view(pool, 1, 10, 0), w["emb"][7] + w["pos"][10])
    assert np.array_equal(case.view(pool, 2, 62, 0), w["emb"][9] + w["pos"][62])
    LL2 = np.array([0, 0, 63, 0], np.int32)      # 63+1 >= 64 -> reset, nothing written
    snap = pool.copy()
    orc.orc_paged_decoder(H.p(score), H.p(dec), H.p(LL2), H.p(tab), H.p(w["pos"]), H.p(w["emb"]),
                          B, V, S, d, 1, 0)
    assert LL2.tolist() == [0, 0, 0, 0] and np.array_equal(snap, pool)
    del before


CASES = [
    dict(B=4, S=64, d=32, V=1024, n_blocks=16, n_req=10, lo=1, hi=40),
    dict(B=8, S=128, d=32, V=1024, n_blocks=36, n_req=24, lo=20, hi=64),   # pre-emption
    dict(B=6, S=64, d=32, V=1024, n_blocks=6 * 4, n_req=6, lo=1, hi=30),   # InsertAllItemsTest shape
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_paged_engine_fixed_equals_dense_engine(case, dist):
    """pre-emption is recompute: tokens must not depend on the schedule"""
    w = H.make_weights(31, case["d"], case["V"], case["S"], dist)
    offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
    rc, paged, order, st = H.run_oracle_engine("paged", case, w, offs, toks, fix=1)
    rc2, dense, _, st2 = H.run_oracle_engine("dense", case, w, offs, toks)
    assert rc == 0 and rc2 == 0 and st.n_finished == case["n_req"] == st2.n_finished
    for i in range(case["n_req"]):
        assert np.array_equal(paged[i], dense[i])
        n0 = offs[i + 1] - offs[i]
        assert np.array_equal(paged[i][:n0], toks[offs[i]:offs[i + 1]])
        assert len(paged[i]) == case["S"] or paged[i][-1] == H.EOF
    if case["n_blocks"] == 36:
        assert st.preemptions > 0, "this case is sized to force tail pre-emption"


def test_quirk_q1_stale_lengths_reproduced():
    """paged_item_storage.cpp:73-75,:114-118: with a free row present every step, every in-flight
    row re-emits its first token until the host-side list reaches S"""
    case = dict(B=4, S=64, d=32, V=1024, n_blocks=16, n_req=3, lo=5, hi=20)   # 3 requests, 4 rows
    w = H.make_weights(31, case["d"], case["V"], case["S"], "Z")
    offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
    rc, q1, _, _ = H.run_oracle_engine("paged", case, w, offs, toks, fix=0)
    rc, ok, _, _ = H.run_oracle_engine("paged", case, w, offs, toks, fix=1)
    for i in range(3):
        n0 = offs[i + 1] - offs[i]
        gen = q1[i][n0:]
        assert len(set(gen.tolist())) == 1 and gen[0] == ok[i][n0]
        assert len(set(ok[i][n0:].tolist())) > 1


def test_admission_rule_q6():
    """needs free >= 4 and free >= ceil((len+R)/16); allocates max(that, 4) pages
    (paged_item_storage.cpp:84-113; tests/paged_item_storage_test.cpp InsertNewItemsTest)"""
    case = dict(B=8, S=128, d=32, V=1024, n_blocks=9, n_req=3, lo=70, hi=70)   # 5 pages each
    w = H.make_weights(1, 32, 1024, 128, "Z")
    offs, toks = H.make_prompts(2, 3, 70, 70)
    rc, res, order, st = H.run_oracle_engine("paged", case, w, offs, toks, fix=1)
    assert rc == 0 and st.n_finished == 3
    # only one 5-page request fits in 9 pages at a time (second needs 5 > 4 left) -> strictly serial
    assert order.tolist() == [0, 1, 2]
    assert st.steps == sum(128 - 70 for _ in range(3)) or st.generated_tokens <= 3 * 58


@pytest.mark.skipif(not list(GOLDEN.glob("ref_cuda_*.npz")), reason="no golden vectors committed yet")
@pytest.mark.parametrize("path", sorted(GOLDEN.glob("ref_cuda_*.npz")), ids=lambda p: p.stem)
def test_oracle_against_reference_cuda_golden(path):
    with np.load(path, allow_pickle=False) as z:
        g = {k: (z[k].item() if z[k].ndim == 0 else np.ascontiguousarray(z[k])) for k in z.files}  # keep arrays alive: H.p() takes raw pointers
    kind = str(g["kind"])
    if kind == "stages":
        B, S, d, V = (int(g[k]) for k in ("B", "S", "d", "V"))
        dist = str(g["dist"])
        L = g["lengths"]
        case = H.PagedCase(int(g["case_seed"]), B, S, d, L, dist)
        w = H.make_weights(int(g["w_seed"]), d, V, S, dist, eof_ratio=float(g["eof_ratio"]))
        pool, tab = case.host()
        q = np.zeros((B, d), np.float32)
        qkt = np.zeros((B, S), np.float32)
        attn = np.zeros((B, d), np.float32)
        n_new = int(g["n_new"])
        orc.orc_paged_encoder(H.p(w["emb"]), H.p(w["pos"]), H.p(g["inp"]), H.p(tab), H.p(L),
                              H.p(g["new_idx"]), B, S, d, n_new)
        orc.orc_paged_attention(H.p(tab), H.p(L), H.p(w["wk"]), H.p(w["wq"]), H.p(w["wv"]),
                                H.p(g["new_idx"]), H.p(q), H.p(qkt), H.p(attn), n_new, B, S, d)
        assert np.array_equal(pool, g["pool_after_attention"])
        assert np.array_equal(q, g["q_output"])
        assert np.abs(qkt - g["softmax"]).max() < 2e-6
        assert H.rel_err(attn, g["attention_result"]) < 2e-6
        score = np.zeros((B, V), np.float32)
        orc.orc_logits(H.p(g["attention_result"]), H.p(w["emb"]), H.p(score), B, V, d)
        assert np.array_equal(score, g["logits"])
        LL, dec = L.copy(), np.zeros((B, 1), np.int32)
        orc.orc_paged_decoder(H.p(score), H.p(dec), H.p(LL), H.p(tab), H.p(w["pos"]), H.p(w["emb"]),
                              B, V, S, d, 1, 0)
        assert np.array_equal(dec, g["tokens"]) and np.array_equal(LL, g["lengths_after"])
    else:
        case = {k: int(g[k]) for k in ("B", "S", "d", "V", "n_blocks", "n_req", "lo", "hi")}
        dist = str(g["dist"])
        w = H.make_weights(int(g["w_seed"]), case["d"], case["V"], case["S"], dist)
        offs, toks = H.make_prompts(int(g["p_seed"]), case["n_req"], case["lo"], case["hi"])
        for name, ek, fix in (("paged", "paged", 0), ("dense", "dense", 0)):
            rc, res, order, st = H.run_oracle_engine(ek, case, w, offs, toks, fix=fix)
            assert rc == 0
            assert order.tolist() == g[f"{name}_order"].tolist()
            fo, ft = g[f"{name}_offsets"], g[f"{name}_tokens"]
            for k, rid in enumerate(order):
                assert np.array_equal(res[int(rid)], ft[fo[k]:fo[k + 1]])


def test_oracle_policy_flags():
    """the two opt-in scheduling policies the product engine adds (mli_engine_cfg.max_new_tokens /
    max_prefill_positions) as restated in the oracle: the token cap truncates every request's generation,
    the admission throttle changes WHEN requests are admitted but -- with corrected lengths -- never a token"""
    cfg = dict(B=6, S=64, d=32, V=1024, n_blocks=40, R=1)
    w = H.make_weights(11, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(13, 15, 3, 30)
    plen = np.diff(offs)
    rc, base, order, st = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1)
    assert rc == 0 and st.n_finished == 15
    rc, capped, _, st_c = H.run_oracle_engine("paged", dict(cfg, max_new=5), w, offs, toks, fix=1)
    assert rc == 0 and st_c.n_finished == 15
    for i in range(15):
        assert len(capped[i]) - plen[i] <= 5
        assert np.array_equal(capped[i], base[i][:len(capped[i])])
    assert st_c.steps < st.steps
    rc, thr, _, st_t = H.run_oracle_engine("paged", dict(cfg, max_prefill=24), w, offs, toks, fix=1)
    assert rc == 0 and st_t.n_finished == 15
    for i in range(15):
        assert np.array_equal(thr[i], base[i])
    assert st_t.steps >= st.steps      # admissions are spread over more iterations


def test_oracle_chunked_prefill():
    """chunked prefill as restated in the oracle: the same token lists as the unchunked job, first tokens later"""
    cfg = dict(B=4, S=128, d=32, V=1024, n_blocks=48, R=1, max_new=6)
    w = H.make_weights(21, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(23, 9, 30, 100)
    rc, base, _, st = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1)
    rc2, chunked, _, st_c = H.run_oracle_engine("paged", dict(cfg, chunk=32), w, offs, toks, fix=1)
    assert rc == 0 and rc2 == 0 and st.n_finished == st_c.n_finished == 9
    for i in range(9):
        assert np.array_equal(base[i], chunked[i])
    assert st_c.steps > st.steps
    # needs corrected lengths
    rc3, *_ = H.run_oracle_engine("paged", dict(cfg, chunk=32), w, offs, toks, fix=0)
    assert rc3 == -1


def test_oracle_stops_where_the_reference_would_spin():
    """a request that outgrows the whole pool can never be admitted again after it pre-empted itself: the
    reference's loop spins (paged_item_storage.cpp:84-113 keeps admitting nothing), the oracle returns -5;
    with a token cap that keeps every request inside the pool the same job completes"""
    cfg = dict(B=4, S=192, d=64, V=1024, n_blocks=10, R=1, n_req=6, lo=3, hi=12)
    w = H.make_weights(81, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(83, cfg["n_req"], cfg["lo"], cfg["hi"])
    rc, res, order, st = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1, max_steps=20000)
    assert rc == -5 and st.n_finished < cfg["n_req"] and st.preemptions >= 1
    rc, res, order, st = H.run_oracle_engine("paged", dict(cfg, max_new=20), w, offs, toks, fix=1, max_steps=20000)
    assert rc == 0 and st.n_finished == cfg["n_req"]
