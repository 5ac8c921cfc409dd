"""Pins the CPU oracle (oracle/oracle.c) to the reference's own CUDA kernels on a GPU box.

The reference has no golden vectors, so this is the primary pin: same seeded inputs
through oracle/_ref (the unmodified reference) and through the C restatement.  Contractions are
k-ascending FMA chains on both sides -> bit-exact; softmax involves expf -> 2e-6.
"""
import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("B,S,d,V", [(6, 64, 64, 1024), (9, 128, 256, 1024), (4, 128, 516, 1000)])
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_oracle_equals_reference_cuda(torch_cuda, ref, B, S, d, V, dist):
    torch = torch_cuda
    orc = H.load_oracle()
    rng = np.random.default_rng(500 + B + d)
    L = rng.integers(1, S - 1, size=B).astype(np.int32)
    L[0] = 0
    case = H.PagedCase(41, B, S, d, L, dist)
    w = H.make_weights(43, d, V, S, dist, eof_ratio=1.3)
    new_idx = np.zeros(B, np.int32)
    cand = np.flatnonzero(L > 0)
    n_new = len(cand) // 2 + 1
    new_idx[:n_new] = cand[:n_new]
    inp = rng.integers(0, min(V, 1023), size=(B, S)).astype(np.int32)
    # ---- oracle (host) ----
    hpool, htab = case.host()
    hq = np.zeros((B, d), np.float32)
    hqkt = np.zeros((B, S), np.float32)
    hattn = np.zeros((B, d), np.float32)
    orc.orc_paged_encoder(H.p(w["emb"]), H.p(w["pos"]), H.p(inp), H.p(htab), H.p(L), H.p(new_idx), B, S, d, n_new)
    orc.orc_paged_attention(H.p(htab), H.p(L), H.p(w["wk"]), H.p(w["wq"]), H.p(w["wv"]), H.p(new_idx),
                            H.p(hq), H.p(hqkt), H.p(hattn), n_new, B, S, d)
    # ---- reference CUDA ----
    dw = {k: dev(torch, v) for k, v in w.items()}
    pool, tab = case.device(torch)
    dL, dnew, dinp = dev(torch, L), dev(torch, new_idx), dev(torch, inp)
    dq = torch.zeros((B, d), device="cuda")
    dqkt = torch.zeros((B, S), device="cuda")
    dattn = torch.zeros((B, d), device="cuda")
    H.check_ref(ref.ref_paged_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(dinp), H.p(tab), H.p(dL),
                                      H.p(dnew), B, S, d, n_new))
    H.check_ref(ref.ref_paged_attention(H.p(tab), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]), H.p(dw["wv"]),
                                        H.p(dnew), H.p(dq), H.p(dqkt), H.p(dattn), n_new, B, S, d, 0))
    assert np.array_equal(hpool, pool.cpu().numpy()), "oracle pages (embeddings, K, V) not bit-exact"
    assert np.array_equal(hq, dq.cpu().numpy()), "oracle q_output not bit-exact"
    assert np.abs(hqkt - dqkt.cpu().numpy()).max() < 2e-6, "oracle softmax differs"
    assert H.rel_err(hattn, dattn.cpu().numpy()) < 2e-6, "oracle attention differs"
    # ---- decoder ----
    hscore = np.zeros((B, V), np.float32)
    hL, hdec = L.copy(), np.zeros((B, 1), np.int32)
    orc.orc_logits(H.p(hattn), H.p(w["emb"]), H.p(hscore), B, V, d)
    dscore = torch.zeros((B, V), device="cuda")
    ddec = torch.zeros((B, 1), dtype=torch.int32, device="cuda")
    dattn_h = dev(torch, hattn)   # same attention input on both sides
    H.check_ref(ref.ref_paged_decoder(H.p(dattn_h), H.p(dw["emb"]), H.p(dscore), H.p(dw["pos"]),
                                      H.p(tab), H.p(dL), H.p(ddec), B, V, S, d, 1, 0, 0))
    orc.orc_paged_decoder(H.p(hscore), H.p(hdec), H.p(hL), H.p(htab), H.p(w["pos"]), H.p(w["emb"]),
                          B, V, S, d, 1, 0)
    assert np.array_equal(hscore, dscore.cpu().numpy()), "oracle logits not bit-exact"
    assert np.array_equal(hdec, ddec.cpu().numpy()), "oracle tokens differ"
    assert np.array_equal(hL, dL.cpu().numpy()), "oracle lengths differ"
    assert np.array_equal(hpool, pool.cpu().numpy()), "oracle next embedding differs"


@pytest.mark.parametrize("dist", ["R", "Z"])
def test_oracle_engines_equal_reference_engines(torch_cuda, ref, dist):
    from test_gpu_forward_engine import run_ref_engine
    case = dict(B=8, S=128, d=128, V=1024, n_blocks=36, n_req=24, lo=20, hi=64)
    w = H.make_weights(31, case["d"], case["V"], case["S"], dist)
    offs, toks = H.make_prompts(33, case["n_req"], case["lo"], case["hi"])
    theirs, torder, _ = run_ref_engine(ref, "paged", case, w, offs, toks, 0)
    rc, mine, order, st = H.run_oracle_engine("paged", case, w, offs, toks, fix=0)
    assert rc == 0 and order.tolist() == torder.tolist()
    assert all(np.array_equal(mine[i], theirs[i]) for i in range(case["n_req"]))
    theirs, torder, _ = run_ref_engine(ref, "dense", case, w, offs, toks)
    rc, mine, order, st = H.run_oracle_engine("dense", case, w, offs, toks)
    assert rc == 0 and order.tolist() == torder.tolist()
    assert all(np.array_equal(mine[i], theirs[i]) for i in range(case["n_req"]))
