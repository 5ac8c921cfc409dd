"""Stage-level parity of the NON-paged path (BASELINE configs[0]) against the
reference's own CUDA kernels compiled into oracle/_ref:

    mli_dense_encoder   vs launch_inference_optimized_encoder_kernel   src/kernels/encoder.cu:56-92
    mli_self_attention  vs inference_self_attention                    src/kernels/self_attention_inference_optimized.cu:27-383
    mli_dense_decoder   vs launch_decoder                              src/kernels/decoder.cu:25-112
    mli_dense_forward   vs InferenceModel::forward (the three above)   src/inference_model.cpp:14-39

Shapes follow tests/self_attention_inference_optimized_test.cpp:6-190 (incl. rows of length 0) and the
BASELINE configs[0] shape (B=32, d=256, S=256).  Bar: K^T / V caches, q, logits, tokens, lengths and the
next input embedding bit-exact (both sides are k-ascending fp32 FMA chains); softmax probabilities and
the attention result within rel 1e-4 (expf and the P.V summation order differ).
"""
import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli

pytestmark = pytest.mark.gpu

SHAPES = [(6, 64, 64, 1024), (32, 256, 256, 1024), (5, 128, 132, 1000), (16, 96, 512, 1024)]


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def setup(torch, ref, B, S, d, V, dist, seed):
    """a batch in mid-flight: every row embedded and prefilled by the REFERENCE, then a subset of rows is
    re-admitted as "new" with fresh prompts"""
    rng = np.random.default_rng(seed)
    w = H.make_weights(seed + 1, d, V, S, dist, eof_ratio=1.001)
    dw = {k: dev(torch, v) for k, v in w.items()}
    L = rng.integers(1, S - 1, size=B).astype(np.int32)
    L[rng.random(B) < 0.2] = 0
    if B > 2:
        L[1] = 0
        L[2] = S - 1
    inp = rng.integers(0, min(V, 1023), size=(B, S)).astype(np.int32)
    all_idx = dev(torch, np.arange(B, dtype=np.int32))
    dL, dinp = dev(torch, L), dev(torch, inp)
    x = torch.zeros((B, S, d), device="cuda")
    H.check_ref(ref.ref_dense_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(dinp), H.p(x), H.p(dL), H.p(all_idx),
                                      B, S, d, B))
    kt = torch.zeros((B, d, S), device="cuda")
    vc = torch.zeros((B, S, d), device="cuda")
    scratch = [torch.zeros((B, d), device="cuda"), torch.zeros((B, S), device="cuda"),
               torch.zeros((B, d), device="cuda")]
    H.check_ref(ref.ref_self_attention(H.p(x), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]), H.p(dw["wv"]), H.p(all_idx),
                                       H.p(kt), H.p(vc), H.p(scratch[0]), H.p(scratch[1]), H.p(scratch[2]),
                                       B, B, S, d, d))
    cand = np.flatnonzero(L > 0)
    n_new = max(1, len(cand) // 3)
    new_idx = np.zeros(B, np.int32)
    new_idx[:n_new] = rng.permutation(cand)[:n_new]
    inp2 = inp.copy()
    inp2[new_idx[:n_new]] = rng.integers(0, min(V, 1023), size=(n_new, S))
    return dict(w=w, dw=dw, L=L, inp=dev(torch, inp2), x=x, kt=kt, vc=vc, new_idx=dev(torch, new_idx), n_new=n_new)


@pytest.mark.parametrize("B,S,d,V", SHAPES)
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_dense_stages_match_reference(torch_cuda, ctx, ref, B, S, d, V, dist):
    torch = torch_cuda
    st = setup(torch, ref, B, S, d, V, dist, 700 + B + d)
    dw = st["dw"]
    # ---- encoder: new rows get E[tok] + P ----
    xs = [st["x"].clone(), st["x"].clone()]
    dL = dev(torch, st["L"])
    ctx.call("mli_dense_encoder", dw["emb"], dw["pos"], st["inp"], xs[0], dL, st["new_idx"], B, S, d, st["n_new"])
    ctx.synchronize()
    H.check_ref(ref.ref_dense_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(st["inp"]), H.p(xs[1]), H.p(dL),
                                      H.p(st["new_idx"]), B, S, d, st["n_new"]))
    assert torch.equal(xs[0], xs[1]), "inp_embedding differs from the reference encoder"
    # ---- self attention: prefill of the new rows + latest QKV + attention ----
    outs = []
    for side in range(2):
        kt, vc = st["kt"].clone(), st["vc"].clone()
        q = torch.full((B, d), -3.0, device="cuda")
        qkt = torch.full((B, S), -3.0, device="cuda")
        att = torch.full((B, d), -3.0, device="cuda")
        if side == 0:
            ctx.call("mli_self_attention", xs[0], dL, dw["wk"], dw["wq"], dw["wv"], st["new_idx"], kt, vc, q, qkt,
                     att, st["n_new"], B, S, d, d)
            ctx.synchronize()
        else:
            H.check_ref(ref.ref_self_attention(H.p(xs[1]), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]), H.p(dw["wv"]),
                                               H.p(st["new_idx"]), H.p(kt), H.p(vc), H.p(q), H.p(qkt), H.p(att),
                                               st["n_new"], B, S, d, d))
        outs.append((kt, vc, q, qkt, att))
    live = torch.from_numpy(st["L"] > 0).cuda()
    assert torch.equal(outs[0][0], outs[1][0]), "kt_cache differs"
    assert torch.equal(outs[0][1], outs[1][1]), "v_cache differs"
    assert torch.equal(outs[0][2][live], outs[1][2][live]), "q_output differs"
    p_mine, p_ref = outs[0][3][live].cpu().numpy(), outs[1][3][live].cpu().numpy()
    assert H.rel_err(p_mine, p_ref) < 1e-4, "softmax probabilities differ"
    # positions past a row's length hold zeros on both sides (self_attention_inference_optimized.cu:226-241)
    Lh = st["L"][st["L"] > 0]
    for i in range(len(Lh)):
        assert not p_mine[i, Lh[i]:].any() and not p_ref[i, Lh[i]:].any()
    a_mine, a_ref = outs[0][4][live].cpu().numpy(), outs[1][4][live].cpu().numpy()
    assert H.rel_err(a_mine, a_ref) < 1e-4, "attention_result differs"
    # ---- decoder on the REFERENCE's attention result: logits, argmax, lengths, next embedding ----
    att = outs[1][4].clone()
    att[~live] = 0.0
    res = []
    for side in range(2):
        x = xs[side].clone()
        ln = dev(torch, st["L"])
        score = torch.zeros((B, V), device="cuda")
        dec = torch.full((B,), -7, dtype=torch.int32, device="cuda")
        if side == 0:
            ctx.call("mli_dense_decoder", att, dw["emb"], score, dw["pos"], x, ln, dec, B, V, S, d)
            ctx.synchronize()
        else:
            H.check_ref(ref.ref_dense_decoder(H.p(att), H.p(dw["emb"]), H.p(score), H.p(dw["pos"]), H.p(x), H.p(ln),
                                              H.p(dec), B, V, S, d))
        res.append((x, ln, score, dec))
    assert torch.equal(res[0][3], res[1][3]), "tokens differ from the reference decoder"
    assert torch.equal(res[0][1], res[1][1]), "lengths differ from the reference decoder"
    assert torch.equal(res[0][2][live], res[1][2][live]), "logits differ from the reference decoder"
    assert torch.equal(res[0][0], res[1][0]), "next input embedding differs"


@pytest.mark.parametrize("B,S,d,V", SHAPES[:3])
@pytest.mark.parametrize("dist", ["R", "Z"])
def test_dense_forward_matches_reference_chain(torch_cuda, ctx, ref, B, S, d, V, dist):
    """three consecutive InferenceModel::forward steps (new rows only in the first)"""
    torch = torch_cuda
    st = setup(torch, ref, B, S, d, V, dist, 900 + B + d)
    dw = st["dw"]
    state = []
    for side in range(2):
        x, kt, vc = st["x"].clone(), st["kt"].clone(), st["vc"].clone()
        ln = dev(torch, st["L"])
        decs = []
        for step in range(3):
            n_new = st["n_new"] if step == 0 else 0
            dec = torch.full((B,), -7, dtype=torch.int32, device="cuda")
            if side == 0:
                ctx.call("mli_dense_forward", st["inp"], ln, st["new_idx"], dec, n_new, dw["emb"], dw["pos"], dw["wk"],
                         dw["wq"], dw["wv"], x, kt, vc, None, None, B, S, d, V)
                ctx.synchronize()
            else:
                q = torch.zeros((B, d), device="cuda")
                qkt = torch.zeros((B, S), device="cuda")
                att = torch.zeros((B, d), device="cuda")
                score = torch.zeros((B, V), device="cuda")
                H.check_ref(ref.ref_dense_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(st["inp"]), H.p(x), H.p(ln),
                                                  H.p(st["new_idx"]), B, S, d, n_new))
                H.check_ref(ref.ref_self_attention(H.p(x), H.p(ln), H.p(dw["wk"]), H.p(dw["wq"]), H.p(dw["wv"]),
                                                   H.p(st["new_idx"]), H.p(kt), H.p(vc), H.p(q), H.p(qkt), H.p(att),
                                                   n_new, B, S, d, d))
                H.check_ref(ref.ref_dense_decoder(H.p(att), H.p(dw["emb"]), H.p(score), H.p(dw["pos"]), H.p(x),
                                                  H.p(ln), H.p(dec), B, V, S, d))
            decs.append(dec.clone())
        state.append((decs, ln, kt, vc, x))
    for step in range(3):
        assert torch.equal(state[0][0][step], state[1][0][step]), f"tokens of step {step} differ"
    assert torch.equal(state[0][1], state[1][1]), "lengths differ"
    assert torch.equal(state[0][2], state[1][2]) and torch.equal(state[0][3], state[1][3]), "caches differ"
    assert torch.equal(state[0][4], state[1][4]), "input embeddings differ"
