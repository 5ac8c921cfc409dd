"""Generates tests/golden/ref_cuda_*.npz from the UNMODIFIED reference's CUDA path
(oracle/_ref/libmli_ref.so) on a GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy into tests/golden/

The reference ships no golden vectors of its own; these pin the CPU oracle in the
CPU-only suite (tests/test_cpu_oracle.py::test_oracle_against_reference_cuda_golden).
Inputs are regenerated from the stored seeds by tests/harness.py, so the files stay small.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import harness as H  # noqa: E402


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def stages(out, name, B, S, d, V, dist, case_seed, w_seed, eof_ratio=1.3):
    ref = H.load_ref()
    rng = np.random.default_rng(case_seed + 1000)
    L = rng.integers(1, S - 1, size=B).astype(np.int32)
    L[0] = 0
    case = H.PagedCase(case_seed, B, S, d, L, dist)
    w = H.make_weights(w_seed, d, V, S, dist, eof_ratio=eof_ratio)
    cand = np.flatnonzero(L > 0)
    n_new = len(cand) // 2 + 1
    new_idx = np.zeros(B, np.int32)
    new_idx[:n_new] = cand[:n_new]
    inp = rng.integers(0, min(V, 1023), size=(B, S)).astype(np.int32)
    dw = {k: dev(v) for k, v in w.items()}
    pool, tab = case.device(torch)
    dL, dnew, dinp = dev(L), dev(new_idx), dev(inp)
    q = torch.zeros((B, d), device="cuda")
    qkt = torch.zeros((B, S), device="cuda")
    attn = torch.zeros((B, d), device="cuda")
    H.check_ref(ref.ref_paged_encoder(H.p(dw["emb"]), H.p(dw["pos"]), H.p(dinp), H.p(tab), H.p(dL),
                                      H.p(dnew), B, S, d, n_new))
    H.check_ref(ref.ref_paged_attention(H.p(tab), H.p(dL), H.p(dw["wk"]), H.p(dw["wq"]), H.p(dw["wv"]),
                                        H.p(dnew), H.p(q), H.p(qkt), H.p(attn), n_new, B, S, d, 0))
    pool_after = pool.cpu().numpy().copy()
    score = torch.zeros((B, V), device="cuda")
    dec = torch.zeros((B, 1), dtype=torch.int32, device="cuda")
    H.check_ref(ref.ref_paged_decoder(H.p(attn), H.p(dw["emb"]), H.p(score), H.p(dw["pos"]), H.p(tab),
                                      H.p(dL), H.p(dec), B, V, S, d, 1, 0, 0))
    np.savez_compressed(out / f"ref_cuda_{name}.npz", kind="stages", B=B, S=S, d=d, V=V, dist=dist,
                        case_seed=case_seed, w_seed=w_seed, eof_ratio=eof_ratio, lengths=L,
                        n_new=n_new, new_idx=new_idx, inp=inp, pool_after_attention=pool_after,
                        q_output=q.cpu().numpy(), softmax=qkt.cpu().numpy(),
                        attention_result=attn.cpu().numpy(), logits=score.cpu().numpy(),
                        tokens=dec.cpu().numpy(), lengths_after=dL.cpu().numpy())


def engines(out, name, case, dist, w_seed, p_seed):
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from test_gpu_forward_engine import run_ref_engine
    ref = H.load_ref()
    w = H.make_weights(w_seed, case["d"], case["V"], case["S"], dist)
    offs, toks = H.make_prompts(p_seed, case["n_req"], case["lo"], case["hi"])
    d = dict(kind="engines", dist=dist, w_seed=w_seed, p_seed=p_seed, **case)
    for label, kind in (("paged", "paged"), ("dense", "dense")):
        res, order, _ = run_ref_engine(ref, kind, case, w, offs, toks, 0)
        fo = np.zeros(len(order) + 1, np.int32)
        ft = []
        for k, rid in enumerate(order):
            ft.append(res[int(rid)])
            fo[k + 1] = fo[k] + len(res[int(rid)])
        d[f"{label}_order"] = order
        d[f"{label}_offsets"] = fo
        d[f"{label}_tokens"] = np.concatenate(ft).astype(np.int32)
    np.savez_compressed(out / f"ref_cuda_{name}.npz", **d)


if __name__ == "__main__":
    out = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
    out.mkdir(parents=True, exist_ok=True)
    stages(out, "stages_R_6x64x64", 6, 64, 64, 1024, "R", 41, 43)
    stages(out, "stages_Z_6x64x64", 6, 64, 64, 1024, "Z", 41, 43)
    stages(out, "stages_Z_5x128x132", 5, 128, 132, 1000, "Z", 45, 47)
    engines(out, "engine_Z_pressure", dict(B=8, S=128, d=32, V=1024, n_blocks=36, n_req=24, lo=20, hi=64), "Z", 31, 33)
    engines(out, "engine_R_small", dict(B=4, S=64, d=32, V=1024, n_blocks=16, n_req=10, lo=1, hi=40), "R", 31, 33)
    print("golden vectors written to", out)
