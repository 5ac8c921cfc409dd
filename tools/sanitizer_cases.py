#!/usr/bin/env python
"""Small, fast invocations of every hand-rolled synchronisation protocol in the library, meant to be run
under compute-sanitizer (tools/run_sanitizer.sh): both fused attention kernels with dynamic slices forced
(cross-CTA arrival counters, partial-row merges, the producer's `issued` counter), the cluster split-K
tcgen05 GEMM (DSMEM reduce, cluster barriers), the device scheduler + six-kernel step graph with PDL,
streaming enqueue / poll, and the dense path.  Each case also checks its result, so a sanitizer-clean run
is a correct run."""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import harness as H  # noqa: E402
import min_llm_inference_b200 as mli  # noqa: E402


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def attention_case(ctx, kernel, B, S, d, min_dyn):
    rng = np.random.default_rng(B + d + kernel)
    L = rng.integers(0, S, size=B).astype(np.int32)
    L[1] = 0
    case = H.PagedCase(3, B, S, d, L, "Z")
    pool, tab = case.device(torch)
    q = torch.rand((B, d), device="cuda") - 0.5
    out = torch.zeros((B, d), device="cuda")
    ctx.set_option(mli.OPT_ATTN_KERNEL, kernel)
    ctx.set_option(mli.OPT_ATTN_MIN_DYN, min_dyn)
    for _ in range(2):   # twice: the self-resetting counters must be back at zero
        ctx.call("mli_decode_attention_paged", q, tab, dev(L), out, None, B, S, d)
    ctx.synchronize()
    ctx.set_option(mli.OPT_ATTN_KERNEL, 0)
    ctx.set_option(mli.OPT_ATTN_MIN_DYN, 4096)
    K = torch.from_numpy(case.gather(pool, 1)).cuda().double()
    V = torch.from_numpy(case.gather(pool, 2)).cuda().double()
    sc = torch.einsum("bd,bsd->bs", q.double(), K) / np.sqrt(np.float64(d))
    dL = dev(L)
    mask = torch.arange(S, device="cuda")[None, :] < dL[:, None]
    p = torch.nan_to_num(torch.softmax(sc.masked_fill(~mask, float("-inf")), dim=1), nan=0.0)
    want = torch.einsum("bs,bsd->bd", p, V)
    err = float((out.double() - want).abs().max() / want.abs().max())
    assert err < 1e-4, err
    print(f"attention kernel {kernel} B={B} S={S} d={d} min_dyn={min_dyn}: rel err {err:.1e}", flush=True)


def engine_case(ctx, mode, name):
    ctx.set_option(mli.OPT_GEMM_MODE, mode)
    cfg = dict(B=8, S=128, d=128, V=1024, n_blocks=36, n_req=14, lo=20, hi=64, max_new=6)
    w = H.make_weights(31, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(33, cfg["n_req"], cfg["lo"], cfg["hi"])
    rc, want, order, ost = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1)
    dw = {k: dev(v) for k, v in w.items()}
    ec = mli.EngineCfg(cfg["B"], cfg["S"], cfg["d"], cfg["V"], cfg["n_blocks"], 1, 0, cfg["n_req"], None, cfg["max_new"], 0)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    eng.submit(offs[:9], toks[:offs[8]])
    eng.run()
    eng.enqueue((offs[8:] - offs[8]).astype(np.int32), toks[offs[8]:])
    eng.run()
    eng.poll_finished()              # arms the device side, returns nothing
    polled, _ = eng.poll_finished()
    got, gorder = eng.results()
    eng.close()
    ties, errors = H.classify_token_mismatches(w, got, want)
    assert not errors and len(polled) == cfg["n_req"], (errors, len(polled))
    print(f"engine ({name}): {len(got)} requests, {len(ties)} tie flips", flush=True)


def gemm_case(ctx):
    """latest-token QKV at a shape whose K is split over a cluster (DSMEM reduce)"""
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    B, S, d = 40, 64, 1024
    rng = np.random.default_rng(9)
    L = rng.integers(1, S - 1, size=B).astype(np.int32)
    case = H.PagedCase(5, B, S, d, L, "Z")
    w = H.make_weights(7, d, 1024, S, "Z")
    pool, tab = case.device(torch)
    dw = {k: dev(v) for k, v in w.items()}
    q = torch.zeros((B, d), device="cuda")
    ctx.call("mli_qkv_latest_paged", tab, dev(L), dw["wk"], dw["wq"], dw["wv"], q, B, S, d)
    ctx.synchronize()
    x = np.stack([case.view(case.pool, r, int(L[r]) - 1, 0) for r in range(B)]).astype(np.float64)
    want = x @ w["wq"].astype(np.float64)
    err = H.rel_err(q.cpu().numpy(), want)
    assert err < 1e-5, err
    print(f"cluster split-K tcgen05 GEMM: rel err {err:.1e}", flush=True)
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


def main():
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "attn"):
        attention_case(ctx, 1, 24, 256, 256, 1)
        attention_case(ctx, 2, 24, 256, 256, 1)
        attention_case(ctx, 1, 12, 128, 1024, 4096)
        attention_case(ctx, 2, 12, 128, 1024, 4096)
    if which in ("all", "gemm"):
        gemm_case(ctx)
    if which in ("all", "engine"):
        engine_case(ctx, mli.GEMM_TCGEN05, "tcgen05")
        engine_case(ctx, mli.GEMM_SIMT_EXACT, "exact")
    ctx.close()
    print("sanitizer cases ok", flush=True)


if __name__ == "__main__":
    main()
