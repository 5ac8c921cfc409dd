#!/usr/bin/env python
"""Phase timing of the tcgen05 GEMM (clock64 stamps written by the kernel itself) plus CUDA-event
time per launch, for the three contractions of the decode step.  Diagnostic tool, not a benchmark.

    python tools/gemm_timing.py [B d V S]
"""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import harness as H  # noqa: E402
import min_llm_inference_b200 as mli  # noqa: E402

NAMES = ["setup", "to first MMA", "mainloop (conv done)", "MMA tail", "epilogue+barrier", "reduce", "tail barrier"]


def main():
    B, d, V, S = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (256, 1024, 1024, 128)
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    rng = np.random.default_rng(0)
    L = rng.integers(1, S - 1, size=B).astype(np.int32)
    case = H.PagedCase(1, B, S, d, L, "Z")
    w = H.make_weights(5, d, V, S, "Z")
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    pool, tab = case.device(torch)
    dL = torch.from_numpy(L).cuda()
    q = torch.zeros((B, d), device="cuda")
    attn = torch.rand((B, d), device="cuda") - 0.5
    score = torch.zeros((B, V), device="cuda")
    dec = torch.zeros((B, 1), dtype=torch.int32, device="cuda")
    new_idx = torch.arange(B, dtype=torch.int32, device="cuda")
    ctx.register_weights(dw["wk"], dw["wq"], dw["wv"], dw["emb"], d, V)
    stamps = torch.zeros((4096, 8), dtype=torch.int64, device="cuda")

    def latest():
        ctx.call("mli_qkv_latest_paged", tab, dL, dw["wk"], dw["wq"], dw["wv"], q, B, S, d)

    def logits():
        L2 = dL.clone()
        ctx.call("mli_paged_decoder", attn, dw["emb"], score, dw["pos"], tab, L2, dec, B, V, S, d, 1, 0)

    def prefill(n_new):
        ctx.call("mli_prefill_kv_paged", tab, new_idx, dL, dw["wk"], dw["wv"], n_new, B, S, d)

    for name, fn in (("latest QKV", latest), ("logits+decoder", logits), ("prefill 2 rows", lambda: prefill(2)),
                     ("prefill all rows", lambda: prefill(B))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        stamps.zero_()
        ctx.call("mli_debug_set_gemm_stamps", stamps)
        fn()
        torch.cuda.synchronize()
        ctx.call("mli_debug_set_gemm_stamps", None)
        st = stamps.cpu().numpy()
        st = st[st[:, 0] != 0]
        dl = np.diff(st, axis=1).astype(np.float64)
        print(f"{name}: {us:.1f} us/call (events, back to back, incl. any helper kernels); {len(st)} CTAs stamped")
        for i, nm in enumerate(NAMES):
            col = dl[:, i]
            print(f"    {nm:24s} mean {col.mean():9.0f} cyc   min {col.min():9.0f}   max {col.max():9.0f}")
        tot = (st[:, 7] - st[:, 0]).astype(np.float64)
        print(f"    {'total in-kernel':24s} mean {tot.mean():9.0f} cyc   max {tot.max():9.0f}")
    ctx.close()


if __name__ == "__main__":
    main()
