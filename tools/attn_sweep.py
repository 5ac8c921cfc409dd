#!/usr/bin/env python
"""Fused decode attention alone over a sweep of shapes and both KV page formats: us per launch and
achieved bytes/s against the measured HBM peak.  KV pages are generated on the device (no host pool),
so the configs[2] / configs[3] shapes run in seconds.  Diagnostic tool, not a benchmark line.

    python tools/attn_sweep.py [--json out.jsonl]
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import min_llm_inference_b200 as mli  # noqa: E402

# name, B, d, S, (lo, hi) context lengths, fraction of rows active
SHAPES = [
    ("bench step (configs[1])", 256, 1024, 128, (1, 126), 0.58),
    ("configs[2]", 1024, 2048, 2048, (64, 2048), 1.0),
    ("d=512 long", 2048, 512, 2048, (64, 2048), 1.0),
    ("d=1024 long", 1024, 1024, 2048, (64, 2048), 1.0),
    ("d=2048 mid", 512, 2048, 1024, (1, 1023), 1.0),
    ("d=4096 short", 128, 4096, 2048, (1, 2047), 1.0),
    ("configs[3]", 128, 4096, 32768, (12000, 20000), 1.0),
    ("B=2048 d=1024", 2048, 1024, 2304, (64, 2176), 1.0),
    ("B=4096 d=1024", 4096, 1024, 2304, (64, 2176), 1.0),
    ("configs[4] step (N=1)", 8192, 1024, 2304, (64, 2176), 1.0),
]


def one(ctx, name, B, d, S, lohi, frac, kv_bf16):
    rng = np.random.default_rng(7)
    L = rng.integers(lohi[0], lohi[1], size=B).astype(np.int32)
    L[rng.random(B) > frac] = 0
    W = S // 16
    need = (L + 15) // 16
    n_pages = int(need.sum())
    page_words = 16 * (2 if kv_bf16 else 3) * d
    pool = torch.empty((n_pages, page_words), device="cuda", dtype=torch.float32)
    for i in range(0, n_pages, 4096):
        if kv_bf16:
            pool[i:i + 4096].view(torch.bfloat16).uniform_(-1.0, 1.0)
        else:
            pool[i:i + 4096].uniform_(-1.0, 1.0)
    perm = rng.permutation(n_pages)
    tab = np.zeros((B, W), np.uint64)
    k = 0
    for r in range(B):
        ids = perm[k:k + need[r]]
        tab[r, :need[r]] = np.uint64(pool.data_ptr()) + ids.astype(np.uint64) * np.uint64(page_words * 4)
        k += need[r]
    dtab = torch.from_numpy(tab.view(np.int64)).cuda()
    dL = torch.from_numpy(L).cuda()
    q = (torch.rand((B, d), device="cuda") - 0.5) * 0.1
    out = torch.empty((B, d), device="cuda")

    def run():
        ctx.call("mli_decode_attention_paged", q, dtab, dL, out, None, B, S, d)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    kv_row = (4.0 if kv_bf16 else 8.0) * d
    nbytes = float(np.sum((kv_row * L + 8.0 * d + 8.0 * need + 4.0) * (L > 0)))
    n = 20 if nbytes < 2e9 else 5
    best = None
    for _ in range(3):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            run()
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) * 1e3 / n
        best = us if best is None else min(best, us)
    del pool
    return {"shape": name, "B": B, "emb_dim": d, "S": S, "kv_format": "bf16" if kv_bf16 else "fp32",
            "MB": nbytes / 1e6, "us_per_launch": best, "GBps": nbytes / best / 1e3}


def main():
    torch.cuda.set_device(0)
    peak_f = REPO / "MEASURED_PEAKS.json"
    peak = json.loads(peak_f.read_text())["hbm_gbs"] if peak_f.exists() else 6650.0
    rows = []
    for kv_bf16 in (0, 1):
        ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
        if kv_bf16:
            ctx.set_option(mli.OPT_KV_FORMAT, 1)
        if os.environ.get("ATTN_KERNEL"):
            ctx.set_option(mli.OPT_ATTN_KERNEL, int(os.environ["ATTN_KERNEL"]))   # 1 column-split, 2 warp-per-position
        only = os.environ.get("ATTN_ONLY", "")
        for sh in SHAPES:
            if only and only not in sh[0]:
                continue
            r = one(ctx, *sh, kv_bf16)
            r["frac_of_measured_peak"] = r["GBps"] / peak
            rows.append(r)
            print(f"{r['shape']:26s} {r['kv_format']:5s} {r['MB']:10.1f} MB {r['us_per_launch']:10.1f} us "
                  f"{r['GBps']:8.0f} GB/s  {r['frac_of_measured_peak']:.2f} of peak", flush=True)
        ctx.close()
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
