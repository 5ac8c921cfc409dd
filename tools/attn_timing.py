#!/usr/bin/env python
"""Phase timing of the fused decode-attention kernel (clock64 stamps written by the kernel) at the
bench workload's in-job shape: B=256, d=1024, S=128, lengths like a mid-job engine step.
Diagnostic tool, not a benchmark.      python tools/attn_timing.py [B d S meanL]
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import harness as H  # noqa: E402
import min_llm_inference_b200 as mli  # noqa: E402

NAMES = ["setup+wait", "length scan", "zero-fill", "first data", "main loop(last seg)", "epilogue"]


def main():
    B, d, S, frac = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (256, 1024, 128, 58)
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    if os.environ.get("ATTN_KV_BF16"):
        ctx.set_option(mli.OPT_KV_FORMAT, 1)   # timing only: the pool below keeps the fp32 page size
    if os.environ.get("ATTN_KERNEL"):
        ctx.set_option(mli.OPT_ATTN_KERNEL, int(os.environ["ATTN_KERNEL"]))   # 1 column-split, 2 warp-per-position
    if os.environ.get("ATTN_CTAS"):
        ctx.set_option(mli.OPT_ATTN_CTAS_PER_SM, int(os.environ["ATTN_CTAS"]))
    rng = np.random.default_rng(int(os.environ.get("ATTN_SEED", "0")))
    L = rng.integers(1, S - 1, size=B).astype(np.int32)
    L[rng.random(B) > frac / 100.0] = 0          # ~58 % of rows active, as in the bench job
    case = H.PagedCase(1, B, S, d, L, "Z")
    pool, tab = case.device(torch)
    if os.environ.get("ATTN_ALIAS_MB"):
        # experiment: fold the page table onto the first N MB of the pool (same logical work, few
        # distinct 2 MB translations) -- separates address-translation effects from the rest
        page_bytes = 16 * 3 * d * 4
        n_alias = max(1, int(float(os.environ["ATTN_ALIAS_MB"]) * 2**20) // page_bytes)
        base = pool.data_ptr()
        live = tab != 0
        ids = (tab - base) // page_bytes
        tab = torch.where(live, base + (ids % n_alias) * page_bytes, tab)
        print(f"page table folded onto {n_alias} pages ({n_alias * page_bytes / 2**20:.1f} MB)")
    dL = torch.from_numpy(L).cuda()
    q = (torch.rand((B, d), device="cuda") - 0.5) * 0.1
    out = torch.empty((B, d), device="cuda")
    nbytes = float(np.sum((8.0 * d * L + 8.0 * d + 8.0 * ((L + 15) // 16) + 4.0) * (L > 0)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stamps = torch.zeros((1024, 16), dtype=torch.int64, device="cuda")

    def run():
        ctx.call("mli_decode_attention_paged", q, tab, dL, out, None, B, S, d)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    for cold in (False, True):
        ts = []
        for _ in range(10):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts))
        print(f"{'cold L2' if cold else 'warm L2'}: {us:.1f} us/launch (events around one launch), "
              f"{nbytes / 1e6:.1f} MB algorithmic -> {nbytes / us / 1e3:.0f} GB/s")
    # back-to-back launches (launch overhead amortised)
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"back to back (warm L2): {us:.1f} us/launch -> {nbytes / us / 1e3:.0f} GB/s")
    flush.zero_()
    ctx.call("mli_debug_set_gemm_stamps", stamps)
    run()
    torch.cuda.synchronize()
    ctx.call("mli_debug_set_gemm_stamps", None)
    st = stamps.cpu().numpy()
    st = st[st[:, 0] != 0]
    smid = st[:, 7]
    gt = st[:, 8:11].astype(np.float64)
    merged = st[:, 11]
    sm_of_cta = st[:, 12]
    st = st[:, :7]
    # wall-clock picture (globaltimer, common to all SMs): when CTAs start, stop streaming, and end
    t0 = gt[:, 0].min()
    rel = (gt - t0) / 1e3
    print("globaltimer us after the first CTA start, pct 0/10/50/90/100:")
    for i, nm in enumerate(["CTA start", "last segment done", "CTA end"]):
        print(f"    {nm:18s}", np.percentile(rel[:, i], [0, 10, 50, 90, 100]).round(1))
    order = np.argsort(-rel[:, 2])[:10]
    print("last CTAs to end (cta, start, last-seg, end us; stages, segments; rows merged, segments merged):")
    for i in order:
        print(f"    {i:4d} {rel[i, 0]:8.1f} {rel[i, 1]:8.1f} {rel[i, 2]:8.1f}   {int(smid[i] & 0xffffffff):6d} {int(smid[i] >> 32):4d}"
              f"   {int(merged[i] & 0xffffffff):4d} {int(merged[i] >> 32):5d}")
    have = st[:, 4] != 0
    t_first = (st[have, 4] - st[have, 3]).astype(np.float64)
    t_main = (st[have, 5] - st[have, 4]).astype(np.float64)
    t_tot = (st[:, 6] - st[:, 0]).astype(np.float64)
    print("first-data pct 10/50/90/100:", np.percentile(t_first, [10, 50, 90, 100]).round())
    print("main loop  pct 10/50/90/100:", np.percentile(t_main, [10, 50, 90, 100]).round())
    print("total      pct 10/50/90/100:", np.percentile(t_tot, [10, 50, 90, 100]).round())
    start0 = st[:, 0].min()
    print("kernel span (first CTA start .. last CTA end, cycles; SM clocks are not synchronised):",
          int(st[:, 6].max() - start0))
    nseg, nstage = (smid >> 32)[have], (smid & 0xffffffff)[have]
    # main-loop time ~ a * stages + b * segments + c  (least squares over the CTAs)
    A = np.stack([nstage, nseg, np.ones_like(nseg)], axis=1).astype(np.float64)
    coef, *_ = np.linalg.lstsq(A, t_main, rcond=None)
    print(f"main loop ~ {coef[0]:.0f} cyc/stage + {coef[1]:.0f} cyc/segment + {coef[2]:.0f}   "
          f"(stages per CTA {nstage.min()}..{nstage.max()}, segments {nseg.min()}..{nseg.max()})")
    order = np.argsort(-t_main)[:8]
    print("slowest main loops (stages, segments, cycles):", [(int(nstage[i]), int(nseg[i]), int(t_main[i])) for i in order])
    order = np.argsort(t_main)[:8]
    print("fastest main loops (stages, segments, cycles):", [(int(nstage[i]), int(nseg[i]), int(t_main[i])) for i in order])
    # which SMs are slow?  (ATTN_SM_DUMP=file: per-SM mean main-loop cycles, to compare runs / seeds)
    if os.environ.get("ATTN_SM_DUMP"):
        sm = sm_of_cta[have]
        per_sm = np.array([t_main[sm == i].mean() if np.any(sm == i) else np.nan for i in range(int(sm.max()) + 1)])
        np.save(os.environ["ATTN_SM_DUMP"], per_sm)
        print("per-SM mean main loop: min/median/max", np.nanmin(per_sm).round(), np.nanmedian(per_sm).round(),
              np.nanmax(per_sm).round(), " slowest SMs:", np.argsort(-np.nan_to_num(per_sm))[:12].tolist())
    dl = np.diff(st, axis=1).astype(np.float64)
    print(f"{len(st)} CTAs stamped (cold L2)")
    for i, nm in enumerate(NAMES):
        col = dl[:, i]
        print(f"    {nm:22s} mean {col.mean():8.0f} cyc   min {col.min():8.0f}   max {col.max():8.0f}")
    tot = (st[:, 6] - st[:, 0]).astype(np.float64)
    print(f"    {'total in-kernel':22s} mean {tot.mean():8.0f} cyc   max {tot.max():8.0f}")
    ctx.close()


if __name__ == "__main__":
    main()
