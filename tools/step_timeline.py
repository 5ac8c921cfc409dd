#!/usr/bin/env python
"""In-graph timeline of the engine step on the bench workload (BASELINE configs[1]): every kernel of
the captured step graph records %globaltimer once its dependencies are satisfied, so consecutive
stamps give the time each kernel (plus its launch gap) really takes inside the graph -- which ncu's
serialised, cold-cache per-kernel durations cannot show.  Diagnostic tool, not a benchmark.

    python tools/step_timeline.py [--pdl 0|1] [--out profiles/xyz.md]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import harness as H  # noqa: E402
import min_llm_inference_b200 as mli  # noqa: E402
from bench import WORKLOADS  # noqa: E402
WORKLOAD = WORKLOADS["c2a"]

SLOTS = ["sched_step", "encoder", "QKV+prefill GEMM", "decode attention", "logits GEMM", "decoder"]


def measure(device=0, pdl=1, workload="c2a", world=1, chunk=0):
    """one traced bench job -> (text report, dict of per-kernel in-graph times).  workload / world: the job of
    rank 0 of `bench.py --workload W` on `world` GPUs (strong workloads shrink with the world size)"""
    import bench
    wl = dict(WORKLOADS[workload])
    S, d, V = wl["S"], wl["d"], wl["V"]
    strong = wl["n_blocks"] == 0
    B = wl["B"] // world if strong else wl["B"]
    torch.cuda.set_device(device)
    ctx = mli.Context(device, torch.cuda.current_stream().cuda_stream)
    ctx.set_option(mli.OPT_PDL, pdl)
    cap = 4096
    trace = torch.zeros(8 + 8 * cap, dtype=torch.int64, device="cuda")
    trace[1] = cap
    ctx.call("mli_debug_set_step_trace", trace)
    w = H.make_weights(1001, d, V, S, "Z")
    offs, toks, _ = bench.rank_requests(wl, 0, world)
    if strong:
        plen = np.diff(offs).astype(np.int64)
        wl["n_blocks"] = int(np.maximum((plen + wl["max_new"] + 1 + 15) // 16, 4).sum()) + 64
    wl["n_req"] = len(offs) - 1
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(B, S, d, V, wl["n_blocks"], wl["R"], 0, wl["n_req"], None, wl["max_new"], 0, chunk)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    d_offs, d_toks = torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda()
    for _ in range(2):
        eng.submit(d_offs, d_toks, is_device=True)
        eng.run()
    torch.cuda.synchronize()
    trace[0] = 0
    trace[8:] = 0
    torch.cuda.synchronize()
    eng.submit(d_offs, d_toks, is_device=True)
    eng.run()
    st = eng.stats()
    torch.cuda.synchronize()
    t = trace.cpu().numpy()
    n = int(min(t[0], cap))
    full = t[8:8 + 8 * n].reshape(n, 8)
    tl = full[:, :6].astype(np.float64)
    n_act, n_gran = full[:, 6].astype(np.int64), full[:, 7].astype(np.int64)
    real = int(st.steps)
    # duration of kernel k of step i = stamp of the next kernel - its own stamp; kernels that are
    # not part of the graph (e.g. the encoder once it is folded into the GEMM) have no stamps
    present = [k for k in range(6) if tl[:real - 1, k].any()]
    dur = np.full((real - 1, 6), np.nan)
    for a, k in enumerate(present):
        nxt = tl[:real - 1, present[a + 1]] if a + 1 < len(present) else tl[1:real, 0]
        dur[:, k] = (nxt - tl[:real - 1, k]) / 1e3
    lines = [f"# In-graph step timeline, bench workload {workload} on {world} GPU(s): rank 0 (B={B}, d={d}, S={S}), pdl={pdl}",
             "",
             f"{real} engine iterations, {st.generated_tokens} tokens, device job time {st.gpu_ms:.2f} ms "
             f"({1e3 * st.gpu_ms / real:.1f} us / iteration).  Stamp = %globaltimer when the kernel's "
             "dependencies were satisfied; a kernel's time = next stamp - its stamp (includes the launch gap).",
             "",
             "| kernel | mean us | median us | p90 us | share |", "|---|---:|---:|---:|---:|"]
    tot = np.nansum(np.nanmean(dur, axis=0))
    summary = {"iterations": real, "job_ms": float(st.gpu_ms), "us_per_iteration": 1e3 * st.gpu_ms / real,
               "kernels_mean_us": {}, "timed_by": "%globaltimer stamp of every kernel once its dependencies are satisfied"}
    for k, name in enumerate(SLOTS):
        col = dur[:, k]
        col = col[~np.isnan(col)]
        if not len(col):
            lines.append(f"| {name} | - | - | - | not a separate kernel |")
            continue
        summary["kernels_mean_us"][name] = float(col.mean())
        lines.append(f"| {name} | {col.mean():.1f} | {np.median(col):.1f} | {np.percentile(col, 90):.1f} | "
                     f"{100 * col.mean() / tot:.1f}% |")
    summary["attention_total_us"] = float(np.nansum(dur[:, 3]))
    summary["attention_steps"] = int(np.sum(~np.isnan(dur[:, 3])))
    lines.append(f"\nSum of means {tot:.1f} us per iteration.")
    whole = np.nansum(dur, axis=1)
    lines.append(f"Iteration time (us): median {np.median(whole):.1f}, p99 {np.percentile(whole, 99):.1f}, max {whole.max():.1f}"
                 + (f"  [chunked prefill, {chunk} positions per step]" if chunk else ""))
    summary["iteration_us_max"] = float(whole.max())
    # the merged GEMM against the rows it saw (active rows padded to 16 + 16 per prefill granule)
    rows = ((n_act + 15) // 16 * 16 + 16 * n_gran)[:real - 1]
    g = dur[:, 2]
    lines += ["", "QKV+prefill GEMM by activation rows of the step (256 rows = one tile):", "",
              "| rows | steps | mean us | mean active rows | mean granules |", "|---|---:|---:|---:|---:|"]
    for lo, hi in ((0, 128), (129, 256), (257, 512), (513, 1 << 30)):
        m = (rows >= lo) & (rows <= hi) & ~np.isnan(g)
        if m.any():
            lines.append(f"| {lo}-{hi if hi < 1 << 30 else 'inf'} | {int(m.sum())} | {g[m].mean():.1f} | "
                         f"{n_act[:real - 1][m].mean():.0f} | {n_gran[:real - 1][m].mean():.1f} |")
    e = dur[:, 1] if 1 in present else np.full(real - 1, np.nan)
    for has in (False, True):
        m = ((n_gran[:real - 1] > 0) == has) & ~np.isnan(e)
        if m.any():
            lines.append(f"\nencoder, steps {'with' if has else 'without'} new rows: {int(m.sum())} steps, "
                         f"mean {e[m].mean():.1f} us")
    eng.close()
    ctx.close()
    return "\n".join(lines), summary


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--out", default="")
    ap.add_argument("--workload", default="c2a")
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--chunk", type=int, default=0, help="prefill_chunk_positions (0 = off)")
    args = ap.parse_args()
    text, _ = measure(0, args.pdl, args.workload, args.world, args.chunk)
    print(text)
    if args.out:
        Path(args.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
