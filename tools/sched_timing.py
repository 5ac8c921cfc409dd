#!/usr/bin/env python
"""Phase timing of the on-device scheduler inside the captured step graph (bench workload): the
scheduler's thread 0 records %globaltimer at its phase boundaries when the step trace is in extended
mode (trace[2] != 0).  Diagnostic tool, not a benchmark.      python tools/sched_timing.py
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
from bench import WORKLOADS  # noqa: E402
WORKLOAD = WORKLOADS["c2a"]

PHASES = ["own state + ring window + queue head fetched (before the wait)", "dependency wait (decoder tail)",
          "lengths loaded", "phase 1: decoder results, retire", "phase 2: free pages, compact used list",
          "phase 3: grow / pre-empt", "phase 4: admit", "mirrors + work lists", "counters, done flag"]


def main():
    wl = WORKLOAD
    B, S, d, V = wl["B"], wl["S"], wl["d"], wl["V"]
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    cap = 4096
    trace = torch.zeros(8 + 24 * cap, dtype=torch.int64, device="cuda")
    trace[1] = cap
    trace[2] = 1
    ctx.call("mli_debug_set_step_trace", trace)
    w = H.make_weights(1001, d, V, S, "Z")
    offs, toks = H.make_prompts(2002, wl["n_req"], wl["lo"], wl["hi"])
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(B, S, d, V, wl["n_blocks"], wl["R"], 0, wl["n_req"], None)
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
    n = int(st.steps)
    steps = t[8:8 + 8 * n].reshape(n, 8)
    ph = t[8 + 8 * cap:8 + 8 * cap + 16 * n].reshape(n, 16)[1:, :10].astype(np.float64)   # skip the first (plain launch)
    n_gran = steps[1:, 7]
    dur = np.diff(ph, axis=1) / 1e3
    print(f"{n} engine iterations; scheduler phases, us (mean | median | mean on the {int((n_gran > 0).sum())} admission steps)")
    for k, name in enumerate(PHASES):
        col = dur[:, k]
        print(f"  {name:68s} {col.mean():6.2f} | {np.median(col):6.2f} | {col[n_gran > 0].mean():6.2f}")
    tot = (ph[:, 9] - ph[:, 0]) / 1e3
    post = (ph[:, 9] - ph[:, 2]) / 1e3
    print(f"  {'kernel entry .. end':68s} {tot.mean():6.2f} | {np.median(tot):6.2f}")
    print(f"  {'after the wait .. end (the part on the critical path)':68s} {post.mean():6.2f} | {np.median(post):6.2f}")
    eng.close()
    ctx.close()


if __name__ == "__main__":
    main()
