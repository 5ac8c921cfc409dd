#!/usr/bin/env python
"""Engine jobs at the larger BASELINE.json configurations (configs[2] and configs[3]); these are not
bench lines (bench.py measures configs[1]) but show the same engine in the HBM-bound regime.

    python tools/run_config.py c3      # B=1024, d=2048, S=2048, prompts U[64,1024], 2048 requests to completion
    python tools/run_config.py c4      # B=128, d=4096, S=32768, prompts U[12k,20k], ~110 GB of KV pages,
                                       # prefill + 192 decode steps
Prints one JSON line per run: tokens/s, decode-step time, attention bytes/s against the measured peak.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import harness as H  # noqa: E402
import min_llm_inference_b200 as mli  # noqa: E402

PRESETS = {
    "c3": dict(B=1024, d=2048, S=2048, V=1024, n_req=2048, lo=64, hi=1024, pool_gb=60, max_steps=0),
    "c4": dict(B=128, d=4096, S=32768, V=1024, n_req=128, lo=12000, hi=20000, pool_gb=110, max_steps=192),
}


def run(name, device=0, reps=2, kv_bf16=0):
    """kv_bf16 = 1: the opt-in compact page format (MLI_OPT_KV_FORMAT = 1: K, V stored as bf16)"""
    p = PRESETS[name]
    B, d, S, V = p["B"], p["d"], p["S"], p["V"]
    torch.cuda.set_device(device)
    peak = json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (REPO / "MEASURED_PEAKS.json").exists() else 6650.0
    ctx = mli.Context(device, torch.cuda.current_stream().cuda_stream)
    if kv_bf16:
        ctx.set_option(mli.OPT_KV_FORMAT, 1)
    page_bytes = 16 * (2 if kv_bf16 else 3) * d * 4
    n_blocks = int(p["pool_gb"] * 1e9 // page_bytes)
    w = H.make_weights(1001, d, V, S, "Z")
    offs, toks = H.make_prompts(2002, p["n_req"], p["lo"], p["hi"])
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(B, S, d, V, n_blocks, 1, 0, p["n_req"], None)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    d_offs, d_toks = torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda()
    out = {"config": name, "kv_format": "compact (K, V bf16)" if kv_bf16 else "reference (fp32)", "n_batch": B, "emb_dim": d, "n_sequence": S, "kv_pool_gb": n_blocks * page_bytes / 1e9,
           "requests": p["n_req"], "prompt_tokens": int(offs[-1])}
    if p["max_steps"]:
        # step 1 = admission + prefill of every prompt; then a fixed number of decode steps
        eng.submit(d_offs, d_toks, is_device=True)
        t0 = time.perf_counter()
        eng.run(max_steps=1)
        torch.cuda.synchronize()
        out["prefill_s"] = time.perf_counter() - t0
        g0 = eng.stats().generated_tokens
        t0 = time.perf_counter()
        eng.run(max_steps=p["max_steps"])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = eng.stats()
        gen = st.generated_tokens - g0
        out.update(decode_steps=p["max_steps"], decode_tokens=int(gen), decode_s=dt,
                   tokens_per_s=gen / dt, ms_per_step=1e3 * dt / p["max_steps"])
        # rows hold ~prompt + steps tokens: attention bytes per step
        lens = np.diff(offs).astype(np.float64) + p["max_steps"] / 2
        attn_bytes = float(np.sum((4.0 if kv_bf16 else 8.0) * d * lens))
        out.update(attention_bytes_per_step=attn_bytes,
                   step_floor_ms=1e3 * attn_bytes / (peak * 1e9),
                   whole_step_GBps=attn_bytes / (dt / p["max_steps"]) / 1e9, hbm_peak_GBps=peak)
    else:
        for rep in range(reps):
            eng.submit(d_offs, d_toks, is_device=True)
            eng.run()
            st = eng.stats()
        out.update(job_ms=st.gpu_ms, steps=int(st.steps), generated_tokens=int(st.generated_tokens),
                   preemptions=int(st.preemptions), tokens_per_s=st.generated_tokens / (st.gpu_ms / 1e3),
                   us_per_step=1e3 * st.gpu_ms / max(1, st.steps))
        eng.submit(d_offs, d_toks, is_device=True)
        eng.run(profile_attention=True)
        ps = eng.stats()
        out.update(attention_GBps=ps.attn_bytes / max(ps.attn_ms, 1e-9) / 1e6, hbm_peak_GBps=peak,
                   attention_share_of_job=ps.attn_ms / max(ps.gpu_ms, 1e-9),
                   attention_ms_per_launch=ps.attn_ms / max(1, ps.attn_launches))
    eng.close()
    ctx.close()
    return out


if __name__ == "__main__":
    print(json.dumps(run(sys.argv[1] if len(sys.argv) > 1 else "c3",
                         kv_bf16=1 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else 0)), flush=True)
