#!/usr/bin/env python
"""Engine jobs at the BASELINE.json configurations other than the bench headline, at their NAMED shapes
(bench.py reports them as extras; SURVEY 8d gives the concrete numbers):

    python tools/run_config.py c2a     # configs[1]: B=256, d=1024, S=128, 1024 pages, 512 requests U[1,64], to completion
    python tools/run_config.py c3      # configs[2]: B=1024, d=2048, S=4096, 2048 requests U[64,2048], 256-token cap,
                                       #             40 GB of KV pages (continuous batching under pool pressure)
    python tools/run_config.py c4      # configs[3]: B=128, d=4096, S=32768, prompts U[24000,32511], KV pages fill the
                                       #             HBM that is free (~165 GB), prefill + 192 decode steps
    python tools/run_config.py c3 bf16 # the same with the opt-in compact page format
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
    "c2a": dict(B=256, d=1024, S=128, V=1024, n_req=512, lo=1, hi=64, n_blocks=1024, max_new=0, max_steps=0),
    # the same with the opt-in admission throttle: at most 256 prompt positions admitted per step
    "c2a_pf": dict(B=256, d=1024, S=128, V=1024, n_req=512, lo=1, hi=64, n_blocks=1024, max_new=0, max_steps=0,
                   max_prefill=256),
    # chunked prefill: at most 256 prompt positions prefilled per step, long prompts cut into chunks
    "c2a_chunk": dict(B=256, d=1024, S=128, V=1024, n_req=512, lo=1, hi=64, n_blocks=1024, max_new=0, max_steps=0,
                      chunk=256),
    # n_forward_rounds = 4: four decode rounds per scheduler iteration
    "c2a_r4": dict(B=256, d=1024, S=128, V=1024, n_req=512, lo=1, hi=64, n_blocks=1024, max_new=0, max_steps=0, R=4),
    "c3": dict(B=1024, d=2048, S=4096, V=1024, n_req=2048, lo=64, hi=2048, pool_gb=40, max_new=256, max_steps=0),
    "c4": dict(B=128, d=4096, S=32768, V=1024, n_req=128, lo=24000, hi=32768 - 257, pool_gb=-8, max_new=0,
               max_steps=192),
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
    w = H.make_weights(1001, d, V, S, "Z")
    offs, toks = H.make_prompts(2002, p["n_req"], p["lo"], p["hi"])
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    d_offs, d_toks = torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda()
    if "n_blocks" in p:
        n_blocks = p["n_blocks"]
    elif p["pool_gb"] > 0:
        n_blocks = int(p["pool_gb"] * 1e9 // page_bytes)
    else:
        # "KV pages sized to fill 180 GB HBM": everything that is free now, less a margin for the engine's
        # own tables, the split weight copies and the attention workspaces
        torch.cuda.empty_cache()
        free, total = torch.cuda.mem_get_info()
        n_blocks = int((free + p["pool_gb"] * 1e9) // page_bytes)
    ec = mli.EngineCfg(B, S, d, V, n_blocks, p.get("R", 1), 0, p["n_req"], None, p["max_new"], p.get("max_prefill", 0), p.get("chunk", 0))
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    out = {"config": name, "kv_format": "compact (K, V bf16)" if kv_bf16 else "reference (fp32)", "n_batch": B,
           "emb_dim": d, "n_sequence": S, "kv_pool_gb": n_blocks * page_bytes / 1e9, "kv_pages": n_blocks,
           "requests": p["n_req"], "prompt_lengths": f"U[{p['lo']},{p['hi']}]", "prompt_tokens": int(offs[-1]),
           "max_new_tokens": p["max_new"], "max_prefill_positions": p.get("max_prefill", 0),
           "n_forward_rounds": p.get("R", 1), "prefill_chunk_positions": p.get("chunk", 0)}
    if p["max_steps"]:
        # step 1 = admission + prefill of every prompt that fits; then a fixed number of decode steps
        eng.submit(d_offs, d_toks, is_device=True)
        t0 = time.perf_counter()
        eng.run(max_steps=1)
        torch.cuda.synchronize()
        out["prefill_s"] = time.perf_counter() - t0
        st0 = eng.stats()
        g0 = st0.generated_tokens
        t0 = time.perf_counter()
        eng.run(max_steps=p["max_steps"])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = eng.stats()
        gen = st.generated_tokens - g0
        resident = st.peak_resident_rows
        out.update(decode_steps=p["max_steps"], decode_tokens=int(gen), decode_s=dt, resident_rows=int(resident),
                   peak_pool_occupancy=1.0 - st.min_free_pages / n_blocks, preemptions=int(st.preemptions),
                   tokens_per_s=gen / dt, ms_per_step=1e3 * dt / p["max_steps"])
        # the resident rows hold ~prompt + steps/2 tokens: attention bytes per step.  Requests are
        # admitted in queue order, so the resident ones are the first `resident`
        lens = np.diff(offs).astype(np.float64)[:resident] + p["max_steps"] / 2
        attn_bytes = float(np.sum((4.0 if kv_bf16 else 8.0) * d * lens))
        out.update(attention_bytes_per_step=attn_bytes,
                   step_floor_ms=1e3 * attn_bytes / (peak * 1e9),
                   whole_step_GBps=attn_bytes / (dt / p["max_steps"]) / 1e9, hbm_peak_GBps=peak,
                   prefill_positions=int(np.diff(offs)[:resident].sum()),
                   prefill_tflops_fp32_equiv=4.0 * float(np.diff(offs)[:resident].sum()) * d * d / out["prefill_s"] / 1e12)
    else:
        for rep in range(reps):
            eng.submit(d_offs, d_toks, is_device=True)
            eng.run()
            st = eng.stats()
        out.update(job_ms=st.gpu_ms, steps=int(st.steps), generated_tokens=int(st.generated_tokens),
                   preemptions=int(st.preemptions), peak_resident_rows=int(st.peak_resident_rows),
                   peak_pool_occupancy=1.0 - st.min_free_pages / n_blocks,
                   tokens_per_s=st.generated_tokens / (st.gpu_ms / 1e3),
                   us_per_step=1e3 * st.gpu_ms / max(1, st.steps))
        eng.submit(d_offs, d_toks, is_device=True)
        eng.run(profile_attention=True)
        ps = eng.stats()
        out.update(attention_GBps=ps.attn_bytes / max(ps.attn_ms, 1e-9) / 1e6, hbm_peak_GBps=peak,
                   attention_share_of_job=ps.attn_ms / max(ps.gpu_ms, 1e-9),
                   attention_ms_per_launch=ps.attn_ms / max(1, ps.attn_launches),
                   gemm_share_of_job=ps.gemm_ms / max(ps.gpu_ms, 1e-9))
        if ps.gemm_max_ms > 0:
            out.update(prefill_launch_ms=ps.gemm_max_ms,
                       prefill_launch_tflops_fp32_equiv=ps.gemm_max_flops / ps.gemm_max_ms / 1e9)
    eng.close()
    ctx.close()
    return out


if __name__ == "__main__":
    print(json.dumps(run(sys.argv[1] if len(sys.argv) > 1 else "c3",
                         kv_bf16=1 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else 0)), flush=True)
