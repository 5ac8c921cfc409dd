#!/usr/bin/env python
"""Workloads for the ncu passes of round 2 (run under ncu by tools/run_ncu.sh; never a bench number):

    python tools/ncu_jobs.py attn_job [c4|c2a]   # one whole bench job, un-captured (every fused-attention launch is a
                                                 # plain launch): ncu collects dram__bytes of each; prints the job's own
                                                 # algorithmic bytes so the two can be compared launch for launch in total
    python tools/ncu_jobs.py prefill             # ONE engine step at 1024 rows that admits 256 prompts of ~1900 tokens
                                                 # (+ 768 short ones) at emb_dim 1024: a prefill-sized launch of the
                                                 # merged tcgen05 GEMM (~0.5 M positions, 2 TFLOP fp32-equivalent) on a
                                                 # 6 GB pool, small enough for --set full
    python tools/ncu_jobs.py prefill_d4096       # the same at emb_dim 4096 (16 prompts of 24-32 k tokens): K passes and the
                                                 # grouped item order; MLI_TC_KV_GROUP=32 = plain order for the A/B
    python tools/ncu_jobs.py attn_one            # ONE fused-attention launch, B=1024, d=1024, L~U[64,2176] (9.5 GB)
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import harness as H  # noqa: E402
import min_llm_inference_b200 as mli  # noqa: E402


def attn_job(key):
    import bench
    wl = bench.WORKLOADS[key]
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    w = H.make_weights(bench.SEED_W, wl["d"], wl["V"], wl["S"], "Z")
    offs, toks, n_total = bench.rank_requests(wl, 0, 1)
    if wl["n_blocks"] == 0:
        plen = np.diff(offs).astype(np.int64)
        n_blocks = int(np.maximum((plen + wl["max_new"] + 1 + 15) // 16, 4).sum()) + 64
    else:
        n_blocks = wl["n_blocks"]
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(wl["B"], wl["S"], wl["d"], wl["V"], n_blocks, wl["R"], 0, n_total, None, wl["max_new"], 0)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    eng.submit(torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda(), is_device=True)
    eng.run(profile_attention=True)
    ps = eng.stats()
    print(json.dumps({"workload": key, "attn_launches": int(ps.attn_launches), "algorithmic_bytes_total": ps.attn_bytes,
                      "algorithmic_bytes_per_launch": ps.attn_bytes / max(1, ps.attn_launches),
                      "engine_iterations": int(ps.steps)}), flush=True)
    eng.close()
    ctx.close()


def prefill(stamps=False):
    """one engine step at B = 1024 rows (the launch plan of the large configurations: no split-K, 32 tile
    walkers per feature tile) that admits 256 prompts of ~1900 tokens and 768 short ones: ~0.5 M positions"""
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    B, S, d, V, n_req = 1024, 2304, 1024, 1024, 1024
    w = H.make_weights(1001, d, V, S, "Z")
    rng = np.random.default_rng(2002)
    lens = np.where(np.arange(n_req) % 4 == 0, rng.integers(1700, 2100, size=n_req), rng.integers(4, 12, size=n_req))
    offs = np.zeros(n_req + 1, np.int32)
    offs[1:] = np.cumsum(lens)
    toks = rng.integers(0, 1023, size=int(offs[-1])).astype(np.int32)
    n_blocks = int(np.maximum((lens + 2 + 15) // 16, 4).sum()) + 64
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(B, S, d, V, n_blocks, 1, 0, n_req, None, 4, 0)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    ms = []
    if stamps:   # before the step graph is captured (the pointer is baked into the kernel arguments)
        buf = torch.zeros((4096, 8), dtype=torch.int64, device="cuda")
        ctx.call("mli_debug_set_gemm_stamps", buf)
    for rep in range(2):
        eng.submit(torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda(), is_device=True)
        eng.run(max_steps=1)
        ms.append(eng.stats().gpu_ms)
    pos = int(offs[-1])
    if stamps:
        # clock64 phase stamps of the LAST item every CTA processed (slots: 0 start, 1 set-up done, 2 first MMA of the
        # item issued, 3 converters done, 4 accumulators complete, 5 after the exchange barrier, 6 rows stored, 7 end)
        torch.cuda.synchronize()
        buf.zero_()
        eng.submit(torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda(), is_device=True)
        eng.run(max_steps=1)
        torch.cuda.synchronize()
        st = buf.cpu().numpy()
        nz = st[st[:, 0] != 0]
        print(f"{len(nz)} stamped CTAs; kernel cycles (t7 - t0): {np.sort((nz[:, 7] - nz[:, 0]))[::max(1, len(nz) // 12)].tolist()}")
        print("sample rows (deltas):", np.diff(nz[-4:], axis=1).tolist())
        mm = buf.cpu().numpy()[2048:2048 + 160]
        mm = mm[mm[:, 3] != 0]
        if len(mm):
            print(f"MMA issuer of {len(mm)} pairs, last item: wait for weights {mm[:, 0].mean():.0f}, wait for activations "
                  f"{mm[:, 1].mean():.0f}, whole issue loop {mm[:, 2].mean():.0f} cycles over {mm[0, 3]} k-blocks; items per pair "
                  f"{mm[:, 4].mean():.1f}")
        cv = buf.cpu().numpy()[2304:2304 + 160]
        for nm, sel in (("leader CTAs", cv[0::2]), ("peer CTAs", cv[1::2])):
            sel = sel[sel[:, 1] != 0]
            if len(sel):
                print(f"converter warp 4 of {len(sel)} {nm}, last item (its 16 k-blocks): waiting for a free stage {sel[:, 0].mean():.0f}, "
                      f"convert + store (incl. waiting for the loads) {sel[:, 1].mean():.0f}, proxy fence {sel[:, 2].mean():.0f}, "
                      f"syncwarp + arrive {sel[:, 3].mean():.0f} cycles")
        st = nz[(nz[:, 7] - nz[:, 0]) > 1000000]
        if not len(st):
            return
        names = ["main loop of the last item (first MMA -> converters done)", "MMA tail (converters done -> accumulators complete)",
                 "TMEM -> smem + barrier", "rows to global"]
        print(f"{len(st)} CTAs with a long run; kernel {np.mean(st[:, 7] - st[:, 0]):.0f} cycles")
        for i, nm in zip((2, 3, 4, 5), names):
            col = (st[:, i + 1] - st[:, i]).astype(np.float64)
            print(f"  {nm:60s} mean {col.mean():8.0f}  min {col.min():8.0f}  max {col.max():8.0f} cycles")
    print(json.dumps({"prefill_positions": pos, "fp32_equivalent_flops": 4.0 * (pos - n_req) * d * d + 6.0 * B * d * d,
                      "rows": B, "emb_dim": d, "kv_pool_gb": n_blocks * 16 * 3 * d * 4 / 1e9, "step_ms": ms}), flush=True)
    eng.close()
    ctx.close()


def attn_one():
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    B, S, d = 1024, 2304, 1024
    rng = np.random.default_rng(7)
    L = rng.integers(64, 2176, size=B).astype(np.int32)
    W = S // 16
    need = (L + 15) // 16
    n_pages = int(need.sum())
    pool = torch.empty((n_pages, 16 * 3 * d), device="cuda")
    for i in range(0, n_pages, 4096):
        pool[i:i + 4096].uniform_(-1.0, 1.0)
    perm = rng.permutation(n_pages)
    tab = np.zeros((B, W), np.uint64)
    k = 0
    for r in range(B):
        ids = perm[k:k + need[r]]
        tab[r, :need[r]] = np.uint64(pool.data_ptr()) + ids.astype(np.uint64) * np.uint64(16 * 3 * d * 4)
        k += need[r]
    dtab = torch.from_numpy(tab.view(np.int64)).cuda()
    dL = torch.from_numpy(L).cuda()
    q = (torch.rand((B, d), device="cuda") - 0.5) * 0.1
    out = torch.empty((B, d), device="cuda")
    for _ in range(3):
        ctx.call("mli_decode_attention_paged", q, dtab, dL, out, None, B, S, d)
    ctx.synchronize()
    nbytes = float(np.sum(8.0 * d * L + 8.0 * d + 8.0 * need + 4.0))
    print(json.dumps({"B": B, "emb_dim": d, "algorithmic_bytes_per_launch": nbytes}), flush=True)
    ctx.close()


def prefill_d4096():
    """one engine step at emb_dim 4096 that admits 16 prompts of 24-32 k tokens (configs[3] dimensions, ~0.45 M
    positions, 23 GB of pages): the merged GEMM on CTA pairs with two K passes.  MLI_TC_KV_GROUP=32 in the environment
    selects the plain item order (all 32 feature pairs of an activation tile, then the next tile) for the A/B of the
    DRAM traffic"""
    torch.cuda.set_device(0)
    ctx = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    B, S, d, V, n_req = 16, 32768, 4096, 1024, 16
    w = H.make_weights(1001, d, V, S, "Z")
    offs, toks = H.make_prompts(2002, n_req, 24000, 32511)
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(B, S, d, V, B * (S // 16), 1, 0, n_req, None, 2, 0)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    ms = []
    for rep in range(2):
        eng.submit(torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda(), is_device=True)
        eng.run(max_steps=1)
        ms.append(eng.stats().gpu_ms)
    pos = int(offs[-1])
    print(json.dumps({"emb_dim": d, "prompt_positions": pos, "step_ms": ms,
                      "fp32_equivalent_tflop": 4.0 * pos * d * d / 1e12,
                      "algorithmic_bytes": {"activations_read": 4.0 * pos * d, "kv_written": 8.0 * pos * d,
                                            "weights_hi_lo": 2 * 3 * 4.0 * d * d}}), flush=True)
    eng.close()
    ctx.close()


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "attn_job":
        attn_job(sys.argv[2] if len(sys.argv) > 2 else "c4")
    elif what == "prefill":
        prefill()
    elif what == "prefill_stamps":
        prefill(stamps=True)
    elif what == "prefill_d4096":
        prefill_d4096()
    else:
        attn_one()
