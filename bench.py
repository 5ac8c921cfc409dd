#!/usr/bin/env python
"""bench.py -- decode tokens/s of the paged-attention decode path (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- paged attention, page_size 16, n_batch 256,
emb_dim 1024, n_sequence 128, vocab 1024, pool 1024 pages, 512 requests with prompt lengths
U[1,64], n_forward_rounds 1 (the reference's tests/paged_for_profile.cpp workload at the
BASELINE-named shape; SURVEY 8d "C2a").  Synthetic, fixed seeds, zero-mean weights (dist "Z": the
reference's U(0,1] fixtures make softmax one-hot and every row emit the same token; SURVEY
finding 9), corrected lengths (the reference's quirk Q1 is replayed only in the parity tests).

A "step" = one whole engine job: all requests admitted, prefilled, decoded to EOS / n_sequence and
retired by the on-device scheduler.  `value` = generated tokens / device time with the prompts
already resident in HBM; `e2e` = the same job through the host-buffer C ABI (prompts H2D from
pinned memory, finished token lists D2H) by wall clock.  With N > 1 (torchrun) every rank runs the
same-sized job on its own requests (weak scaling, request sharding; no collective on the data
path) and the final tokens are all-gathered over NCCL inside the timed region.

`--impl reference` times the reference's own HOST implementation of the path (oracle/_ref:
tests/test_utils.cpp host loops driven by the reference's scheduler; single-threaded as written) on
a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

WORKLOAD = dict(name="BASELINE.json configs[1]: paged attention, n_batch=256, emb_dim=1024, "
                     "n_sequence=128, page_size=16, pool=1024 pages, 512 requests, prompts U[1,64]",
                B=256, S=128, d=1024, V=1024, n_blocks=1024, n_req=512, lo=1, hi=64, R=1)
METRIC = "decode tokens/sec (paged attention, continuous batching)"
UNIT = "tokens/s"


def attention_traffic():
    """DRAM bytes per launch of the fused attention from the committed `ncu --set full` capture
    (profiles/r1_attn_traffic.json), with the algorithmic bytes of the captured launch beside it"""
    f = REPO / "profiles" / "r1_attn_traffic.json"
    if not f.exists():
        return None, None
    t = json.loads(f.read_text())
    return t["dram_bytes_per_launch"], t


def peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """dense bf16 TFLOP/s (sustained: the GEMM is timed inside a long job); tf32 is half of it"""
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        m = json.loads(f.read_text())
        return float(m.get("bf16_tflops_sustained", m.get("bf16_tflops", 1400.0))), "measured (MEASURED_PEAKS.json)"
    return 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9])
                          if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class StdoutToStderr:
    """the reference's ThroughputCounter printf()s to fd 1; keep stdout to the one JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        import ctypes
        try:
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ---------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """the reference's own CPU implementation of the path (oracle/_ref), bounded sample"""
    if rank != 0:
        return
    import harness as H
    wl = WORKLOAD
    sample_rows, sample_iters = 16, 8
    w = H.make_weights(1001, wl["d"], wl["V"], wl["S"], "Z")
    offs, toks = H.make_prompts(2002, wl["n_req"], wl["lo"], wl["hi"])
    offs_s = offs[:sample_rows + 1].copy()
    toks_s = toks[:offs_s[-1]].copy()
    sample = (f"first {sample_rows} of {wl['n_req']} requests on {sample_rows} rows, prefill + "
              f"{sample_iters} engine iterations per step, same d/S/V/prompt distribution")
    times, gens = [], []
    import torch  # the reference's host tensors are cudaHostAlloc'd: it needs a CUDA context
    if H.ref_available() and torch.cuda.is_available():
        torch.cuda.init()
        ref = H.load_ref()
        kind, cores = "reference", 1
        for i in range(args.warmup + args.steps):
            gen, steps, sec = C.c_longlong(0), C.c_longlong(0), C.c_double(0)
            with StdoutToStderr():
                H.check_ref(ref.ref_run_host_engine(sample_rows, wl["S"], wl["d"], wl["V"], H.p(w["emb"]),
                                                    H.p(w["pos"]), H.p(w["wk"]), H.p(w["wq"]), H.p(w["wv"]),
                                                    sample_rows, H.p(offs_s), H.p(toks_s), sample_iters,
                                                    C.byref(gen), C.byref(steps), C.byref(sec), None, None,
                                                    None, None))
            if i >= args.warmup:
                times.append(sec.value)
                gens.append(gen.value)
    else:
        kind, cores = "port", os.cpu_count() or 1
        cfg = dict(B=sample_rows, S=wl["S"], d=wl["d"], V=wl["V"], n_blocks=sample_rows * 8, R=1)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            rc, _, _, st = H.run_oracle_engine("paged", cfg, w, offs_s, toks_s, fix=1,
                                               max_steps=sample_iters, threads=cores)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
                gens.append(st.generated_tokens)
    total_t, total_g = float(sum(times)), float(sum(gens))
    value = total_g / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD["name"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def cpu_baseline_leg():
    """oracle port (OpenMP over rows) on all host cores, bounded sample of the same workload"""
    import harness as H
    wl = WORKLOAD
    cores = os.cpu_count() or 1
    iters = 24
    w = H.make_weights(1001, wl["d"], wl["V"], wl["S"], "Z")
    offs, toks = H.make_prompts(2002, wl["n_req"], wl["lo"], wl["hi"])
    cfg = dict(B=wl["B"], S=wl["S"], d=wl["d"], V=wl["V"], n_blocks=wl["n_blocks"], R=1)
    t0 = time.perf_counter()
    rc, _, _, st = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1, max_steps=iters, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": st.generated_tokens / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"full workload shape, first {iters} engine iterations incl. the initial prefill of "
                      f"{wl['B']} rows ({st.generated_tokens} tokens in {dt:.1f} s), OpenMP over rows"}


def reference_cuda_leg(w, offs, toks):
    """the reference's CUDA engines (rebuilt for sm_100) on the same tensors, their own metric:
    generated tokens / wall time including host scheduling (src/throughput_counter.cpp)"""
    import harness as H
    if not H.ref_available():
        return None
    ref = H.load_ref()
    wl = WORKLOAD
    n_req, S = wl["n_req"], wl["S"]
    out = {}
    for name, variant in (("warp_tiling_cublas", 1), ("naive", 0)):
        ids = np.zeros(n_req, np.int32)
        fo = np.zeros(n_req + 1, np.int32)
        ft = np.zeros(n_req * S, np.int32)
        nf, sec = C.c_int(0), C.c_double(0)
        best = None
        for _ in range(2):
            H.check_ref(ref.ref_run_paged_engine(variant, wl["B"], S, wl["d"], wl["V"], wl["n_blocks"], 1,
                                                 H.p(w["emb"]), H.p(w["pos"]), H.p(w["wk"]), H.p(w["wq"]),
                                                 H.p(w["wv"]), n_req, H.p(offs), H.p(toks), H.p(ids),
                                                 H.p(fo), H.p(ft), C.byref(nf), C.byref(sec)))
            gen = int(fo[nf.value]) - int(offs[-1])
            tps = gen / sec.value
            best = tps if best is None else max(best, tps)
        out[name + "_tok_s"] = best
        out[name + "_tokens"] = gen
    out["note"] = ("reference engines replay quirk Q1 (stale lengths): their rows attend over at most the "
                   "prompt length; same prompts/weights as our arm")
    return out


def long_context_leg(ctx, torch, hbm_peak):
    """fused decode attention alone at a BASELINE configs[2]-like shape (HBM-bound regime):
    B=1024, d=2048, context lengths U[64,2048]; inputs (26 GB of KV pages) far larger than L2"""
    import harness as H
    import min_llm_inference_b200 as mli
    B, S, d = 1024, 2048, 2048
    rng = np.random.default_rng(7)
    L = rng.integers(64, S, size=B).astype(np.int32)
    W = S // 16
    need = (L + 15) // 16
    n_pages = int(need.sum())
    page_floats = 16 * 3 * d
    try:
        pool = torch.empty((n_pages, page_floats), device="cuda", dtype=torch.float32)
    except Exception as e:  # not enough memory on a shared box
        return {"skipped": str(e)[:80]}
    for i in range(0, n_pages, 4096):
        pool[i:i + 4096].uniform_(-1.0, 1.0)
    perm = rng.permutation(n_pages)
    tab = np.zeros((B, W), np.uint64)
    k = 0
    for r in range(B):
        ids = perm[k:k + need[r]]
        tab[r, :need[r]] = np.uint64(pool.data_ptr()) + ids.astype(np.uint64) * np.uint64(page_floats * 4)
        k += need[r]
    dtab = torch.from_numpy(tab.view(np.int64)).cuda()
    dL = torch.from_numpy(L).cuda()
    q = (torch.rand((B, d), device="cuda") - 0.5) * 0.1
    out = torch.empty((B, d), device="cuda")
    stream = torch.cuda.current_stream()
    for _ in range(3):
        ctx.call("mli_decode_attention_paged", q, dtab, dL, out, None, B, S, d)
    torch.cuda.synchronize()
    n = 10
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(n):
        ctx.call("mli_decode_attention_paged", q, dtab, dL, out, None, B, S, d)
    t1.record(stream)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n
    nbytes = float(np.sum(8.0 * d * L + 8.0 * d + 8.0 * need + 4.0))
    gbs = nbytes / ms / 1e6
    del pool
    return {"workload": "fused decode attention alone, B=1024, d=2048, L~U[64,2048] (BASELINE configs[2] shape)",
            "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
            "ms_per_launch": ms, "algorithmic_bytes": nbytes,
            "kernel": "decode_attention_wp_kernel (warp-per-position consumers; the auto rule picks it for launches "
                      "that can hold >= 1024 positions per SM)",
            "note": "one fused launch per call, 10 calls back to back between one CUDA-event pair"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / reference CUDA / long-context legs")
    ap.add_argument("--no-large", action="store_true", help="skip the engine jobs at the configs[2] / configs[3] shapes")
    ap.add_argument("--pdl", type=int, default=-1, help="override MLI_OPT_PDL (1 programmatic dependent launch, 0 off)")
    ap.add_argument("--gemm-mode", type=int, default=-1, help="override MLI_OPT_GEMM_MODE (0 tcgen05, 1 SIMT exact)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import harness as H
    import min_llm_inference_b200 as mli

    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOAD
    B, S, d, V = wl["B"], wl["S"], wl["d"], wl["V"]
    hbm_peak, peak_src = peaks()

    ctx = mli.Context(local_rank, torch.cuda.current_stream().cuda_stream)
    if args.gemm_mode >= 0:
        ctx.set_option(mli.OPT_GEMM_MODE, args.gemm_mode)
    gemm_mode = ctx.get_option(mli.OPT_GEMM_MODE)
    if args.pdl >= 0:
        ctx.set_option(mli.OPT_PDL, args.pdl)

    # weights are replicated (every rank regenerates them from the seed); the global request set is
    # n_req * world requests from one seed, sharded by contiguous blocks (weak scaling)
    from min_llm_inference_b200.sharding import gather_tokens, shard_requests
    w = H.make_weights(1001, d, V, S, "Z")
    g_offs, g_toks = H.make_prompts(2002, wl["n_req"] * world, wl["lo"], wl["hi"])
    offs, toks, _ = shard_requests(g_offs, g_toks, rank, world)
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(B, S, d, V, wl["n_blocks"], wl["R"], 0, wl["n_req"], None)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    d_offs, d_toks = torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda()
    p_offs = torch.from_numpy(offs).pin_memory()
    p_toks = torch.from_numpy(toks).pin_memory()
    tok_buf = torch.zeros((wl["n_req"], S), dtype=torch.int32, device="cuda")
    cnt_buf = torch.zeros((wl["n_req"],), dtype=torch.int32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def job_device():
        """one step, prompts resident in HBM; returns (device ms, tokens generated)"""
        eng.submit(d_offs, d_toks, is_device=True)
        eng.run()
        st = eng.stats()
        ms = st.gpu_ms
        if world > 1:   # the only collective of the path: final token gather
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.copy_tokens(tok_buf, cnt_buf)
            gather_tokens(tok_buf, cnt_buf, wl["n_req"] * world)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        return ms, st.generated_tokens

    def job_e2e():
        """one step through the host-buffer C ABI: prompts H2D, finished token lists D2H"""
        eng.submit(p_offs.numpy(), p_toks.numpy(), is_device=False)
        eng.run()
        res, order = eng.results()
        return eng.stats().generated_tokens, len(order)

    for _ in range(args.warmup):
        job_device()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms, gen_total = 0.0, 0
    for _ in range(args.steps):
        ms, gen = job_device()
        dev_ms += ms
        gen_total += gen
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    st = eng.stats()

    # e2e (wall clock, host buffers)
    job_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_gen = 0
    for _ in range(args.steps):
        g, nfin = job_e2e()
        e2e_gen += g
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = 4 * (len(offs) + len(toks))
    d2h = 4 * (nfin + wl["n_req"] + wl["n_req"] * S) + 128

    # roofline pass: every fused-attention launch of one job bracketed by CUDA events
    eng.submit(d_offs, d_toks, is_device=True)
    eng.run(profile_attention=True)
    ps = eng.stats()
    attn_gbs = ps.attn_bytes / max(ps.attn_ms, 1e-9) / 1e6
    traffic, traffic_info = attention_traffic()
    if traffic_info:
        traffic_info = {"algorithmic_bytes_of_captured_launch": traffic_info["algorithmic_bytes_per_launch"],
                        "ratio": traffic_info["traffic_over_algorithmic"], "source": "profiles/r1_attn_traffic.json"}

    # reduce over ranks: max time, sum tokens
    stats_t = torch.tensor([dev_ms, wall * 1e3, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    toks_t = torch.tensor([float(gen_total), float(e2e_gen), float(launches)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(toks_t, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, e2e_ms_max = (float(x) for x in stats_t.tolist())
    gen_all, e2e_gen_all, launches_all = (float(x) for x in toks_t.tolist())

    if rank == 0:
        line = {
            "metric": METRIC, "value": gen_all / (dev_ms_max / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["name"], "n_batch": B, "emb_dim": d, "n_sequence": S, "n_vocab": V,
                       "kv_pages": wl["n_blocks"], "requests_per_gpu": wl["n_req"], "n_forward_rounds": 1,
                       "distribution": "Z (zero-mean), fixed seeds", "lengths": "corrected (no Q1 replay)",
                       "gemm_mode": "tcgen05 3xTF32" if gemm_mode == 0 else "SIMT fp32 exact-order",
                       "pdl": ctx.get_option(mli.OPT_PDL),
                       "l2": "inputs larger than L2 (KV pool 201 MB + tables > 126 MB); no flush",
                       "step_graph": "6 kernels per engine iteration (scheduler, encoder, merged QKV+prefill GEMM, "
                                     "fused attention, split-K logits GEMM, decoder), 4 iterations per CUDA graph",
                       "parallelism": f"request-sharded dp{world}",
                       "step": "one whole engine job (512 requests per GPU to completion)"},
            "clocks": clocks,
            "e2e": {"value": e2e_gen_all / (e2e_ms_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "timed_by": "wall clock, host buffers through the C ABI"},
            "gpu_launches": int(launches_all),
            "tokens_per_step": gen_all / args.steps, "engine_iterations_per_step": st.steps,
            "preemptions_per_step": st.preemptions, "wall_ms_per_step": wall_ms_max / args.steps,
            "roofline": {"kernel": "decode_attention_kernel (one fused launch: qkt + masked softmax + softmax_v, "
                                   "flattened position-space slices, last-arriver merge)",
                         "bound": "hbm", "achieved": attn_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": attn_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "launches": ps.attn_launches, "avg_launch_us": 1e3 * ps.attn_ms / max(1, ps.attn_launches),
                         "algorithmic_bytes_per_launch": ps.attn_bytes / max(1, ps.attn_launches),
                         "traffic_capture": traffic_info,
                         "note": "ATTN_BYTES (SURVEY 8d) / CUDA-event time of every attention launch of one job "
                                 "(un-captured pass, an event pair per launch).  In-job launches move only ~70-100 MB "
                                 "(14 us at peak), so fixed costs dominate; the same kernel at the configs[2] shape is "
                                 "in roofline_long_context"},
        }
        if ps.gemm_launches > 0:
            bf16_peak, tsrc = tensor_peak()
            tflops = ps.gemm_flops / max(ps.gemm_ms, 1e-9) / 1e9
            line["roofline_gemm"] = {
                "kernel": "gemm_tf32x3_kernel (merged latest-token QKV + prefill projection, 3xTF32 on tcgen05)",
                "bound": "tensor", "achieved": 3.0 * tflops, "peak": bf16_peak / 2.0, "unit": "TFLOP/s",
                "frac": 3.0 * tflops / (bf16_peak / 2.0), "fp32_equivalent_tflops": tflops,
                "peak_source": tsrc + ": dense bf16 sustained / 2 = tf32",
                "launches": ps.gemm_launches, "avg_launch_us": 1e3 * ps.gemm_ms / ps.gemm_launches,
                "algorithmic_flops_per_launch": ps.gemm_flops / ps.gemm_launches,
                "note": "achieved = 3 x fp32-equivalent FLOPs (each product is three tf32 MMAs) / CUDA-event time of "
                        "every launch of one job.  A decode-size launch is ~1 GFLOP (150-250 rows): it is bound by "
                        "fixed costs, not by the tensor pipe (ncu: pipe active 41 % on the busiest SM during the "
                        "main loop's share of the kernel, profiles/r1_gemm_ncu.json); by launch share of the step the "
                        "two GEMMs (46 %) exceed the attention (31 %)"}
        if world == 1 and not args.no_extras:
            try:
                # its own context: the engine's captured graph pins the first context's workspaces
                ctx2 = mli.Context(local_rank, torch.cuda.current_stream().cuda_stream)
                line["roofline_long_context"] = long_context_leg(ctx2, torch, hbm_peak)
                ctx2.close()
            except Exception as e:  # never lose the headline line to an extra
                line["roofline_long_context"] = {"error": str(e)[:200]}
            try:
                with StdoutToStderr():
                    line["reference_cuda"] = reference_cuda_leg(w, offs, toks)
            except Exception as e:
                line["reference_cuda"] = {"error": str(e)[:200]}
            try:
                line["cpu_baseline"] = cpu_baseline_leg()
            except Exception as e:
                line["cpu_baseline"] = {"error": str(e)[:200]}
            try:
                # in-graph time of every kernel of the step (globaltimer stamps), and the attention
                # roofline recomputed with the in-graph time instead of the event-bracketed eager launch
                sys.path.insert(0, str(REPO / "tools"))
                import step_timeline
                _, tl = step_timeline.measure(local_rank, ctx.get_option(mli.OPT_PDL))
                line["step_timeline"] = tl
                # the last step of the job has no successor stamp: scale the bytes to the steps timed
                frac_steps = tl["attention_steps"] / max(1, ps.attn_launches)
                in_graph_gbs = ps.attn_bytes * frac_steps / max(tl["attention_total_us"], 1e-9) / 1e3
                line["roofline"]["in_graph"] = {
                    "avg_launch_us": tl["attention_total_us"] / max(1, tl["attention_steps"]),
                    "achieved": in_graph_gbs, "frac": in_graph_gbs / hbm_peak, "timed_by": tl["timed_by"]}
            except Exception as e:
                line["step_timeline"] = {"error": str(e)[:200]}
            if not args.no_large:
                # the same engine in the HBM-bound regime (not the headline workload): whole jobs at the
                # BASELINE configs[2] shape and a 32k-context decode at the configs[3] shape
                sys.path.insert(0, str(REPO / "tools"))
                eng.close()
                for key, preset, kvb in (("engine_configs2", "c3", 0), ("engine_configs3", "c4", 0),
                                         ("engine_configs2_compact_kv", "c3", 1)):   # opt-in bf16 K/V pages
                    try:
                        import run_config
                        torch.cuda.empty_cache()
                        line[key] = run_config.run(preset, local_rank, reps=1, kv_bf16=kvb)
                    except Exception as e:
                        line[key] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    eng.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
