#!/usr/bin/env python
"""bench.py -- decode tokens/s of the paged-attention decode path (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload c4|c2a]

Workload (config.workload): BASELINE.json configs[4], the configuration the metric ("decode tokens/sec at
1/2/4/8 B200") is quoted on -- the request-sharded sweep: 8192 requests / batch rows IN TOTAL, emb_dim 1024,
vocab 1024, page_size 16, prompts U[64,2048], every request generates up to 128 tokens (EOF ends it
earlier), n_sequence 2304, n_forward_rounds 1.  With N GPUs every rank owns 8192/N rows and 8192/N requests
of the SAME global request set (fixed seed): STRONG scaling.  KV pages are sized so that every request of
the rank is resident (about 116 GB of fp32 pages at N=1).  Synthetic, fixed seeds, zero-mean weights (dist
"Z": the reference's U(0,1] fixtures make softmax one-hot and every row emit the same token; SURVEY
finding 9), corrected lengths (the reference's quirk Q1 is replayed only in the parity tests and in the
reference-CUDA comparison legs).

A "step" = one whole job: all requests of the rank admitted, prefilled, decoded and retired by the
on-device scheduler, then the final token lists all-gathered over NCCL (mli_comm_gather_tokens; N > 1).
`value` = generated tokens of all ranks / device time (CUDA events around submit .. gather, max over
ranks) with the prompts already resident in HBM; `e2e` = the same job through the host-buffer C ABI
(prompts H2D from pinned memory, gather, finished token lists D2H) by wall clock.

`--impl reference` times the reference's own HOST implementation of the path (oracle/_ref: the loops of
tests/test_utils.cpp driven by the reference's scheduler; single-threaded as written, so one process per
host core, each on its own requests -- the same request sharding the GPUs use) on a bounded sample of the
same workload.
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

WORKLOADS = {
    # BASELINE.json configs[4] (the headline): totals over all GPUs
    "c4": dict(name="BASELINE.json configs[4]: request-sharded paged attention, n_batch=8192 total, emb_dim=1024, "
                    "n_vocab=1024, page_size=16, n_sequence=2304, 8192 requests, prompts U[64,2048], up to 128 new "
                    "tokens per request, KV pool holds every resident request",
               B=8192, S=2304, d=1024, V=1024, n_req=8192, lo=64, hi=2048, R=1, max_new=128, n_blocks=0),
    # BASELINE.json configs[1] (round 1's headline; launch-latency-bound): per GPU, weak scaling
    "c2a": dict(name="BASELINE.json configs[1]: paged attention, n_batch=256, emb_dim=1024, n_sequence=128, "
                     "page_size=16, pool=1024 pages, 512 requests, prompts U[1,64] (per GPU)",
                B=256, S=128, d=1024, V=1024, n_req=512, lo=1, hi=64, R=1, max_new=0, n_blocks=1024),
}
METRIC = "decode tokens/sec (paged attention, continuous batching)"
UNIT = "tokens/s"
SEED_W, SEED_P = 1001, 2002


def workload_config(wl, world):
    """the `config` object both arms print (identical keys and values for a given --gpus)"""
    strong = wl["n_blocks"] == 0
    return {"workload": wl["name"], "n_batch_total": wl["B"] if strong else wl["B"] * world, "emb_dim": wl["d"],
            "n_sequence": wl["S"], "n_vocab": wl["V"], "requests_total": wl["n_req"] if strong else wl["n_req"] * world,
            "prompt_lengths": f"U[{wl['lo']},{wl['hi']}]", "max_new_tokens": wl["max_new"], "n_forward_rounds": wl["R"],
            "distribution": "Z (zero-mean), fixed seeds", "lengths": "corrected (no Q1 replay)",
            "parallelism": f"request-sharded dp{world}" + (" (prompt-length-balanced shards)" if strong else ""),
            "step": "one whole job (every request of the rank admitted, prefilled, decoded, retired; token gather)"}


def peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured copy bandwidth (MEASURED_PEAKS.json)"
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
        try:
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def rank_requests(wl, rank, world):
    """(local offsets, local tokens) of this rank: strong scaling shards ONE global request set"""
    import harness as H
    from min_llm_inference_b200.sharding import shard_requests, shard_requests_balanced
    strong = wl["n_blocks"] == 0
    n_total = wl["n_req"] if strong else wl["n_req"] * world
    g_offs, g_toks = H.make_prompts(SEED_P, n_total, wl["lo"], wl["hi"])
    # strong scaling: the router balances prompt tokens between the ranks (the job time is the max over ranks)
    split = shard_requests_balanced if strong else shard_requests
    offs, toks, _ = split(g_offs, g_toks, rank, world)
    return offs, toks, n_total


# ---------------------------------------------------------------------------------------------
# --impl reference: the reference's host loops, one process per core
# ---------------------------------------------------------------------------------------------
_W = {}


def _ref_worker_init(d, V, S, use_ref):
    sys.stdout = sys.stderr
    os.dup2(2, 1)   # the reference printf()s its throughput counter
    import harness as H
    _W["H"] = H
    _W["w"] = H.make_weights(SEED_W, d, V, S, "Z")
    _W["ref"] = H.load_ref() if use_ref else None
    _W["dims"] = (d, V, S)


def _ref_worker_run(task):
    offs, toks, iters = task
    H, w = _W["H"], _W["w"]
    d, V, S = _W["dims"]
    rows = len(offs) - 1
    if rows == 0:
        return 0, 0.0
    gen, steps, sec = C.c_longlong(0), C.c_longlong(0), C.c_double(0)
    if _W["ref"] is not None:
        H.check_ref(_W["ref"].ref_run_host_engine(rows, S, d, V, H.p(w["emb"]), H.p(w["pos"]), H.p(w["wk"]),
                                                  H.p(w["wq"]), H.p(w["wv"]), rows, H.p(offs), H.p(toks), iters,
                                                  C.byref(gen), C.byref(steps), C.byref(sec), None, None, None, None))
        return gen.value, sec.value
    cfg = dict(B=rows, S=S, d=d, V=V, n_blocks=rows * (S // 16), R=1)
    t0 = time.perf_counter()
    rc, _, _, st = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1, max_steps=iters, threads=1)
    return st.generated_tokens, time.perf_counter() - t0


def reference_arm(args, rank, world, wl):
    if rank != 0:
        return
    import multiprocessing as mp
    import harness as H
    cores = os.cpu_count() or 1
    iters = 4
    offs, toks, n_total = rank_requests(wl, 0, 1)
    # bounded sample: the first `cores` requests of the global set, one per host core, prefill + `iters`
    # decode iterations each (the reference's host prefill of ONE 1056-token prompt at emb_dim 1024 is
    # ~9 GFLOP of scalar loops; the whole 8192-request job would take hours)
    n_sample = min(cores, n_total)
    # ... taken at the MEDIAN prompt length of the set (ties by id): every core then works for the whole step (with
    # the first `cores` requests the step lasted as long as its one ~2000-token prompt and the other cores idled,
    # which both understated the reference and made a 20-step run take six minutes)
    plen = np.diff(offs)
    pick = np.argsort(np.abs(plen - np.median(plen)), kind="stable")[:n_sample]
    tasks = []
    for k in pick.tolist():
        o = (offs[k:k + 2] - offs[k]).astype(np.int32)
        tasks.append((o, toks[offs[k]:offs[k + 1]].copy(), iters))
    have_ref = H.ref_available()
    if have_ref:
        try:
            import torch
            have_ref = torch.cuda.is_available()   # the reference's host tensors are cudaHostAlloc'd
        except Exception:
            have_ref = False
    kind = "reference" if have_ref else "port"
    sample = (f"the {n_sample} requests of {n_total} closest to the median prompt length ({int(np.median(plen))} tokens), one "
              f"per host core ({cores} processes), prefill + {iters} engine iterations per step; same emb_dim / "
              f"n_sequence / vocab as the GPU arm")
    ctx = mp.get_context("spawn")
    times, gens = [], []
    with ctx.Pool(n_sample, initializer=_ref_worker_init, initargs=(wl["d"], wl["V"], wl["S"], have_ref)) as pool:
        # page in libraries / contexts on an 8-token prompt
        pool.map(_ref_worker_run, [(np.array([0, 8], np.int32), t[1][:8].copy(), 1) for t in tasks])
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker_run, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
                gens.append(sum(r[0] for r in res))
    total_t, total_g = float(sum(times)), float(sum(gens))
    value = total_g / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong" if wl["n_blocks"] == 0 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(wl, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_sample, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def cpu_baseline_leg(wl):
    """oracle port (OpenMP over rows) on all host cores, bounded sample of the same workload"""
    import harness as H
    cores = os.cpu_count() or 1
    iters = 8
    n_sample = min(4 * cores, wl["n_req"])
    w = H.make_weights(SEED_W, wl["d"], wl["V"], wl["S"], "Z")
    offs, toks, n_total = rank_requests(wl, 0, 1)
    offs_s = offs[:n_sample + 1].copy()
    toks_s = toks[:offs_s[-1]].copy()
    W = wl["S"] // 16
    cfg = dict(B=n_sample, S=wl["S"], d=wl["d"], V=wl["V"], n_blocks=n_sample * W, R=1, max_new=wl["max_new"])
    t0 = time.perf_counter()
    rc, _, _, st = H.run_oracle_engine("paged", cfg, w, offs_s, toks_s, fix=1, max_steps=iters, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": st.generated_tokens / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n_sample} of {n_total} requests on {n_sample} rows, prefill + {iters} engine iterations "
                      f"({st.generated_tokens} tokens in {dt:.1f} s), OpenMP over rows"}


def reference_cuda_pairs(ctx, torch):
    """The reference's CUDA engines (rebuilt for sm_100) and OUR engine on the same tensors, at the two
    shapes the reference can run: BASELINE configs[1] (C2a) and the reference's own profile workload
    tests/paged_for_profile.cpp:11-23 (C2b, the only shape it publishes numbers for: README.md:66-82).
    Their metric: generated tokens / wall time including host scheduling (src/throughput_counter.cpp).
    Both sides replay quirk Q1 here (ours: compat_stale_lengths = 1), i.e. they do the SAME work and, in
    exact GEMM mode, produce identical tokens.  Median of 5 runs after a warm-up (cuBLAS init)."""
    import harness as H
    import min_llm_inference_b200 as mli
    if not H.ref_available():
        return None
    ref = H.load_ref()
    shapes = {
        # BASELINE configs[0]: the NON-paged engine (start_inference_engine, src/inferencer.cpp:11-41; no Q1 quirk
        # there), ours = the same device engine with a pool that cannot run dry (what the C++ drop-in
        # start_inference_engine runs, host/src/engine.cpp)
        "configs0_C1_dense": dict(B=32, S=256, d=256, V=1024, n_blocks=32 * 16, n_req=96, lo=1, hi=128, dist="Z",
                                  eof=1.0001, dense=True),
        # the non-paged mode at a larger shape
        "dense_larger_shape": dict(B=128, S=512, d=1024, V=1024, n_blocks=128 * 32, n_req=256, lo=32, hi=384, dist="Z",
                                   eof=1.0001, dense=True),
        "configs1_C2a": dict(B=256, S=128, d=1024, V=1024, n_blocks=1024, n_req=512, lo=1, hi=64, dist="Z", eof=1.0001),
        "paged_for_profile_C2b": dict(B=1024, S=128, d=2048, V=1024, n_blocks=4096, n_req=2048, lo=1, hi=64, dist="R",
                                      eof=1.0001),
    }
    out = {}
    for key, c in shapes.items():
        w = H.make_weights(SEED_W, c["d"], c["V"], c["S"], c["dist"], eof_ratio=c["eof"])
        offs, toks = H.make_prompts(SEED_P, c["n_req"], c["lo"], c["hi"])
        n_req, S = c["n_req"], c["S"]
        res = {"shape": {k: c[k] for k in ("B", "S", "d", "V", "n_blocks", "n_req", "lo", "hi", "dist")}}
        dense = c.get("dense", False)
        for name, variant, runs in ((("dense_engine", -1, 5),) if dense else
                                    (("warp_tiling_cublas", 1, 5), ("naive", 0, 3))):
            ids = np.zeros(n_req, np.int32)
            fo = np.zeros(n_req + 1, np.int32)
            ft = np.zeros(n_req * S, np.int32)
            nf, sec = C.c_int(0), C.c_double(0)
            walls, evs, gen = [], [], 0
            for i in range(runs + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if dense:
                    H.check_ref(ref.ref_run_dense_engine(c["B"], S, c["d"], c["V"], H.p(w["emb"]), H.p(w["pos"]),
                                                         H.p(w["wk"]), H.p(w["wq"]), H.p(w["wv"]), n_req, H.p(offs),
                                                         H.p(toks), H.p(ids), H.p(fo), H.p(ft), C.byref(nf),
                                                         C.byref(sec)))
                else:
                    H.check_ref(ref.ref_run_paged_engine(variant, c["B"], S, c["d"], c["V"], c["n_blocks"], 1,
                                                         H.p(w["emb"]), H.p(w["pos"]), H.p(w["wk"]), H.p(w["wq"]),
                                                         H.p(w["wv"]), n_req, H.p(offs), H.p(toks), H.p(ids),
                                                         H.p(fo), H.p(ft), C.byref(nf), C.byref(sec)))
                e1.record()
                torch.cuda.synchronize()
                gen = int(fo[nf.value]) - int(offs[-1])
                if i > 0:   # run 0 = warm-up
                    walls.append(sec.value)
                    evs.append(e0.elapsed_time(e1) / 1e3)
            res[name] = {"tokens": gen, "tok_s_median_wall": gen / float(np.median(walls)),
                         "tok_s_best_wall": gen / min(walls), "tok_s_worst_wall": gen / max(walls),
                         "tok_s_median_device_events": gen / float(np.median(evs)), "runs": runs}
        dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
        for name, compat in ((("ours_device_engine", 0),) if dense else
                             (("ours_same_work_q1_replayed", 1), ("ours_corrected_lengths", 0))):
            ec = mli.EngineCfg(c["B"], S, c["d"], c["V"], c["n_blocks"], 1, compat, n_req, None)
            eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
            walls, gen = [], 0
            for i in range(6):
                t0 = time.perf_counter()
                eng.submit(offs, toks)
                eng.run()
                eng.results()
                dt = time.perf_counter() - t0
                gen = eng.stats().generated_tokens
                if i > 0:
                    walls.append(dt)
            eng.close()
            res[name] = {"tokens": int(gen), "tok_s_median_wall": gen / float(np.median(walls)), "runs": 5,
                         "timed": "wall clock: host prompts in, finished token lists out (the reference's metric)"}
        if dense:
            res["speedup_vs_reference_dense_engine"] = (res["ours_device_engine"]["tok_s_median_wall"] /
                                                        res["dense_engine"]["tok_s_median_wall"])
        else:
            res["speedup_same_work_vs_warp_tiling_cublas"] = (res["ours_same_work_q1_replayed"]["tok_s_median_wall"] /
                                                              res["warp_tiling_cublas"]["tok_s_median_wall"])
        out[key] = res
    out["note"] = ("reference engines replay quirk Q1 (stale lengths): rows attend over at most the prompt length and "
                   "re-emit their first token until n_sequence; 'ours_same_work' replays it too (identical token lists "
                   "in exact mode, tests/test_gpu_forward_engine.py P3), 'ours_corrected' is the fixed behaviour")
    return out


def attention_traffic(wl_key):
    """DRAM bytes per attention launch of THIS workload's job, from the committed ncu pass
    (profiles/r2_attn_traffic_<workload>.json: dram__bytes_read.sum + dram__bytes_write.sum of every
    decode-attention launch of one job, tools/attn_traffic.py); None when no capture of this workload exists"""
    f = REPO / "profiles" / f"r2_attn_traffic_{wl_key}.json"
    if not f.exists():
        return None, None
    t = json.loads(f.read_text())
    return t.get("dram_bytes_per_launch_mean"), t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / reference CUDA / other-config legs")
    ap.add_argument("--no-large", action="store_true", help="skip the engine jobs at the configs[2] / configs[3] shapes")
    ap.add_argument("--pdl", type=int, default=-1, help="override MLI_OPT_PDL (1 programmatic dependent launch, 0 off)")
    ap.add_argument("--gemm-mode", type=int, default=-1, help="override MLI_OPT_GEMM_MODE (0 tcgen05, 1 SIMT exact)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, rank, world, wl)
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
    strong = wl["n_blocks"] == 0
    S, d, V = wl["S"], wl["d"], wl["V"]
    B = wl["B"] // world if strong else wl["B"]
    hbm_peak, peak_src = peaks()

    ctx = mli.Context(local_rank, torch.cuda.current_stream().cuda_stream)
    if args.gemm_mode >= 0:
        ctx.set_option(mli.OPT_GEMM_MODE, args.gemm_mode)
    gemm_mode = ctx.get_option(mli.OPT_GEMM_MODE)
    if args.pdl >= 0:
        ctx.set_option(mli.OPT_PDL, args.pdl)

    # weights are replicated (every rank regenerates them from the seed); requests are one global set from
    # one seed, sharded by contiguous blocks
    w = H.make_weights(SEED_W, d, V, S, "Z")
    offs, toks, n_total = rank_requests(wl, rank, world)
    n_local = len(offs) - 1
    per_rank = (n_total + world - 1) // world
    if strong:
        # every request of the rank resident: pages for prompt + new tokens (+ the admission minimum of 4)
        plen = np.diff(offs).astype(np.int64)
        n_blocks = int(np.maximum((plen + wl["max_new"] + 1 + 15) // 16, 4).sum()) + 64
    else:
        n_blocks = wl["n_blocks"]
    page_bytes = 16 * 3 * d * 4
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    ec = mli.EngineCfg(B, S, d, V, n_blocks, wl["R"], 0, per_rank, None, wl["max_new"], 0)
    eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
    d_offs, d_toks = torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda()
    p_offs = torch.from_numpy(offs).pin_memory()
    p_toks = torch.from_numpy(toks).pin_memory()
    comm = None
    all_tok = all_cnt = None
    if world > 1:
        # the unique id travels over the torch.distributed bootstrap; the collective itself is ours (C ABI)
        box = [mli.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = mli.Comm(ctx, world, rank, box[0])
        all_tok = torch.empty((world * per_rank, S), dtype=torch.int32, device="cuda")
        all_cnt = torch.empty((world * per_rank,), dtype=torch.int32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def job_device():
        """one step, prompts resident in HBM; returns (device ms, tokens generated)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.submit(d_offs, d_toks, is_device=True)
        eng.run()
        if comm is not None:   # the only collective of the path: final token gather
            comm.gather_tokens(eng, per_rank, all_tok, all_cnt)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), eng.stats().generated_tokens

    def job_e2e():
        """one step through the host-buffer C ABI: prompts H2D, token gather, finished token lists D2H"""
        eng.submit(p_offs.numpy(), p_toks.numpy(), is_device=False)
        eng.run()
        if comm is not None:
            comm.gather_tokens(eng, per_rank, all_tok, all_cnt)
        res, order = eng.results()
        if comm is not None:
            torch.cuda.synchronize()
        return eng.stats().generated_tokens, len(order), sum(len(v) for v in res.values())

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
        g, nfin, ntok = job_e2e()
        e2e_gen += g
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = 4 * (len(offs) + len(toks))
    d2h = 4 * (2 * nfin + 1 + ntok) + 128

    # roofline pass: every fused-attention launch (and every merged GEMM launch) of one job bracketed by
    # CUDA events on the engine's stream
    eng.submit(d_offs, d_toks, is_device=True)
    eng.run(profile_attention=True)
    ps = eng.stats()
    attn_gbs = ps.attn_bytes / max(ps.attn_ms, 1e-9) / 1e6
    traffic, traffic_info = attention_traffic(args.workload) if world == 1 else (None, None)

    # reduce over ranks: max time, sum tokens
    stats_t = torch.tensor([dev_ms, wall * 1e3, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    toks_t = torch.tensor([float(gen_total), float(e2e_gen), float(launches), float(st.preemptions)], device="cuda",
                          dtype=torch.float64)
    occ_t = torch.tensor([float(st.peak_resident_rows), 1.0 - st.min_free_pages / n_blocks, n_blocks * page_bytes / 1e9,
                          float(st.preemptions), float(st.steps)], device="cuda", dtype=torch.float64)
    occ_all = [occ_t.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(stats_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(toks_t, op=dist.ReduceOp.SUM)
        dist.all_gather(occ_all, occ_t)
    dev_ms_max, wall_ms_max, e2e_ms_max = (float(x) for x in stats_t.tolist())
    gen_all, e2e_gen_all, launches_all, preempt_all = (float(x) for x in toks_t.tolist())

    if rank == 0:
        attn_kernel = ("decode_attention_wp_kernel (one fused launch: qkt + masked softmax + softmax_v; warp-per-position "
                       "consumers, flattened position-space slices, last-arriver merge)" if B * S >= 1024 * 148 else
                       "decode_attention_kernel (one fused launch: qkt + masked softmax + softmax_v; column-split consumers)")
        cfg = workload_config(wl, world)
        line = {
            "metric": METRIC, "value": gen_all / (dev_ms_max / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "details": {"rows_per_gpu": B, "requests_per_gpu": n_local,
                        "gemm_mode": "tcgen05 3xTF32" if gemm_mode == 0 else "SIMT fp32 exact-order",
                        "pdl": ctx.get_option(mli.OPT_PDL),
                        "l2": "inputs larger than L2 (KV pages read per decode step >> 126 MB); no flush",
                        "step_graph": "6 kernels per engine iteration (scheduler, encoder, merged QKV+prefill GEMM, fused "
                                      "attention, split-K logits GEMM, decoder), 4 iterations per CUDA graph",
                        "collective": "mli_comm_gather_tokens (NCCL all-gather of the request tables) inside value and e2e"
                                      if world > 1 else "none (single GPU)",
                        "per_gpu": [{"peak_resident_rows": int(o[0]), "peak_pool_occupancy": round(float(o[1]), 4),
                                     "kv_pool_gb": round(float(o[2]), 2), "preemptions": int(o[3]),
                                     "engine_iterations": int(o[4])} for o in (x.tolist() for x in occ_all)]},
            "clocks": clocks,
            "e2e": {"value": e2e_gen_all / (e2e_ms_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "timed_by": "wall clock, host buffers through the C ABI, per rank; "
                                                           "max over ranks"},
            "gpu_launches": int(launches_all),
            "tokens_per_step": gen_all / args.steps, "engine_iterations_per_step": st.steps,
            "preemptions_per_step": preempt_all, "wall_ms_per_step": wall_ms_max / args.steps,
            "roofline": {"kernel": attn_kernel,
                         "bound": "hbm", "achieved": attn_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": attn_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "launches": ps.attn_launches, "avg_launch_us": 1e3 * ps.attn_ms / max(1, ps.attn_launches),
                         "algorithmic_bytes_per_launch": ps.attn_bytes / max(1, ps.attn_launches),
                         "traffic_capture": traffic_info,
                         "frac_of_nominal_8TBps": attn_gbs / 8000.0,
                         "share_of_job_time": ps.attn_ms / max(ps.gpu_ms, 1e-9),
                         "note": "ATTN_BYTES (SURVEY 8d) of every attention launch of one job / the sum of their CUDA-event "
                                 "times (un-captured pass, an event pair per launch on the engine's stream, rank 0).  frac is "
                                 "against the driver-measured COPY bandwidth (read + write traffic); a read-only stream can "
                                 "exceed it, so frac > 1 is possible -- frac_of_nominal_8TBps is the fraction of the HBM3e "
                                 "spec number"},
        }
        if ps.gemm_launches > 0:
            bf16_peak, tsrc = tensor_peak()
            tflops = ps.gemm_flops / max(ps.gemm_ms, 1e-9) / 1e9
            line["roofline_gemm"] = {
                "kernel": "gemm_tf32x3_kernel (merged latest-token QKV + prefill projection, 3xTF32 on tcgen05), all launches "
                          "of one job",
                "bound": "tensor", "achieved": 3.0 * tflops, "peak": bf16_peak / 2.0, "unit": "TFLOP/s",
                "frac": 3.0 * tflops / (bf16_peak / 2.0), "fp32_equivalent_tflops": tflops,
                "peak_source": tsrc + ": dense bf16 sustained / 2 = tf32",
                "launches": ps.gemm_launches, "avg_launch_us": 1e3 * ps.gemm_ms / ps.gemm_launches,
                "share_of_job_time": ps.gemm_ms / max(ps.gpu_ms, 1e-9)}
            if ps.gemm_max_ms > 0:
                ptf = ps.gemm_max_flops / ps.gemm_max_ms / 1e9
                line["roofline_prefill"] = {
                    "kernel": "gemm_tf32x3_kernel, the largest launch of the job (bulk prefill of the admitted prompts: the "
                              "tensor-pipe regime)",
                    "bound": "tensor", "achieved": 3.0 * ptf, "peak": bf16_peak / 2.0, "unit": "TFLOP/s",
                    "frac": 3.0 * ptf / (bf16_peak / 2.0), "fp32_equivalent_tflops": ptf, "ms": ps.gemm_max_ms,
                    "algorithmic_flops": ps.gemm_max_flops,
                    "note": "achieved = 3 x fp32-equivalent FLOPs (each product is three tf32 MMAs) / CUDA-event time"}
        if world == 1 and not args.no_extras:
            eng.close()
            eng = None
            torch.cuda.empty_cache()
            try:
                line["cpu_baseline"] = cpu_baseline_leg(wl)
            except Exception as e:
                line["cpu_baseline"] = {"error": str(e)[:200]}
            try:
                with StdoutToStderr():
                    ctx3 = mli.Context(local_rank, torch.cuda.current_stream().cuda_stream)
                    line["reference_cuda"] = reference_cuda_pairs(ctx3, torch)
                    ctx3.close()
            except Exception as e:
                line["reference_cuda"] = {"error": str(e)[:200]}
            if not args.no_large:
                # the same engine at the other BASELINE configurations, at their named shapes
                sys.path.insert(0, str(REPO / "tools"))
                for key, preset, kvb in (("engine_configs1", "c2a", 0), ("engine_configs1_admission_throttle", "c2a_pf", 0),
                                         ("engine_configs1_chunked_prefill", "c2a_chunk", 0),
                                         ("engine_configs1_four_rounds", "c2a_r4", 0),
                                         ("engine_configs2", "c3", 0),
                                         ("engine_configs3", "c4", 0), ("engine_configs2_compact_kv", "c3", 1)):
                    try:
                        import run_config
                        torch.cuda.empty_cache()
                        line[key] = run_config.run(preset, local_rank, reps=1, kv_bf16=kvb)
                    except Exception as e:
                        line[key] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if comm is not None:
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
