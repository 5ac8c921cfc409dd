"""N > 1 on real GPUs (skipped with fewer than two devices): the token gather behind the C ABI
(mli_comm_*, NCCL all-gather of the request tables) and the request-sharded C++ drop-in engine.

  * single process, two contexts, mli_comm_init_all + grouped mli_comm_gather_tokens: the gathered
    table equals the two engines' own results;
  * the drop-in driver (tests/dropin/engine_driver.cpp, written against the reference's API:
    include/inferencer.h:23-32) run with MLI_NUM_GPUS=2 returns the same token list per request id
    as with one GPU;
  * torchrun --nproc-per-node 2 bench.py (multi-process, mli_comm_init_rank) runs and reports.
"""
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
MLI_DRIVER = REPO / "tests" / "dropin" / "_build" / "dropin_driver_mli"


@pytest.fixture(scope="module")
def two_gpus(torch_cuda):
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    return torch_cuda


def test_comm_gather_single_process(two_gpus):
    torch = two_gpus
    lib = mli.load_library()
    cfg = dict(B=8, S=128, d=128, V=1024, n_blocks=64, n_req=24, lo=4, hi=60)
    w = H.make_weights(111, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(113, cfg["n_req"], cfg["lo"], cfg["hi"])
    per = cfg["n_req"] // 2
    ctxs, engs, outs, results = [], [], [], []
    for g in range(2):
        torch.cuda.set_device(g)
        ctx = mli.Context(g, torch.cuda.current_stream().cuda_stream)
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
        dw = {k: torch.from_numpy(v).cuda(g) for k, v in w.items()}
        ec = mli.EngineCfg(cfg["B"], cfg["S"], cfg["d"], cfg["V"], cfg["n_blocks"], 1, 0, per, None, 12, 0)
        eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        lo, hi = g * per, (g + 1) * per
        eng.submit((offs[lo:hi + 1] - offs[lo]).astype(np.int32), toks[offs[lo]:offs[hi]].copy())
        eng.run()
        results.append(eng.results()[0])
        ctxs.append(ctx); engs.append(eng)
        outs.append((torch.zeros((2 * per, cfg["S"]), dtype=torch.int32, device=f"cuda:{g}"),
                     torch.zeros((2 * per,), dtype=torch.int32, device=f"cuda:{g}")))
    arr = (C.c_void_p * 2)(ctxs[0].h, ctxs[1].h)
    comms = (C.c_void_p * 2)()
    assert lib.mli_comm_init_all(arr, 2, comms) == 0, lib.mli_last_error().decode()
    assert lib.mli_comm_group_start() == 0
    for g in range(2):
        rc = lib.mli_comm_gather_tokens(comms[g], engs[g].h, per, outs[g][0].data_ptr(), outs[g][1].data_ptr())
        assert rc == 0, lib.mli_last_error().decode()
    assert lib.mli_comm_group_end() == 0
    for g in range(2):
        ctxs[g].synchronize()
    for g in range(2):
        tab, cnt = outs[g][0].cpu().numpy(), outs[g][1].cpu().numpy()
        for r in range(2):
            for k in range(per):
                want = results[r][k]
                assert cnt[r * per + k] == len(want)
                assert np.array_equal(tab[r * per + k, :len(want)], want)
    for g in range(2):
        lib.mli_comm_destroy(comms[g])
        engs[g].close()
        ctxs[g].close()
    torch.cuda.set_device(0)


def run_driver(args, env):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([str(MLI_DRIVER)] + [str(a) for a in args], capture_output=True, text=True, env=e, timeout=600)
    assert out.returncode == 0, f"driver failed: {out.stderr[-2000:]}\n{out.stdout[-500:]}"
    return sorted(l for l in out.stdout.splitlines() if l.startswith("RESULT "))


@pytest.mark.parametrize("case", [
    ["paged", 16, 128, 256, 1024, 64, 40, 1, 64, 6, "Z"],
    ["paged_cublas", 8, 128, 128, 1024, 36, 25, 20, 64, 8, "Z"],   # odd request count, pool pressure
])
def test_dropin_driver_two_gpus_same_tokens(two_gpus, case):
    if not MLI_DRIVER.exists():
        pytest.skip("drop-in driver not built")
    env = {"MLI_FIX_STALE_LENGTHS": "1", "MLI_GEMM_MODE": "1"}
    one = run_driver(case, env)
    two = run_driver(case, dict(env, MLI_NUM_GPUS="2"))
    assert len(one) == case[6]
    assert one == two, "token lists differ between the 1-GPU and the request-sharded 2-GPU run"


def test_bench_two_ranks(two_gpus):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", str(REPO / "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "3",
           "--workload", "c2a", "--no-extras"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(REPO))
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["e2e"]["value"] > 0
    assert "mli_comm_gather_tokens" in line["details"]["collective"]
