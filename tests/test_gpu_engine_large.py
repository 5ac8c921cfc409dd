"""Engine parity at the REAL dimensions of BASELINE configs[2] (emb_dim 2048, n_sequence 4096, prompts
U[64,2048]) against the reference's own NON-paged CUDA engine (start_inference_engine,
src/inferencer.cpp:11-41 -- the clean end-to-end pin: it has no stale-length quirk) on a request subset that
engine can hold (its dense caches are n_batch * n_sequence * emb_dim floats each).

The reference decodes every request up to n_sequence; our engine is run with max_new_tokens = 48, so the
comparison is over the prompt + the first 48 generated tokens of every request: exact-order mode must be
bit-identical, tensor-core mode may differ only at classified numerical ties."""
import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli
from test_gpu_forward_engine import run_ref_engine

pytestmark = pytest.mark.gpu

CFG = dict(B=12, S=4096, d=2048, V=1024, n_req=18, lo=64, hi=2048, max_new=48)


@pytest.fixture(scope="module")
def reference_tokens(torch_cuda, ref):
    w = H.make_weights(1001, CFG["d"], CFG["V"], CFG["S"], "Z")
    offs, toks = H.make_prompts(2002, CFG["n_req"], CFG["lo"], CFG["hi"])
    theirs, _, sec = run_ref_engine(ref, "dense", CFG, w, offs, toks)
    assert len(theirs) == CFG["n_req"]
    return w, offs, toks, theirs


@pytest.mark.parametrize("gemm_mode", [mli.GEMM_SIMT_EXACT, mli.GEMM_TCGEN05], ids=["exact", "tcgen05"])
def test_engine_at_configs2_dims_vs_reference_dense_engine(torch_cuda, ctx, reference_tokens, gemm_mode):
    torch = torch_cuda
    w, offs, toks, theirs = reference_tokens
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, gemm_mode)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        W = CFG["S"] // 16
        dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
        # pool under pressure on purpose (about 45 % of what 12 full rows would need): growth and
        # pre-emption + re-prefill happen at these dimensions too
        n_blocks = int(0.45 * CFG["B"] * (2048 + 48) / 16)
        ec = mli.EngineCfg(CFG["B"], CFG["S"], CFG["d"], CFG["V"], n_blocks, 1, 0, CFG["n_req"], None, CFG["max_new"], 0)
        eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        eng.submit(offs, toks)
        eng.run()
        mine, order = eng.results()
        st = eng.stats()
        eng.close()
        assert st.n_finished == CFG["n_req"]
        plen = np.diff(offs)
        want = {i: theirs[i][:min(len(theirs[i]), plen[i] + CFG["max_new"])] for i in theirs}
        # the reference stops a request at EOF too: both lists end at the same place unless the cap cut ours
        ties, errors = H.classify_token_mismatches(w, mine, want)
        assert not errors, f"(request, position, margin) beyond a numerical tie: {errors[:4]}"
        if gemm_mode == mli.GEMM_SIMT_EXACT:
            assert not ties, f"exact mode must be bit-identical: {ties}"
        else:
            assert len(ties) <= 2, f"too many tie flips: {ties}"
        print(f"configs[2] dims: {st.generated_tokens} tokens, {st.steps} steps, {st.preemptions} pre-emptions, "
              f"{len(ties)} tie flips")
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


def test_headline_job_at_full_size(torch_cuda, ctx):
    """The bench's own job -- BASELINE configs[4] on one GPU: 8192 rows, 8192 requests U[64,2048], emb_dim 1024,
    n_sequence 2304, 128 new tokens each, ~120 GB of fp32 KV pages, tensor-core mode -- checked where a whole-job
    oracle run is out of reach (hours of CPU):
      * against the CPU oracle on a SAMPLE of its requests (shortest, longest and ten more), each run alone: requests
        are independent with corrected lengths, so their token lists must match the big job's (numerical ties classified);
      * size-independent properties of every request: prompt kept as prefix, at most 128 new tokens, fewer only after
        EOF; nothing pre-empted (the pool holds every request), every page back in the pool at the end;
      * idempotence: the same job again on the same engine gives the same tokens bit for bit;
      * sharding: rank 3 of the 8-GPU run (1024 rows, its balanced deal of requests) reproduces the one-GPU job's
        token lists for its requests."""
    torch = torch_cuda
    import os
    import bench
    torch.cuda.empty_cache()
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 140e9:
        pytest.skip(f"needs ~125 GB of free HBM, {free_b / 1e9:.0f} GB free")
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        wl = dict(bench.WORKLOADS["c4"])
        B, S, d, V, cap = wl["B"], wl["S"], wl["d"], wl["V"], wl["max_new"]
        offs, toks, n_total = bench.rank_requests(wl, 0, 1)
        n_req = len(offs) - 1
        assert (B, n_req, n_total) == (8192, 8192, 8192)
        plen = np.diff(offs).astype(np.int64)
        n_blocks = int(np.maximum((plen + cap + 1 + 15) // 16, 4).sum()) + 64
        w = H.make_weights(bench.SEED_W, d, V, S, "Z")
        dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
        ec = mli.EngineCfg(B, S, d, V, n_blocks, wl["R"], 0, n_req, None, cap, 0, 0)
        eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        runs = []
        for _ in range(2):
            eng.submit(offs, toks)
            eng.run()
            res, order = eng.results()
            st = eng.stats()
            runs.append((res, order.copy(), (st.steps, st.generated_tokens, st.preemptions, st.n_finished,
                                             st.peak_resident_rows, st.min_free_pages)))
        eng.close()
        (mine, order, stats), (again, order2, stats2) = runs
        # idempotence (fixed reduction orders; a page that did not return to the pool would change min_free_pages)
        assert stats == stats2 and np.array_equal(order, order2)
        assert all(np.array_equal(mine[i], again[i]) for i in range(n_req))
        steps, generated, preempt, n_fin, peak_rows, min_free = stats
        assert (n_fin, preempt, peak_rows) == (n_req, 0, B) and min_free >= 64 and steps == cap
        # per-request properties
        total_new = 0
        for i in range(n_req):
            t = mine[i]
            p = toks[offs[i]:offs[i + 1]]
            assert np.array_equal(t[:len(p)], p), f"request {i}: prompt not kept"
            new = len(t) - len(p)
            assert 1 <= new <= cap
            assert new == cap or t[-1] == mli.EOF_TOKEN_ID, f"request {i}: stopped after {new} tokens without EOF"
            assert not np.any(t[len(p):-1] == mli.EOF_TOKEN_ID)
            total_new += new
        assert total_new == generated
        # the oracle on a sample of the requests, each as its own row of a small job
        rng = np.random.default_rng(5)
        sample = sorted(set([int(np.argmin(plen)), int(np.argmax(plen))] + rng.choice(n_req, 10, replace=False).tolist()))
        s_offs = np.zeros(len(sample) + 1, np.int32)
        s_toks = []
        for k, i in enumerate(sample):
            s_toks.append(toks[offs[i]:offs[i + 1]])
            s_offs[k + 1] = s_offs[k] + len(s_toks[-1])
        s_toks = np.concatenate(s_toks).astype(np.int32)
        cfg = dict(B=len(sample), S=S, d=d, V=V, n_blocks=len(sample) * (S // 16), R=1, max_new=cap)
        rc, want, _, _ = H.run_oracle_engine("paged", cfg, w, s_offs, s_toks, fix=1, threads=min(16, os.cpu_count() or 8))
        assert rc == 0 and len(want) == len(sample)
        got = {k: mine[i] for k, i in enumerate(sample)}
        ties, errors = H.classify_token_mismatches(w, got, want)
        assert not errors, f"(sample index, position, margin) differ from the oracle beyond a numerical tie: {errors[:4]}"
        assert len(ties) <= 1, f"too many tie flips in a sample of {len(sample)}: {ties}"
        print(f"headline job: {generated} tokens, sample of {len(sample)} requests vs oracle: {len(ties)} tie flips")
        # strong scaling shards the SAME request set: rank 3 of 8 (1024 rows, its prompt-length-balanced deal of 1024
        # requests) must produce, for its requests, the token lists of the one-GPU job
        from min_llm_inference_b200.sharding import shard_requests_balanced
        l_offs, l_toks, ids = shard_requests_balanced(offs, toks, 3, 8)
        l_plen = np.diff(l_offs).astype(np.int64)
        l_blocks = int(np.maximum((l_plen + cap + 1 + 15) // 16, 4).sum()) + 64
        ec8 = mli.EngineCfg(B // 8, S, d, V, l_blocks, wl["R"], 0, len(ids), None, cap, 0, 0)
        eng8 = mli.Engine(ctx, ec8, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        eng8.submit(l_offs, l_toks)
        eng8.run()
        shard, _ = eng8.results()
        eng8.close()
        assert len(ids) == n_req // 8
        ties8, errors8 = H.classify_token_mismatches(w, shard, {k: mine[int(i)] for k, i in enumerate(ids)})
        assert not errors8, errors8[:4]
        assert len(ties8) <= len(ids) // 100, f"too many tie flips between the sharded and the one-GPU job: {ties8}"
        print(f"rank 3 of 8 vs the one-GPU job: {len(ties8)} tie flips in {len(ids)} requests")
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
        torch.cuda.empty_cache()


def test_configs2_job_at_full_size(torch_cuda, ctx):
    """BASELINE configs[2] as the bench runs it (tools/run_config.py c3): 1024 rows, emb_dim 2048, n_sequence 4096,
    2048 requests U[64,2048], 256-token cap, 40 GB of KV pages -- continuous batching in which the pool, not the row
    count, limits admission (512 iterations for two waves of requests).  Same three checks as the headline job: a request sample against the CPU oracle
    (each alone in a roomy pool: tokens do not depend on the schedule), per-request properties, idempotence."""
    torch = torch_cuda
    import os
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import run_config
    p = run_config.PRESETS["c3"]
    torch.cuda.empty_cache()
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 60e9:
        pytest.skip(f"needs ~45 GB of free HBM, {free_b / 1e9:.0f} GB free")
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        B, S, d, V, cap, n_req = p["B"], p["S"], p["d"], p["V"], p["max_new"], p["n_req"]
        assert (B, S, d, n_req, p["lo"], p["hi"], cap) == (1024, 4096, 2048, 2048, 64, 2048, 256)
        n_blocks = int(p["pool_gb"] * 1e9 // (16 * 3 * d * 4))
        w = H.make_weights(1001, d, V, S, "Z")
        offs, toks = H.make_prompts(2002, n_req, p["lo"], p["hi"])
        dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
        ec = mli.EngineCfg(B, S, d, V, n_blocks, 1, 0, n_req, None, cap, 0, 0)
        eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        runs = []
        for _ in range(2):
            eng.submit(offs, toks)
            eng.run()
            res, order = eng.results()
            st = eng.stats()
            runs.append((res, order.copy(), (st.steps, st.generated_tokens, st.preemptions, st.n_finished,
                                             st.peak_resident_rows, st.min_free_pages)))
        eng.close()
        (mine, order, stats), (again, order2, stats2) = runs
        assert stats == stats2 and np.array_equal(order, order2)
        assert all(np.array_equal(mine[i], again[i]) for i in range(n_req))
        steps, generated, preempt, n_fin, peak_rows, min_free = stats
        assert n_fin == n_req and peak_rows <= B and min_free >= 0
        total_new = 0
        for i in range(n_req):
            t, pr = mine[i], toks[offs[i]:offs[i + 1]]
            assert np.array_equal(t[:len(pr)], pr), f"request {i}: prompt not kept"
            new = len(t) - len(pr)
            assert 1 <= new <= cap and (new == cap or t[-1] == mli.EOF_TOKEN_ID)
            total_new += new
        # (a pre-empted request generates its tokens once: they travel with it as the new prompt)
        assert total_new == generated
        plen = np.diff(offs)
        rng = np.random.default_rng(6)
        sample = sorted(set([int(np.argmin(plen)), int(np.argmax(plen))] + rng.choice(n_req, 6, replace=False).tolist()))
        s_offs = np.zeros(len(sample) + 1, np.int32)
        parts = []
        for k, i in enumerate(sample):
            parts.append(toks[offs[i]:offs[i + 1]])
            s_offs[k + 1] = s_offs[k] + len(parts[-1])
        s_toks = np.concatenate(parts).astype(np.int32)
        cfg = dict(B=len(sample), S=S, d=d, V=V, n_blocks=len(sample) * (S // 16), R=1, max_new=cap)
        rc, want, _, _ = H.run_oracle_engine("paged", cfg, w, s_offs, s_toks, fix=1, threads=min(16, os.cpu_count() or 8))
        assert rc == 0 and len(want) == len(sample)
        ties, errors = H.classify_token_mismatches(w, {k: mine[i] for k, i in enumerate(sample)}, want)
        assert not errors, f"(sample index, position, margin) differ from the oracle beyond a numerical tie: {errors[:4]}"
        assert len(ties) <= 1, f"too many tie flips in a sample of {len(sample)}: {ties}"
        print(f"configs[2] job: {generated} tokens, {steps} iterations, {preempt} pre-emptions, "
              f"sample of {len(sample)} vs oracle: {len(ties)} tie flips")
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
        torch.cuda.empty_cache()


def test_configs3_dimensions_tensor_core_vs_exact_order(torch_cuda, ctx):
    """BASELINE configs[3] dimensions (emb_dim 4096, n_sequence 32768, prompts U[24000,32511]) with 32 of its 128 rows
    (51 GB of pages instead of 181 GB; the GEMM plan is the same one: CTA pairs, two K passes, grouped item order).
    A CPU oracle run of one such request is ~2 TFLOP of scalar loops, so the witness is the engine's own exact-order
    mode -- whose K, V, q, logits are bit-identical to the reference's naive kernels (tests/test_gpu_stages.py) -- on
    four of the requests: same token lists, numerical ties classified by a float64 replay."""
    torch = torch_cuda
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import run_config
    p = run_config.PRESETS["c4"]
    torch.cuda.empty_cache()
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 75e9:
        pytest.skip(f"needs ~60 GB of free HBM, {free_b / 1e9:.0f} GB free")
    B, S, d, V, cap, n_req = 32, p["S"], p["d"], p["V"], 8, 32
    assert (S, d, p["lo"], p["hi"]) == (32768, 4096, 24000, 32511)
    w = H.make_weights(1001, d, V, S, "Z")
    offs, toks = H.make_prompts(2002, n_req, p["lo"], p["hi"])
    dw = {k: torch.from_numpy(v).cuda() for k, v in w.items()}

    def job(mode, rows, o, t):
        ctx.set_option(mli.OPT_GEMM_MODE, mode)
        n = len(o) - 1
        ec = mli.EngineCfg(rows, S, d, V, rows * (S // 16), 1, 0, n, None, cap, 0, 0)
        eng = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        eng.submit(o, t)
        eng.run()
        res, _ = eng.results()
        st = eng.stats()
        eng.close()
        return res, st

    try:
        try:
            mine, st = job(mli.GEMM_TCGEN05, B, offs, toks)
        except mli.MliError:
            pytest.skip("tcgen05 path not available")
        import ctypes as C
        plan = (C.c_int * 5)()
        ctx._check(ctx.lib.mli_debug_last_gemm_plan(ctx.h, 3, plan))
        assert list(plan)[:2] == [2, 1] and plan[4] == 2, f"expected CTA pairs with two K passes, got {list(plan)}"
        assert st.n_finished == n_req and st.preemptions == 0
        sample = [0, 7, 19, 31]
        s_offs = np.zeros(len(sample) + 1, np.int32)
        parts = []
        for k, i in enumerate(sample):
            parts.append(toks[offs[i]:offs[i + 1]])
            s_offs[k + 1] = s_offs[k] + len(parts[-1])
        want, _ = job(mli.GEMM_SIMT_EXACT, len(sample), s_offs, np.concatenate(parts).astype(np.int32))
        for k, i in enumerate(sample):
            new = len(mine[i]) - (offs[i + 1] - offs[i])
            assert 1 <= new <= cap
        ties, errors = H.classify_token_mismatches(w, {k: mine[i] for k, i in enumerate(sample)}, want)
        assert not errors, f"(sample index, position, margin) differ from the exact-order mode beyond a tie: {errors[:4]}"
        assert len(ties) <= 1, ties
        print(f"configs[3] dims: {st.generated_tokens} tokens after {int(offs[-1])} prompt positions, "
              f"{len(ties)} tie flips against the exact-order mode")
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
        torch.cuda.empty_cache()
