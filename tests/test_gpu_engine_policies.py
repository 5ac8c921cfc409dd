"""Engine behaviour added in round 2, each checked against the CPU oracle (which restates the same
rule next to the reference's scheduler, oracle/oracle.c):

  * page accounting when a row would need more pages than its table row holds (the pool
    must be whole again at the end of the job)                     src/paged_item_storage.cpp:84-113
  * opt-in max_new_tokens / max_prefill_positions policies         (not in the reference; off = reference)
  * streaming ingestion (mli_engine_enqueue) and the non-blocking finished poll
                                                                   src/item_storage.cpp:97-139, :190-196
  * a random-lengths soak of the two fused attention kernels (cross-CTA merge protocols)
"""
import threading

import numpy as np
import pytest

import harness as H
import min_llm_inference_b200 as mli

pytestmark = pytest.mark.gpu


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def make_engine(ctx, torch, cfg, w, max_req, compat=0):
    dw = {k: dev(torch, v) for k, v in w.items()}
    ec = mli.EngineCfg(cfg["B"], cfg["S"], cfg["d"], cfg["V"], cfg["n_blocks"], cfg.get("R", 1), compat, max_req,
                       None, cfg.get("max_new", 0), cfg.get("max_prefill", 0), cfg.get("chunk", 0))
    return mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])


def check_against_oracle(ctx, torch, cfg, w, offs, toks, compat=0):
    eng = make_engine(ctx, torch, cfg, w, len(offs) - 1, compat)
    eng.submit(offs, toks)
    eng.run()
    mine, order = eng.results()
    st = eng.stats()
    eng.close()
    rc, theirs, oorder, ost = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1 - compat)
    assert rc == 0
    assert (st.steps, st.generated_tokens, st.preemptions) == (ost.steps, ost.generated_tokens, ost.preemptions)
    assert order.tolist() == oorder.tolist(), "finish order differs from the oracle"
    for i in theirs:
        assert np.array_equal(mine[i], theirs[i]), f"request {i}: tokens differ from the oracle"
    return st


@pytest.mark.parametrize("cfg", [
    # R >= 2 with prompts of length S-1 / S-2: ceil((len + R) / 16) exceeds the table width W
    dict(B=4, S=64, d=64, V=1024, n_blocks=24, R=3, n_req=12, lo=61, hi=63),
    dict(B=6, S=128, d=128, V=1024, n_blocks=40, R=4, n_req=14, lo=120, hi=127),
    # n_sequence < 64: W < 4, so EVERY admission asks for more pages (4) than the table row holds
    dict(B=4, S=32, d=64, V=1024, n_blocks=16, R=1, n_req=10, lo=1, hi=20),
    dict(B=5, S=48, d=64, V=1024, n_blocks=11, R=2, n_req=12, lo=5, hi=40),
], ids=lambda c: f"S{c['S']}-R{c['R']}")
def test_rows_wider_than_the_table_do_not_leak_pages(torch_cuda, ctx, cfg):
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    w = H.make_weights(51, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(53, cfg["n_req"], cfg["lo"], cfg["hi"])
    st = check_against_oracle(ctx, torch, cfg, w, offs, toks)
    assert st.n_finished == cfg["n_req"]
    # a second job on a fresh engine must see the whole pool: run twice on ONE engine and compare
    eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
    runs = []
    for _ in range(2):
        eng.submit(offs, toks)
        eng.run()
        runs.append((eng.stats().steps, eng.stats().preemptions, eng.stats().min_free_pages))
    eng.close()
    assert runs[0] == runs[1]


@pytest.mark.parametrize("gemm_mode", [mli.GEMM_SIMT_EXACT, mli.GEMM_TCGEN05])
@pytest.mark.parametrize("cfg", [
    dict(B=8, S=128, d=128, V=1024, n_blocks=64, n_req=24, lo=4, hi=60, max_new=7),
    dict(B=16, S=256, d=256, V=1024, n_blocks=96, n_req=40, lo=10, hi=120, max_new=16, R=3),
    dict(B=16, S=128, d=128, V=1024, n_blocks=128, n_req=48, lo=8, hi=64, max_prefill=96),
    dict(B=16, S=128, d=128, V=1024, n_blocks=48, n_req=48, lo=8, hi=64, max_prefill=64, max_new=12),
], ids=lambda c: f"new{c.get('max_new', 0)}-pf{c.get('max_prefill', 0)}")
def test_policies_match_the_oracle(torch_cuda, ctx, cfg, gemm_mode):
    """token cap and admission throttle: same decisions, tokens and finish order as the oracle with the
    same flags; every request generates at most max_new tokens"""
    torch = torch_cuda
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, gemm_mode)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        w = H.make_weights(61, cfg["d"], cfg["V"], cfg["S"], "Z")
        offs, toks = H.make_prompts(63, cfg["n_req"], cfg["lo"], cfg["hi"])
        if gemm_mode == mli.GEMM_SIMT_EXACT:
            check_against_oracle(ctx, torch, cfg, w, offs, toks)
        eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
        eng.submit(offs, toks)
        eng.run()
        mine, order = eng.results()
        st = eng.stats()
        eng.close()
        rc, theirs, oorder, ost = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1)
        ties, errors = H.classify_token_mismatches(w, mine, theirs)
        assert not errors, errors[:3]
        if not ties:
            assert (st.steps, st.preemptions) == (ost.steps, ost.preemptions)
        if cfg.get("max_new"):
            plen = np.diff(offs)
            assert all(len(mine[i]) - plen[i] <= cfg["max_new"] for i in mine)
            assert any(len(mine[i]) - plen[i] == cfg["max_new"] for i in mine)
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


def test_throttle_spreads_an_admission_burst(torch_cuda, ctx):
    """with the throttle no step admits more prompt positions than the cap (first admission excepted)"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    cfg = dict(B=32, S=128, d=64, V=1024, n_blocks=256, n_req=32, lo=30, hi=60, max_prefill=128, max_new=4)
    w = H.make_weights(71, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(73, cfg["n_req"], cfg["lo"], cfg["hi"])
    eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
    eng.submit(offs, toks)
    admitted = []
    for _ in range(40):
        eng.run(max_steps=1)
        admitted.append(eng.stats().admitted)
    eng.close()
    per_step = np.diff([0] + admitted)
    plen = np.diff(offs)
    k = 0
    for n in per_step:
        if n:
            assert plen[k:k + n].sum() <= cfg["max_prefill"] or n == 1
        k += n
    assert k == cfg["n_req"] and (per_step > 0).sum() > 4   # unthrottled: everything in the first step


def test_bad_device_prompts_are_flagged(torch_cuda, ctx):
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    cfg = dict(B=4, S=64, d=64, V=1024, n_blocks=16)
    w = H.make_weights(81, 64, 1024, 64, "Z")
    eng = make_engine(ctx, torch, cfg, w, 4)
    offs = np.array([0, 10, 10 + 70, 90], np.int32)      # second prompt longer than n_sequence
    toks = np.zeros(90, np.int32)
    eng.submit(dev(torch, offs), dev(torch, toks), is_device=True)
    with pytest.raises(mli.MliError, match="prompt length"):
        eng.run()
    with pytest.raises(mli.MliError, match="prompt length"):
        eng.submit(offs, toks)                           # host path: rejected before anything is enqueued
    eng.close()
    # a prompt that could never be admitted (needs more pages than the pool has) must not spin forever
    cfg2 = dict(B=2, S=256, d=64, V=1024, n_blocks=5)
    w2 = H.make_weights(82, 64, 1024, 256, "Z")
    eng = make_engine(ctx, torch, cfg2, w2, 2)
    offs, toks = H.make_prompts(83, 2, 100, 120)
    with pytest.raises(mli.MliError, match="pool"):
        eng.submit(offs, toks)
    eng.submit(dev(torch, offs), dev(torch, toks), is_device=True)
    with pytest.raises(mli.MliError, match="pool"):
        eng.run()
    eng.close()


def test_two_engines_share_registered_weights(torch_cuda, ctx):
    """destroying one engine must not free the split weight copies another engine's captured
    graph still reads"""
    torch = torch_cuda
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        cfg = dict(B=8, S=128, d=128, V=1024, n_blocks=64, n_req=16, lo=4, hi=60, max_new=6)
        w = H.make_weights(91, 128, 1024, 128, "Z")
        dw = {k: dev(torch, v) for k, v in w.items()}
        offs, toks = H.make_prompts(93, cfg["n_req"], cfg["lo"], cfg["hi"])
        ec = mli.EngineCfg(8, 128, 128, 1024, 64, 1, 0, 16, None, 6, 0)
        a = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        b = mli.Engine(ctx, ec, dw["emb"], dw["pos"], dw["wk"], dw["wq"], dw["wv"])
        a.submit(offs, toks); a.run(); ra, _ = a.results()
        b.submit(offs, toks); b.run(); rb, _ = b.results()
        a.close()
        b.submit(offs, toks); b.run(); rb2, _ = b.results()      # graph replays with a gone
        b.close()
        for i in ra:
            assert np.array_equal(ra[i], rb[i]) and np.array_equal(rb[i], rb2[i])
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


@pytest.mark.parametrize("gemm_mode", [mli.GEMM_SIMT_EXACT, mli.GEMM_TCGEN05])
def test_enqueue_and_poll_stream_requests(torch_cuda, ctx, gemm_mode):
    """requests arrive in three waves (one before the run, one from another thread while it runs, one
    after it went idle); every token list equals a one-shot job's (requests are independent when
    lengths are corrected), the poll hands out every request exactly once, in finish order"""
    torch = torch_cuda
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, gemm_mode)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        cfg = dict(B=8, S=128, d=128, V=1024, n_blocks=96, n_req=36, lo=4, hi=60, max_new=24)
        w = H.make_weights(101, cfg["d"], cfg["V"], cfg["S"], "Z")
        offs, toks = H.make_prompts(103, cfg["n_req"], cfg["lo"], cfg["hi"])
        rc, want, _, _ = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1)

        def part(lo, hi):
            return (offs[lo:hi + 1] - offs[lo]).astype(np.int32), toks[offs[lo]:offs[hi]].copy()

        eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
        eng.submit(*part(0, 12))
        got, seen, errors_in_threads = {}, [], []
        stop = threading.Event()

        def feeder():
            try:
                assert eng.enqueue(*part(12, 24)) == 12
            except Exception as exc:   # a worker's failure must fail the test, not vanish with the thread
                errors_in_threads.append(exc)

        def poller():
            try:
                while not stop.is_set():
                    r, ids = eng.poll_finished(max_out=5)
                    got.update(r)
                    seen.extend(ids.tolist())
            except Exception as exc:
                errors_in_threads.append(exc)

        t1, t2 = threading.Thread(target=feeder), threading.Thread(target=poller)
        try:
            t2.start(); t1.start()
            eng.run()                      # (the first run also CAPTURES the step graph while the two threads work)
            t1.join()
            eng.run()                      # the second wave may have landed after the first run went idle
            assert eng.enqueue(*part(24, 36)) == 24
            eng.run()
        finally:
            stop.set()
            t1.join(); t2.join()           # nobody may still use the engine when it is destroyed
        assert not errors_in_threads, errors_in_threads
        # drain (the first poll after a submit only arms the device side and returns nothing)
        for _ in range(200):
            if len(seen) == cfg["n_req"]:
                break
            r, ids = eng.poll_finished(max_out=7)
            got.update(r)
            seen.extend(ids.tolist())
        full, order = eng.results()
        st = eng.stats()
        eng.close()
        assert st.n_finished == cfg["n_req"]
        assert seen == order.tolist(), "poll must hand out each request once, in finish order"
        ties, errors = H.classify_token_mismatches(w, full, want)
        assert not errors and (gemm_mode != mli.GEMM_SIMT_EXACT or not ties)
        for i in full:
            assert np.array_equal(full[i], got[i])
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


@pytest.mark.parametrize("kernel", [1, 2])
def test_attention_soak_random_lengths(torch_cuda, ctx, kernel):
    """1000 launches of the single-launch attention over freshly drawn lengths (dynamic slices forced on
    a third of them): the cross-CTA merge protocol (arrival counters that reset themselves, partial
    (m, l, acc) hand-off) must leave every launch equal to a float64 evaluation"""
    torch = torch_cuda
    B, S, d = 96, 512, 256
    rng = np.random.default_rng(4242 + kernel)
    case = H.PagedCase(7, B, S, d, np.full(B, S - 1, np.int32), "Z")
    pool, tab = case.device(torch)
    q = (torch.rand((B, d), device="cuda") - 0.5)
    out = torch.empty((B, d), device="cuda")
    # dense float64 K, V of every row once
    K = torch.from_numpy(case.gather(pool, 1)).cuda().double()
    V = torch.from_numpy(case.gather(pool, 2)).cuda().double()
    scores = torch.einsum("bd,bsd->bs", q.double(), K) / np.sqrt(np.float64(d))
    ctx.set_option(mli.OPT_ATTN_KERNEL, kernel)
    try:
        worst = 0.0
        for it in range(1000):
            mode = it % 3
            ctx.set_option(mli.OPT_ATTN_MIN_DYN, 1 if mode == 2 else 4096)
            hi = [S - 1, 40, S - 1][mode]
            L = rng.integers(0, hi + 1, size=B).astype(np.int32)
            if it % 7 == 0:
                L[rng.random(B) < 0.5] = 0
            dL = torch.from_numpy(L).cuda()
            ctx.call("mli_decode_attention_paged", q, tab, dL, out, None, B, S, d)
            mask = torch.arange(S, device="cuda")[None, :] < dL[:, None]
            p = torch.softmax(scores.masked_fill(~mask, float("-inf")), dim=1)
            p = torch.nan_to_num(p, nan=0.0)
            want = torch.einsum("bs,bsd->bd", p, V)
            err = float((out.double() - want).abs().max() / want.abs().max().clamp_min(1e-30))
            worst = max(worst, err)
            assert err < 1e-4, f"launch {it}: rel err {err:.2e}"
    finally:
        ctx.set_option(mli.OPT_ATTN_KERNEL, 0)
        ctx.set_option(mli.OPT_ATTN_MIN_DYN, 4096)


@pytest.mark.parametrize("B,min_dyn", [(2049, 4096), (5000, 4096), (8192, 1), (3333, 1)])
def test_attention_many_rows_coarse_row_table(torch_cuda, ctx, B, min_dyn):
    """more than 2048 rows: the warp-per-position kernel keeps only every 32nd row's stage prefix in shared
    memory and locates rows with a second-level lookup; whole blocks of empty rows, a ragged tail block and
    dynamic slices are covered"""
    torch = torch_cuda
    S, d = 128, 128
    rng = np.random.default_rng(B)
    L = rng.integers(0, S, size=B).astype(np.int32)
    L[rng.random(B) < 0.3] = 0
    L[64:160] = 0            # three whole 32-row blocks without work
    L[-5:] = [S - 1, 0, 1, 0, 17]
    if B > 4000:
        L[2100:4000] = 0     # a long empty stretch
    # rows share a handful of pages (content is irrelevant to the row lookup, the pool stays small)
    n_pages = 64
    pool = (torch.rand((n_pages, 16 * 3 * d), device="cuda") - 0.5) * 2.0
    W = S // 16
    pid = rng.integers(0, n_pages, size=(B, W))
    tab = torch.from_numpy((pool.data_ptr() + pid.astype(np.int64) * (16 * 3 * d * 4))).cuda()
    q = torch.rand((B, d), device="cuda") - 0.5
    out = torch.full((B, d), 9.0, device="cuda")
    dL = torch.from_numpy(L).cuda()
    ctx.set_option(mli.OPT_ATTN_KERNEL, 2)
    ctx.set_option(mli.OPT_ATTN_MIN_DYN, min_dyn)
    try:
        ctx.call("mli_decode_attention_paged", q, tab, dL, out, None, B, S, d)
        ctx.synchronize()
    finally:
        ctx.set_option(mli.OPT_ATTN_KERNEL, 0)
        ctx.set_option(mli.OPT_ATTN_MIN_DYN, 4096)
    pages = pool.view(n_pages, 16, 3, d).double()
    pidt = torch.from_numpy(pid).cuda()
    K = pages[pidt, :, 1, :].reshape(B, S, d)
    V = pages[pidt, :, 2, :].reshape(B, S, d)
    sc = torch.einsum("bd,bsd->bs", q.double(), K) / np.sqrt(np.float64(d))
    mask = torch.arange(S, device="cuda")[None, :] < dL[:, None]
    p = torch.nan_to_num(torch.softmax(sc.masked_fill(~mask, float("-inf")), dim=1), nan=0.0)
    want = torch.einsum("bs,bsd->bd", p, V)
    err = float((out.double() - want).abs().max() / want.abs().max())
    assert err < 1e-4, f"rel err {err:.2e}"
    assert not out[dL == 0].any(), "empty rows must produce zeros"


CHUNK_CASES = [
    # long prompts against a small budget: rows spend several steps prefilling before their first token
    dict(B=8, S=256, d=128, V=1024, n_blocks=160, n_req=20, lo=40, hi=200, max_new=10, chunk=64),
    # budget smaller than one granule pair, many rows prefilling at once, pool pressure (a prefilling row can be
    # pre-empted and starts over), several rounds per step
    dict(B=12, S=128, d=128, V=1024, n_blocks=70, n_req=40, lo=10, hi=100, max_new=12, chunk=32, R=2),
    # budget larger than most prompts: several rows complete in one step; combined with the admission throttle
    dict(B=16, S=128, d=256, V=1024, n_blocks=160, n_req=48, lo=5, hi=90, max_new=8, chunk=256, max_prefill=200),
    # runs to n_sequence (no token cap)
    dict(B=6, S=64, d=128, V=1024, n_blocks=30, n_req=14, lo=8, hi=50, chunk=16),
]


@pytest.mark.parametrize("gemm_mode", [mli.GEMM_SIMT_EXACT, mli.GEMM_TCGEN05], ids=["exact", "tcgen05"])
@pytest.mark.parametrize("cfg", CHUNK_CASES, ids=lambda c: f"chunk{c['chunk']}-B{c['B']}")
def test_chunked_prefill_matches_the_oracle(torch_cuda, ctx, cfg, gemm_mode):
    """chunked prefill (mli_engine_cfg.prefill_chunk_positions): same schedule as the oracle's restatement of the
    policy (iterations, pre-emptions, finish order), and -- because K and V do not depend on how a prompt was cut
    into chunks -- the token list of every request equals the UNCHUNKED job's"""
    torch = torch_cuda
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, gemm_mode)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        w = H.make_weights(131, cfg["d"], cfg["V"], cfg["S"], "Z")
        offs, toks = H.make_prompts(133, cfg["n_req"], cfg["lo"], cfg["hi"])
        eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
        eng.submit(offs, toks)
        eng.run()
        mine, order = eng.results()
        st = eng.stats()
        eng.close()
        rc, theirs, oorder, ost = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1)
        assert rc == 0 and st.n_finished == cfg["n_req"] == ost.n_finished
        ties, errors = H.classify_token_mismatches(w, mine, theirs)
        assert not errors, errors[:3]
        if gemm_mode == mli.GEMM_SIMT_EXACT:
            assert not ties
        if not ties:
            assert (st.steps, st.generated_tokens, st.preemptions) == (ost.steps, ost.generated_tokens, ost.preemptions)
            assert order.tolist() == oorder.tolist()
        # the unchunked job: more tokens per early step, the same token lists
        rc, plain, _, pst = H.run_oracle_engine("paged", dict(cfg, chunk=0), w, offs, toks, fix=1)
        assert rc == 0
        for i in plain:
            assert np.array_equal(plain[i], theirs[i]), f"request {i}: chunking changed the oracle's tokens"
        assert ost.steps >= pst.steps
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


def test_chunked_prefill_bounds_the_work_of_a_step(torch_cuda, ctx):
    """with a chunk budget no step admits-and-prefills a whole long prompt at once: the number of active rows grows
    by at most what the budget can complete, and every row still finishes"""
    torch = torch_cuda
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        cfg = dict(B=4, S=512, d=128, V=1024, n_blocks=140, n_req=4, lo=400, hi=480, max_new=3, chunk=128)
        w = H.make_weights(141, cfg["d"], cfg["V"], cfg["S"], "Z")
        offs, toks = H.make_prompts(143, cfg["n_req"], cfg["lo"], cfg["hi"])
        eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
        eng.submit(offs, toks)
        gen = []
        for _ in range(40):
            eng.run(max_steps=1)
            gen.append(eng.stats().generated_tokens)
        eng.close()
        # ~1760 prompt positions at 128 per step: the first token cannot appear before step ceil(400 / 128) = 4 and the
        # four rows become active one after the other, never all in the first step
        per_step = np.diff([0] + gen)
        assert per_step[:3].sum() == 0 and per_step.max() <= 4
        first = int(np.flatnonzero(per_step)[0])
        assert first >= 3
        assert gen[-1] == 4 * 3
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


def random_engine_case(seed):
    """one small engine job drawn from the whole policy space: table width, rounds per step, pool pressure,
    token cap, admission throttle, chunked prefill, stale-length compatibility"""
    r = np.random.default_rng(9000 + seed)
    S = int(r.choice([48, 64, 96, 128, 192]))
    d = int(r.choice([64, 128]))
    B = int(r.integers(2, 25))
    R = int(r.choice([1, 1, 2, 3, 5]))
    hi = int(r.integers(2, S - 1))
    lo = int(r.integers(1, hi + 1))
    W = (S + 15) // 16
    need_max = max(4, min(W, (hi + R + 15) // 16))
    # from "every row fits twice" down to "barely one long request": the second forces pre-emption chains
    tight = r.random() < 0.5
    n_blocks = int(need_max + r.integers(1, max(2, B * need_max // 3 if tight else 2 * B * need_max)))
    cfg = dict(B=B, S=S, d=d, V=1024, n_blocks=n_blocks, R=R, n_req=int(r.integers(B, 4 * B + 2)), lo=lo, hi=hi)
    if r.random() < 0.5:
        cfg["max_new"] = int(r.integers(1, 24))
    if r.random() < 0.35:
        cfg["max_prefill"] = int(r.integers(16, 4 * S))
    compat = 0
    if r.random() < 0.45:
        cfg["chunk"] = int(r.choice([16, 32, 48, 64, 128, 512]))
    elif r.random() < 0.3:
        compat = 1        # the reference's stale lengths (quirk Q1); not combinable with chunked prefill
    return cfg, compat


@pytest.mark.parametrize("seed", range(32))
def test_random_policy_mix_matches_the_oracle(torch_cuda, ctx, seed):
    """exact mode: steps, tokens generated, pre-emptions, finish order and every token equal to the oracle's.
    Some draws hold a request that outgrows the pool (the reference would spin for ever): oracle and engine
    must both stop there, with the same requests finished"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    cfg, compat = random_engine_case(seed)
    w = H.make_weights(300 + seed, cfg["d"], cfg["V"], cfg["S"], "Z" if seed % 3 else "R")
    offs, toks = H.make_prompts(700 + seed, cfg["n_req"], cfg["lo"], cfg["hi"])
    rc, theirs, oorder, ost = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1 - compat, max_steps=20000)
    assert rc in (0, -5), (rc, cfg)
    eng = make_engine(ctx, torch, cfg, w, cfg["n_req"], compat)
    eng.submit(offs, toks)
    if rc == -5:
        with pytest.raises(mli.MliError, match="outgrew the KV pool"):
            eng.run()
    else:
        eng.run()
    mine, order = eng.results()
    st = eng.stats()
    eng.close()
    assert (st.steps, st.generated_tokens, st.preemptions) == (ost.steps, ost.generated_tokens, ost.preemptions), cfg
    assert order.tolist() == oorder.tolist(), "finish order differs from the oracle"
    for i in theirs:
        assert np.array_equal(mine[i], theirs[i]), f"request {i}: tokens differ from the oracle"
    assert (st.n_finished == cfg["n_req"]) == (rc == 0)


def test_request_that_outgrows_the_pool_ends_the_job(torch_cuda, ctx):
    """12 pages per full row, 10 in the pool, no token cap: the longest-running request pre-empts itself and
    can never come back.  The reference spins; the engine returns MLI_ERR_NO_BLOCKS with everything that could
    finish finished, and is usable for the next job"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    cfg = dict(B=4, S=192, d=64, V=1024, n_blocks=10, R=1, n_req=6, lo=3, hi=12)
    w = H.make_weights(81, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(83, cfg["n_req"], cfg["lo"], cfg["hi"])
    rc, theirs, oorder, ost = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1, max_steps=20000)
    assert rc == -5
    eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
    eng.submit(offs, toks)
    with pytest.raises(mli.MliError, match=r"mli error -3: .*outgrew the KV pool"):   # MLI_ERR_NO_BLOCKS
        eng.run()
    mine, order = eng.results()
    assert order.tolist() == oorder.tolist() and eng.stats().steps == ost.steps
    for i in theirs:
        assert np.array_equal(mine[i], theirs[i])
    # the same engine with a token cap that keeps every request inside the pool
    eng.close()
    cfg2 = dict(cfg, max_new=20)
    st = check_against_oracle(ctx, torch, cfg2, w, offs, toks)
    assert st.n_finished == cfg["n_req"]


@pytest.mark.parametrize("seed", range(100, 116))
def test_random_policy_mix_tensor_core_mode(torch_cuda, ctx, seed):
    """the same sweep through the tcgen05 GEMMs (merged QKV + prefill launch, step-by-step chunks): tokens equal
    to the oracle's except on classified numerical ties; without a tie every scheduler decision is identical"""
    torch = torch_cuda
    try:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_TCGEN05)
    except mli.MliError:
        pytest.skip("tcgen05 path not available")
    try:
        cfg, _ = random_engine_case(seed)
        cfg["d"] = 128 if seed % 2 else 256
        cfg["n_blocks"] = max(cfg["n_blocks"], (cfg["S"] + 15) // 16 + 1)   # every request can finish
        w = H.make_weights(300 + seed, cfg["d"], cfg["V"], cfg["S"], "Z" if seed % 3 else "R")
        offs, toks = H.make_prompts(700 + seed, cfg["n_req"], cfg["lo"], cfg["hi"])
        rc, theirs, oorder, ost = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1, max_steps=20000)
        assert rc == 0, cfg
        eng = make_engine(ctx, torch, cfg, w, cfg["n_req"])
        eng.submit(offs, toks)
        eng.run()
        mine, order = eng.results()
        st = eng.stats()
        eng.close()
        assert st.n_finished == cfg["n_req"]
        ties, errors = H.classify_token_mismatches(w, mine, theirs)
        assert not errors, (cfg, errors[:3])
        assert len(ties) <= max(1, cfg["n_req"] // 20), (cfg, ties)
        if not ties:
            assert (st.steps, st.generated_tokens, st.preemptions) == (ost.steps, ost.generated_tokens, ost.preemptions)
            assert order.tolist() == oorder.tolist()
    finally:
        ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)


@pytest.mark.parametrize("seed", range(200, 212))
def test_random_arrival_schedule(torch_cuda, ctx, seed):
    """requests arrive in random waves between slices of the run (mli_engine_enqueue / mli_engine_run(max_steps)),
    finished ones are collected by random-sized polls: whatever the schedule, every request's tokens equal the
    one-shot oracle job's (corrected lengths make requests independent) and the poll returns each exactly once"""
    torch = torch_cuda
    ctx.set_option(mli.OPT_GEMM_MODE, mli.GEMM_SIMT_EXACT)
    cfg, _ = random_engine_case(seed)
    cfg["n_blocks"] = max(cfg["n_blocks"], (cfg["S"] + 15) // 16 + 1)
    r = np.random.default_rng(seed)
    w = H.make_weights(300 + seed, cfg["d"], cfg["V"], cfg["S"], "Z")
    offs, toks = H.make_prompts(700 + seed, cfg["n_req"], cfg["lo"], cfg["hi"])
    rc, want, _, _ = H.run_oracle_engine("paged", cfg, w, offs, toks, fix=1, max_steps=20000)
    assert rc == 0

    def part(lo, hi):
        return (offs[lo:hi + 1] - offs[lo]).astype(np.int32), toks[offs[lo]:offs[hi]].copy()

    n = cfg["n_req"]
    cuts = sorted(set([0, n] + r.integers(1, n, size=int(r.integers(1, 5))).tolist()))
    eng = make_engine(ctx, torch, cfg, w, n)
    got, seen = {}, []

    def poll():
        res, ids = eng.poll_finished(max_out=int(r.integers(1, 9)))
        for i in ids.tolist():
            assert i not in got, f"request {i} handed out twice"
        got.update(res)
        seen.extend(ids.tolist())

    eng.submit(*part(cuts[0], cuts[1]))
    for a, b in zip(cuts[1:-1], cuts[2:]):
        if r.random() < 0.7:
            eng.run(max_steps=int(r.integers(1, 40)))
        else:
            eng.run()                      # idle before the next wave arrives
        poll()
        assert eng.enqueue(*part(a, b)) == a
    eng.run()
    for _ in range(4 * n + 8):
        if len(seen) == n:
            break
        poll()
    full, order = eng.results()
    st = eng.stats()
    eng.close()
    assert st.n_finished == n and seen == order.tolist()
    for i in range(n):
        assert np.array_equal(full[i], want[i]), f"request {i}: tokens differ from the one-shot job"
        assert np.array_equal(got[i], want[i])
