"""N > 1 path on CPU: world_size-2 gloo run of the request sharding + token gather.  Each rank runs
the CPU oracle engine on its shard (standing in for the GPU engine, which needs a device); the
gathered token table must equal a single-process run over all requests, because tokens are
schedule-independent once lengths are handled correctly."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, str(Path(__file__).resolve().parent))
import harness as H  # noqa: E402
from min_llm_inference_b200.sharding import gather_tokens, shard_requests  # noqa: E402

CFG = dict(B=4, S=64, d=32, V=1024, n_blocks=24, R=1)
N_REQ = 13   # deliberately not divisible by the world size


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = H.make_weights(31, CFG["d"], CFG["V"], CFG["S"], "Z")
    offs, toks = H.make_prompts(33, N_REQ, 1, 40)
    lo, lt, ids = shard_requests(offs, toks, rank, world)
    rc, res, order, st = H.run_oracle_engine("paged", CFG, w, lo, lt, fix=1)
    assert rc == 0 and st.n_finished == len(ids)
    table = torch.zeros((len(ids), CFG["S"]), dtype=torch.int32)
    counts = torch.zeros((len(ids),), dtype=torch.int32)
    for k in range(len(ids)):
        counts[k] = len(res[k])
        table[k, :len(res[k])] = torch.from_numpy(res[k])
    all_t, all_c = gather_tokens(table, counts, N_REQ)
    if rank == 0:
        np.save(os.path.join(out_dir, "tokens.npy"), all_t.numpy())
        np.save(os.path.join(out_dir, "counts.npy"), all_c.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    tokens, counts = np.load(tmp_path / "tokens.npy"), np.load(tmp_path / "counts.npy")
    w = H.make_weights(31, CFG["d"], CFG["V"], CFG["S"], "Z")
    offs, toks = H.make_prompts(33, N_REQ, 1, 40)
    rc, res, order, st = H.run_oracle_engine("paged", CFG, w, offs, toks, fix=1)
    assert rc == 0 and st.n_finished == N_REQ
    for i in range(N_REQ):
        assert counts[i] == len(res[i])
        assert np.array_equal(tokens[i, :counts[i]], res[i]), f"request {i} differs after sharding"


def test_shard_requests_partition():
    offs, toks = H.make_prompts(1, 10, 1, 9)
    seen = []
    for world in (1, 2, 3, 4, 8):
        seen.clear()
        for r in range(world):
            lo, lt, ids = shard_requests(offs, toks, r, world)
            assert lo[0] == 0 and lo[-1] == len(lt)
            for k, g in enumerate(ids):
                assert np.array_equal(lt[lo[k]:lo[k + 1]], toks[offs[g]:offs[g + 1]])
            seen.extend(ids.tolist())
        assert sorted(seen) == list(range(10))


def test_balanced_sharding_is_a_partition():
    """shard_requests_balanced: every request lands on exactly one rank, the ranks' prompt-token totals agree to
    within a percent, and the local (offsets, tokens) reproduce the global prompts"""
    from min_llm_inference_b200.sharding import shard_requests_balanced
    offs, toks = H.make_prompts(7, 1001, 4, 500)
    for world in (1, 2, 3, 8):
        seen, totals = [], []
        for rank in range(world):
            lo, lt, ids = shard_requests_balanced(offs, toks, rank, world)
            assert len(lo) - 1 == len(ids)
            for k, g in enumerate(ids):
                assert np.array_equal(lt[lo[k]:lo[k + 1]], toks[offs[g]:offs[g + 1]])
            seen.extend(ids.tolist())
            totals.append(int(lo[-1]))
        assert sorted(seen) == list(range(1001))
        assert max(totals) <= 1.01 * (sum(totals) / world) + 500
