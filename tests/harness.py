"""Shared test/bench harness: library loaders, seeded synthetic inputs, paged fixtures.

Only tests/, __graft_entry__.smoke() and bench.py import this.  It is the one place that loads
oracle/ (the CPU restatement and the compiled reference); the product package never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

ORACLE_SO = REPO / "oracle" / "_build" / "liboracle.so"
REF_SO = REPO / "oracle" / "_ref" / "libmli_ref.so"
PAGE = 16
EOF = 1023

_P, _I = C.c_void_p, C.c_int


# ---------------------------------------------------------------------------------------------
# library loaders
# ---------------------------------------------------------------------------------------------
class OrcCfg(C.Structure):
    _fields_ = [("n_batch", _I), ("n_sequence", _I), ("emb_dim", _I), ("n_vocab", _I),
                ("n_blocks", _I), ("n_forward_rounds", _I), ("fix_stale_lengths", _I),
                ("max_steps", _I), ("max_new_tokens", _I), ("max_prefill_positions", _I),
                ("prefill_chunk_positions", _I)]


class OrcStats(C.Structure):
    _fields_ = [("steps", C.c_longlong), ("generated_tokens", C.c_longlong),
                ("preemptions", C.c_longlong), ("n_finished", _I)]


def build_oracle():
    subprocess.run(["make", "-C", str(REPO / "oracle"), "oracle"], check=True,
                   stdout=subprocess.DEVNULL)


_oracle = None


def load_oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not ORACLE_SO.exists():
            build_oracle()
        lib = C.CDLL(str(ORACLE_SO))
        lib.orc_set_threads.restype = _I
        lib.orc_get_threads.restype = _I
        lib.orc_argmax_device_rule.restype = _I
        lib.orc_argmax_device_rule.argtypes = [_P, _I]
        for name in ("orc_paged_engine_run", "orc_dense_engine_run"):
            fn = getattr(lib, name)
            fn.restype = _I
            fn.argtypes = [C.POINTER(OrcCfg)] + [_P] * 5 + [_I] + [_P] * 5 + [C.POINTER(OrcStats)]
        _oracle = lib
    return _oracle


_ref = None


def ref_available() -> bool:
    return REF_SO.exists()


def load_ref() -> C.CDLL:
    """The unmodified reference compiled into oracle/_ref (needs a GPU to run)."""
    global _ref
    if _ref is None:
        lib = C.CDLL(str(REF_SO))
        lib.ref_last_error.restype = C.c_char_p
        _ref = lib
    return _ref


def p(x):
    """pointer of numpy array / torch tensor / None as c_void_p"""
    if x is None:
        return _P(None)
    if hasattr(x, "data_ptr"):
        return _P(x.data_ptr())
    return _P(x.ctypes.data)


def check_ref(rc):
    if rc != 0:
        raise RuntimeError("reference failed: " + load_ref().ref_last_error().decode())


# ---------------------------------------------------------------------------------------------
# synthetic inputs (fixed seeds; the reference's own fixtures use std::random_device)
#   dist "R": every float tensor i.i.d. U(0,1]  (reference src/kernels/rand_assign.cu:7-15), the EOF
#             row of emb_table scaled by eof_ratio (tests/test_utils.cpp:87-95)
#   dist "Z": zero-mean, scaled so q.k/sqrt(d) = O(1) and the softmax is not one-hot
# ---------------------------------------------------------------------------------------------
def uniform01(rng, shape):
    return (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)  # (0, 1]


def make_weights(seed, d, V, S, dist="R", eof_ratio=1.0001):
    rng = np.random.default_rng(seed)
    if dist == "R":
        w = {k: uniform01(rng, (d, d)) for k in ("wk", "wq", "wv")}
        emb = uniform01(rng, (V, d))
        pos = uniform01(rng, (S, d))
    else:
        sc = np.float32(np.sqrt(12.0 / d))
        w = {k: ((rng.random((d, d), dtype=np.float32) - 0.5) * sc * 1.5).astype(np.float32)
             for k in ("wk", "wq", "wv")}
        emb = ((rng.random((V, d), dtype=np.float32) - 0.5) * 2.0).astype(np.float32)
        pos = ((rng.random((S, d), dtype=np.float32) - 0.5) * 0.5).astype(np.float32)
    if V > EOF:
        emb[EOF] *= np.float32(eof_ratio)
    w["emb"], w["pos"] = emb, pos
    return w


def make_prompts(seed, n_req, lo, hi, V=1024):
    """prompt lengths U[lo,hi], tokens U{0..min(V,EOF)-1} (tests/test_utils.cpp:661-674)"""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n_req)
    offs = np.zeros(n_req + 1, np.int32)
    offs[1:] = np.cumsum(lens)
    toks = rng.integers(0, min(V, EOF), size=int(offs[-1])).astype(np.int32)
    return offs, toks


class PagedCase:
    """A [B] batch with a shuffled page pool, mirroring
    generate_paged_attention_wrapper_device_tensors (tests/test_utils.cpp:695-773): row r owns
    ceil(min(L+1,S)/16) pages drawn from one slab in shuffled order; the slab starts random."""

    def __init__(self, seed, B, S, d, lengths, dist="R", extra_pages=0):
        rng = np.random.default_rng(seed)
        self.B, self.S, self.d, self.W = B, S, d, S // PAGE
        self.lengths = np.asarray(lengths, np.int32).copy()
        need = [(-(-min(int(L) + 1, S) // PAGE) if L > 0 else 0) for L in self.lengths]
        self.n_pages = int(sum(need)) + extra_pages
        self.page_floats = PAGE * 3 * d
        order = rng.permutation(self.n_pages)
        self.page_ids = -np.ones((B, self.W), np.int64)
        k = 0
        for r in range(B):
            for j in range(need[r]):
                self.page_ids[r, j] = order[k]
                k += 1
        if dist == "R":
            self.pool = uniform01(rng, (max(self.n_pages, 1), self.page_floats))
        else:
            self.pool = ((rng.random((max(self.n_pages, 1), self.page_floats), dtype=np.float32)
                          - 0.5) * 2.0).astype(np.float32)

    def table_for(self, base_addr: int) -> np.ndarray:
        """page table of raw pointers for a pool living at base_addr"""
        t = np.zeros((self.B, self.W), np.uint64)
        m = self.page_ids >= 0
        t[m] = np.uint64(base_addr) + self.page_ids[m].astype(np.uint64) * np.uint64(self.page_floats * 4)
        return t

    def host(self):
        """(pool copy, pointer table) on the host for the oracle"""
        pool = self.pool.copy()
        return pool, self.table_for(pool.ctypes.data)

    def device(self, torch, dev="cuda"):
        pool = torch.from_numpy(self.pool.copy()).to(dev)
        tab = torch.from_numpy(self.table_for(pool.data_ptr()).view(np.int64)).to(dev)
        return pool, tab

    def view(self, pool, r, j, off):
        """numpy/torch view of element row (r, j, off) inside a pool laid out like self.pool"""
        pid = int(self.page_ids[r, j // PAGE])
        o = (j % PAGE) * 3 * self.d + off * self.d
        return pool[pid, o:o + self.d]

    def gather(self, pool, off, lengths=None):
        """dense [B,S,d] copy of sub-row `off` for positions < lengths (zeros elsewhere)"""
        lengths = self.lengths if lengths is None else lengths
        pool = pool.cpu().numpy() if hasattr(pool, "cpu") else pool
        out = np.zeros((self.B, self.S, self.d), np.float32)
        for r in range(self.B):
            for j in range(int(lengths[r])):
                out[r, j] = self.view(pool, r, j, off)
        return out


def rel_err(a, b):
    """max |a-b| / max(|b|) over the tensor: the 'rel 1e-4' figure of BASELINE.json north_star"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / denom


def run_oracle_engine(kind, cfg: dict, w, offs, toks, fix=0, max_steps=0, threads=1):
    lib = load_oracle()
    lib.orc_set_threads(threads)
    n_req = len(offs) - 1
    S = cfg["S"]
    c = OrcCfg(cfg["B"], S, cfg["d"], cfg["V"], cfg.get("n_blocks", 0), cfg.get("R", 1), fix, max_steps,
               cfg.get("max_new", 0), cfg.get("max_prefill", 0), cfg.get("chunk", 0))
    ids = np.zeros(n_req, np.int32)
    fo = np.zeros(n_req + 1, np.int32)
    ft = np.zeros(n_req * S, np.int32)
    st = OrcStats()
    fn = lib.orc_paged_engine_run if kind == "paged" else lib.orc_dense_engine_run
    rc = fn(C.byref(c), p(w["emb"]), p(w["pos"]), p(w["wk"]), p(w["wq"]), p(w["wv"]), n_req,
            p(offs), p(toks), p(ids), p(fo), p(ft), C.byref(st))
    res = {int(ids[i]): ft[fo[i]:fo[i + 1]].copy() for i in range(st.n_finished)}
    return rc, res, ids[:st.n_finished].copy(), st


def top2_margin_f64(w, tokens, t):
    """float64 replay of ONE request up to position t (tokens[:t] are given): returns
    (best token, runner-up token, (best - runner-up) / max|logit|).  Used to decide whether a token
    that differs between two fp32 evaluation orders sits on a numerical tie (mismatches are
    classified by the top-2 margin of the logits)."""
    E, P = w["emb"].astype(np.float64), w["pos"].astype(np.float64)
    d = E.shape[1]
    x = E[np.asarray(tokens[:t])] + P[:t]
    K, V = x @ w["wk"].astype(np.float64), x @ w["wv"].astype(np.float64)
    q = x[-1] @ w["wq"].astype(np.float64)
    s = K @ q / np.sqrt(np.float64(d))
    p = np.exp(s - s.max())
    p /= p.sum()
    logits = (p @ V) @ E.T
    order = np.argsort(-logits)
    scale = float(np.max(np.abs(logits)))
    return int(order[0]), int(order[1]), float(logits[order[0]] - logits[order[1]]) / max(scale, 1e-300)


def classify_token_mismatches(w, mine, want, tie_rel=2e-5):
    """requests whose token lists differ -> list of (request, position, margin).  A mismatch is a
    numerical TIE when, at the first differing position, the two tokens are the float64 top-2 and
    their logits differ by less than tie_rel of the largest |logit| (the 3xTF32 / fp32 re-association
    noise level); anything else is a real error (margin = None)."""
    ties, errors = [], []
    for i in sorted(want):
        a, b = np.asarray(mine[i]), np.asarray(want[i])
        if a.shape == b.shape and np.array_equal(a, b):
            continue
        n = min(len(a), len(b))
        diff = np.flatnonzero(a[:n] != b[:n])
        t = int(diff[0]) if len(diff) else n
        if t >= n:
            errors.append((i, t, None))
            continue
        best, second, margin = top2_margin_f64(w, b, t)
        if {int(a[t]), int(b[t])} == {best, second} and margin < tie_rel:
            ties.append((i, t, margin))
        else:
            errors.append((i, t, margin))
    return ties, errors
