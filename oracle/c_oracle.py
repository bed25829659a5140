"""ctypes loader for oracle/radar_oracle.c -- the bit-exact checker for the canonical fp32 arithmetic.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libradar_oracle.so")
_lib = None

MODE_DPR, MODE_KL, MODE_HYBRID = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "radar_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.radar_oracle_search.restype = C.c_int
        _lib.radar_oracle_score_pairs.restype = C.c_int
        _lib.radar_oracle_merge_topk.restype = C.c_int
        _lib.radar_oracle_max_threads.restype = C.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def max_threads() -> int:
    return int(lib().radar_oracle_max_threads())


def prepare_corpus(probs: np.ndarray, eps: float = 1e-8, normalize: bool = False) -> np.ndarray:
    p = _f32(probs)
    out = np.empty((p.shape[0], 16), dtype=np.float32)
    lib().radar_oracle_prepare_corpus(_p(p), C.c_int64(p.shape[0]), C.c_int(p.shape[1]), C.c_float(eps),
                                      C.c_int(int(normalize)), _p(out))
    return out


def prepare_queries(probs: np.ndarray, mask: Optional[np.ndarray] = None, eps: float = 1e-8,
                    normalize: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    p = _f32(probs)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    p16 = np.empty((p.shape[0], 16), dtype=np.float32)
    h = np.empty((p.shape[0],), dtype=np.float32)
    lib().radar_oracle_prepare_queries(_p(p), _p(m), C.c_int64(p.shape[0]), C.c_int(p.shape[1]),
                                       C.c_float(eps), C.c_int(int(normalize)), _p(p16), _p(h))
    return p16, h


def search(mode: int, k: int, q_emb=None, p16=None, entropy=None, c_emb=None, logq16=None,
           alpha: float = 0.5, idx_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    q_emb, p16, entropy, c_emb, logq16 = map(_f32, (q_emb, p16, entropy, c_emb, logq16))
    nq = (q_emb if q_emb is not None else p16).shape[0]
    n = (c_emb if c_emb is not None else logq16).shape[0]
    d = 0 if c_emb is None else c_emb.shape[1]
    k = min(k, n)
    out_s = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    rc = lib().radar_oracle_search(C.c_int(mode), _p(q_emb), _p(p16), _p(entropy), _p(c_emb), _p(logq16),
                                   C.c_int64(nq), C.c_int64(n), C.c_int(d), C.c_int(k), C.c_float(alpha),
                                   C.c_int64(idx_offset), _p(out_s), _p(out_i))
    if rc != 0:
        raise RuntimeError(f"radar_oracle_search failed rc={rc}")
    return out_s, out_i


def score_pairs(mode: int, ids: np.ndarray, q_emb=None, p16=None, entropy=None, c_emb=None, logq16=None,
                alpha: float = 0.5) -> np.ndarray:
    q_emb, p16, entropy, c_emb, logq16 = map(_f32, (q_emb, p16, entropy, c_emb, logq16))
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    nq, k = ids.shape
    d = 0 if c_emb is None else c_emb.shape[1]
    out = np.empty((nq, k), dtype=np.float32)
    lib().radar_oracle_score_pairs(C.c_int(mode), _p(q_emb), _p(p16), _p(entropy), _p(c_emb), _p(logq16),
                                   C.c_int64(nq), C.c_int(d), C.c_int(k), C.c_float(alpha), _p(ids), _p(out))
    return out


def merge_topk(scores: np.ndarray, idx: np.ndarray, k_out: int, ascending: bool
               ) -> Tuple[np.ndarray, np.ndarray]:
    """scores/idx: [parts, Q, k_in]."""
    s = _f32(scores)
    i = np.ascontiguousarray(idx, dtype=np.int64)
    parts, nq, k_in = s.shape
    out_s = np.empty((nq, k_out), dtype=np.float32)
    out_i = np.empty((nq, k_out), dtype=np.int64)
    rc = lib().radar_oracle_merge_topk(_p(s), _p(i), C.c_int64(nq), C.c_int(parts), C.c_int(k_in),
                                       C.c_int(k_out), C.c_int(int(ascending)), _p(out_s), _p(out_i))
    if rc != 0:
        raise RuntimeError(f"radar_oracle_merge_topk failed rc={rc}")
    return out_s, out_i


# ---- packed exchange words (include/radar_retrieval.h: radar_search out_packed / radar_merge_packed) --------------
# numpy restatement used by the CPU tests of the sharded host logic and as the checker of the CUDA merge kernel:
#   word = (orderable bits of the ranking key << 32) | (0xFFFFFFFF - global id);  key = score (DPR, hybrid) or
#   0 - score (KL, where smaller is better);  0 = padding;  larger word = better rank, smaller id wins ties.
def _f2ord(key: np.ndarray) -> np.ndarray:
    u = (np.asarray(key, dtype=np.float32) + np.float32(0.0)).view(np.uint32)  # -0 -> +0
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def _ord2f(o: np.ndarray) -> np.ndarray:
    o = np.asarray(o, dtype=np.uint32)
    u = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    return u.view(np.float32)


def pack_results(mode: int, scores: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """(scores float32[Q,k], global ids int64[Q,k], id < 0 = padding) -> int64[Q,k] holding the uint64 words."""
    s = _f32(scores)
    key = (np.float32(0.0) - s).astype(np.float32) if mode == MODE_KL else s
    i = np.asarray(idx, dtype=np.int64)
    word = (_f2ord(key).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.where(i < 0, 0, i).astype(np.uint64))
    return np.where(i < 0, np.uint64(0), word).view(np.int64)


def merge_packed(packed: np.ndarray, k_out: int, mode: int) -> Tuple[np.ndarray, np.ndarray]:
    """packed: int64[parts, Q, k_in] words -> (scores float32[Q,k_out], ids int64[Q,k_out]) best first."""
    w = np.ascontiguousarray(packed).view(np.uint64)
    parts, nq, k_in = w.shape
    allw = np.transpose(w, (1, 0, 2)).reshape(nq, parts * k_in)
    top = np.sort(allw, axis=1)[:, ::-1][:, :k_out]
    key = _ord2f((top >> np.uint64(32)).astype(np.uint32))
    ids = (np.uint64(0xFFFFFFFF) - (top & np.uint64(0xFFFFFFFF))).astype(np.int64)
    scores = (np.float32(0.0) - key).astype(np.float32) if mode == MODE_KL else key.astype(np.float32)
    empty = top == 0
    scores = np.where(empty, np.float32(np.inf) if mode == MODE_KL else np.float32(-np.inf), scores).astype(np.float32)
    ids = np.where(empty, -1, ids)
    return scores, ids
