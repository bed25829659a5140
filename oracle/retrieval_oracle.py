"""CPU oracle for RADAR's case-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product package (``radar_multimodal_radiology_b200``) never does: it fails
loudly when the CUDA extension is missing.

What is restated here, and from where (paths relative to /root/reference):

* Dense-passage scoring + top-k -- ``faiss.IndexFlatIP`` as it is used at
  ``annotate_retrieve/modeling_dense_passage_retrieval.py:297-300`` (``IndexFlatIP(d)``, ``add``,
  ``ntotal``) and ``:312-314`` (``search(float32[nq,d], k) -> (D float32[nq,k] descending,
  I int64[nq,k])``).  faiss itself is a third-party dependency that is NOT vendored under
  /root/reference and is not version pinned (no requirements/lock file exists in the tree), so the
  published algorithm is restated: exact brute-force fp32 inner product of every query row with
  every stored row, k largest per query sorted descending, ids = insertion order.
* KL-divergence observation retrieval, hybrid fusion, observation masks -- the reference ships NO
  code for these (``src/knowledge/__init__.py`` and ``configs/knowledge.yaml`` are 0-byte files;
  ``hybrid_alpha`` at ``modeling_dense_passage_retrieval.py:187`` is never read).  The definitions
  below are the frozen specification from SURVEY.md section 8c:
      KL[i,n]  = sum_{j<14} p[i,j] * (log p[i,j] - log q[n,j]),  p, q clamped to [eps, 1], eps=1e-8,
                 14 independent sigmoid probabilities in CheXpert-14 order
                 (``train_expert_models.py:50-65``), NOT renormalised; ranked ascending.
      mask     : p <- m * p before both terms (masked-out terms contribute exactly 0).
      hybrid   : s[i,n] = alpha * <e_q[i], e_c[n]> - (1 - alpha) * KL[i,n]; ranked descending.
* Re-rank used by the iterative-RAG rounds -- ``TargetedRetriever.rank_retrieved_passages``
  ``annotate_retrieve/modeling_iterative_rag.py:127-152`` and the substring detector ``:38-49``.
* Retrieval-quality metrics -- ``RetrievalMetrics`` ``evaluate_retrieval_system.py:137-188``.

PARITY STATUS.  The reference holds no golden vector, known-answer test or fixture for this path
(all four files under ``tests/`` are 0-byte; ``test_2.py`` asserts nothing).  The wrapper-level
behaviour (k clamping, ordering, list types, hard-negative split, re-rank scores, RAG call
pattern) IS pinned against outputs of the reference itself, generated in the authoring container
by ``tests/golden/make_reference_fixtures.py`` (which imports /root/reference with a numpy stand-in
for the absent ``faiss`` module).  The arithmetic of KL / hybrid has no reference implementation
at all: for those pieces this oracle is the specification -> **parity unpinned** for K1-K3.

Tie rule (the reference leaves it unspecified): better score first, then smaller id.

Two numeric flavours are provided:
  * ``*_fp64``  : float64 direct-form evaluation -- the ground truth for tolerance tests.
  * ``*_blas32``: float32 BLAS evaluation (``Q @ C.T`` then top-k) -- the faithful restatement of
                  what faiss does on a CPU; this is also the timed CPU baseline in ``bench.py``.
The bit-exact checker for the CUDA path's canonical fp32 arithmetic is the C file
``oracle/radar_oracle.c`` (loaded through ``oracle/c_oracle.py``).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

NUM_OBS = 14          # CheXpert-14, train_expert_models.py:50-65
OBS_PAD = 16          # K=16 padded contraction (north_star)
DEFAULT_EPS = 1e-8
MODE_DPR, MODE_KL, MODE_HYBRID = 0, 1, 2

OBSERVATION_NAMES = [  # canonical order, train_expert_models.py:50-65
    "Enlarged Cardiomediastinum", "Cardiomegaly", "Lung Opacity", "Lung Lesion", "Edema",
    "Consolidation", "Pneumonia", "Atelectasis", "Pneumothorax", "Pleural Effusion",
    "Pleural Other", "Fracture", "Support Devices", "No Finding",
]

# the iterative-RAG detector's own default vocabulary, modeling_iterative_rag.py:30-36
RAG_DEFAULT_VOCAB = [
    "Atelectasis", "Cardiomegaly", "Consolidation", "Edema", "Pleural Effusion", "Pneumonia",
    "Pneumothorax", "No Finding", "Fracture", "Support Devices", "Enlarged Cardiomediastinum",
    "Lung Opacity", "Pulmonary Edema", "Rib Fracture",
]


# --------------------------------------------------------------------------------------------
# canonical fp32 preparation (shared definition with the CUDA path and radar_oracle.c)
# --------------------------------------------------------------------------------------------
def _normalize_rows32(p: np.ndarray) -> np.ndarray:
    """Row / (sum of row), float32, sum accumulated left to right in float32."""
    p = np.ascontiguousarray(p, dtype=np.float32)
    s = np.zeros(p.shape[0], dtype=np.float32)
    for j in range(p.shape[1]):
        s = (s + p[:, j]).astype(np.float32)
    return (p / s[:, None]).astype(np.float32)


def prepare_corpus_logq(probs: np.ndarray, eps: float = DEFAULT_EPS,
                        normalize: bool = False) -> np.ndarray:
    """float32[N,16]: log of the clamped corpus probabilities, columns 14,15 = 0.

    log is evaluated in float64 and rounded once to float32 so that CPU and GPU agree bit for bit.
    """
    p = np.ascontiguousarray(probs, dtype=np.float32)
    assert p.ndim == 2 and p.shape[1] == NUM_OBS
    if normalize:
        p = _normalize_rows32(p)
    eps32 = np.float32(eps)
    pc = np.minimum(np.maximum(p, eps32), np.float32(1.0))
    out = np.zeros((p.shape[0], OBS_PAD), dtype=np.float32)
    out[:, :NUM_OBS] = np.log(pc.astype(np.float64)).astype(np.float32)
    return out


def prepare_queries(probs: np.ndarray, mask: Optional[np.ndarray] = None,
                    eps: float = DEFAULT_EPS, normalize: bool = False
                    ) -> Tuple[np.ndarray, np.ndarray]:
    """(p16 float32[Q,16], entropy float32[Q]).

    p16[:, j] = mask_j ? clamp(p_j, eps, 1) : 0 ; entropy = sum_j p16_j * log p16_j accumulated as a
    float32 fma chain j = 0..13 (masked terms add exactly 0).
    """
    p = np.ascontiguousarray(probs, dtype=np.float32)
    assert p.ndim == 2 and p.shape[1] == NUM_OBS
    if normalize:
        p = _normalize_rows32(p)
    eps32 = np.float32(eps)
    pc = np.minimum(np.maximum(p, eps32), np.float32(1.0))
    lp = np.log(pc.astype(np.float64)).astype(np.float32)
    if mask is not None:
        m = np.ascontiguousarray(mask).astype(bool)
        assert m.shape == p.shape
        pc = np.where(m, pc, np.float32(0.0)).astype(np.float32)
        lp = np.where(m, lp, np.float32(0.0)).astype(np.float32)
    p16 = np.zeros((p.shape[0], OBS_PAD), dtype=np.float32)
    p16[:, :NUM_OBS] = pc
    # float32 fma chain, emulated exactly in float64 (a float32*float32 product is exact in
    # float64 and the sum of it with a float32 rounds to float32 the same way fmaf does, except
    # for double-rounding corner cases that the C oracle -- the actual bit-exact checker -- avoids).
    h = np.zeros(p.shape[0], dtype=np.float32)
    for j in range(NUM_OBS):
        h = (pc[:, j].astype(np.float64) * lp[:, j].astype(np.float64)
             + h.astype(np.float64)).astype(np.float32)
    return p16, h


# --------------------------------------------------------------------------------------------
# float64 ground truth (direct form)
# --------------------------------------------------------------------------------------------
def kl_matrix_fp64(q_probs: np.ndarray, c_probs: np.ndarray, mask: Optional[np.ndarray] = None,
                   eps: float = DEFAULT_EPS, normalize: bool = False) -> np.ndarray:
    """KL[i,n] = sum_j m_ij p_ij (log p_ij - log q_nj) in float64, direct (non-decomposed) form."""
    p = np.asarray(q_probs, dtype=np.float32)
    q = np.asarray(c_probs, dtype=np.float32)
    if normalize:
        p, q = _normalize_rows32(p), _normalize_rows32(q)
    eps32 = np.float32(eps)
    p = np.minimum(np.maximum(p, eps32), np.float32(1)).astype(np.float64)
    q = np.minimum(np.maximum(q, eps32), np.float32(1)).astype(np.float64)
    lp, lq = np.log(p), np.log(q)
    if mask is not None:
        p = p * np.asarray(mask).astype(np.float64)
    # sum_j p_ij*lp_ij  -  sum_j p_ij*lq_nj is algebraically the same thing; evaluate the direct
    # form blockwise to keep the cancellation-free property of the definition.
    out = np.empty((p.shape[0], q.shape[0]), dtype=np.float64)
    blk = max(1, (1 << 22) // max(1, q.shape[0]))
    for s in range(0, p.shape[0], blk):
        pb, lpb = p[s:s + blk], lp[s:s + blk]
        out[s:s + blk] = np.einsum("ij,inj->in", pb, lpb[:, None, :] - lq[None, :, :])
    return out


def kl_operand_scale_fp64(q_probs, c_probs, mask=None, eps=DEFAULT_EPS, normalize=False):
    """|sum p log p| + |sum p log q| per (query, case): the scale the 1e-5 KL tolerance is relative to
    ('hard part 3')."""
    p = np.asarray(q_probs, dtype=np.float32)
    q = np.asarray(c_probs, dtype=np.float32)
    if normalize:
        p, q = _normalize_rows32(p), _normalize_rows32(q)
    eps32 = np.float32(eps)
    p = np.minimum(np.maximum(p, eps32), np.float32(1)).astype(np.float64)
    q = np.minimum(np.maximum(q, eps32), np.float32(1)).astype(np.float64)
    lp, lq = np.log(p), np.log(q)
    if mask is not None:
        p = p * np.asarray(mask).astype(np.float64)
    h = np.abs((p * lp).sum(-1))
    x = np.abs(p @ lq.T)
    return h[:, None] + x


def ip_matrix_fp64(q_emb: np.ndarray, c_emb: np.ndarray) -> np.ndarray:
    return np.asarray(q_emb, dtype=np.float64) @ np.asarray(c_emb, dtype=np.float64).T


def score_matrix_fp64(mode: int, q_emb=None, c_emb=None, q_probs=None, c_probs=None, mask=None,
                      alpha: float = 0.5, eps: float = DEFAULT_EPS, normalize: bool = False
                      ) -> np.ndarray:
    """Returned in the API's own sign: DPR -> inner product, KL -> KL value, hybrid -> fused score."""
    if mode == MODE_DPR:
        return ip_matrix_fp64(q_emb, c_emb)
    if mode == MODE_KL:
        return kl_matrix_fp64(q_probs, c_probs, mask, eps, normalize)
    if mode == MODE_HYBRID:
        a = float(np.float32(alpha))
        oma = float(np.float32(1.0) - np.float32(alpha))
        return a * ip_matrix_fp64(q_emb, c_emb) - oma * kl_matrix_fp64(q_probs, c_probs, mask, eps,
                                                                       normalize)
    raise ValueError(f"unknown mode {mode}")


def topk_rows(scores: np.ndarray, k: int, descending: bool, idx_offset: int = 0
              ) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k per row with the oracle's tie rule (better score first, then smaller id)."""
    s = np.asarray(scores)
    q, n = s.shape
    k = min(k, n)
    key = -s if descending else s
    ids = np.arange(n, dtype=np.int64)
    out_s = np.empty((q, k), dtype=s.dtype)
    out_i = np.empty((q, k), dtype=np.int64)
    for r in range(q):
        order = np.lexsort((ids, key[r]))[:k]
        out_s[r] = s[r, order]
        out_i[r] = order + idx_offset
    return out_s, out_i


def search_fp64(mode: int, k: int, **kw) -> Tuple[np.ndarray, np.ndarray]:
    idx_offset = kw.pop("idx_offset", 0)
    s = score_matrix_fp64(mode, **kw)
    return topk_rows(s, k, descending=(mode != MODE_KL), idx_offset=idx_offset)


# --------------------------------------------------------------------------------------------
# float32 BLAS flavour: the faiss.IndexFlatIP restatement and the timed CPU baseline
# --------------------------------------------------------------------------------------------
def _topk_blas_block(s: np.ndarray, k: int, descending: bool) -> Tuple[np.ndarray, np.ndarray]:
    n = s.shape[1]
    k = min(k, n)
    key = -s if descending else s
    if k < n:
        part = np.argpartition(key, k - 1, axis=1)[:, :k]
    else:
        part = np.broadcast_to(np.arange(n), s.shape).copy()
    pk = np.take_along_axis(key, part, axis=1)
    # sort the k survivors by (key, id)
    order = np.lexsort((part, pk), axis=1)
    idx = np.take_along_axis(part, order, axis=1).astype(np.int64)
    return np.take_along_axis(s, idx, axis=1), idx


def search_blas32(mode: int, k: int, q_emb=None, c_emb=None, q_p16=None, q_entropy=None,
                  c_logq16=None, alpha: float = 0.5, chunk_scores: int = 1 << 27
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """Batched float32 CPU search: chunked ``Q_chunk @ C.T`` (+ ``P @ logQ.T``) then top-k.

    Inputs are the *prepared* tensors (p16/entropy/logq16 from ``prepare_*``) so that this leg and
    the CUDA leg consume identical bits.  ``chunk_scores`` bounds the score block (<= 512 MB fp32).
    """
    nq = (q_emb if q_emb is not None else q_p16).shape[0]
    n = (c_emb if c_emb is not None else c_logq16).shape[0]
    k = min(k, n)
    out_s = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    rows = max(1, chunk_scores // max(1, n))
    a32 = np.float32(alpha)
    oma = np.float32(1.0) - a32
    for s0 in range(0, nq, rows):
        sl = slice(s0, min(nq, s0 + rows))
        if mode == MODE_DPR:
            s = q_emb[sl] @ c_emb.T
        else:
            kl = q_entropy[sl, None] - q_p16[sl] @ c_logq16.T
            if mode == MODE_KL:
                s = kl
            else:
                s = a32 * (q_emb[sl] @ c_emb.T) - oma * kl
        out_s[sl], out_i[sl] = _topk_blas_block(s.astype(np.float32, copy=False), k,
                                                descending=(mode != MODE_KL))
    return out_s, out_i


class IndexFlatIP:
    """numpy restatement of the four faiss members the reference touches
    (modeling_dense_passage_retrieval.py:297-300, :313): ``IndexFlatIP(d)``, ``add``, ``ntotal``,
    ``search``.  Truthiness follows faiss's SWIG objects (always truthy), which the reference
    relies on at ``:310``."""

    def __init__(self, d: int):
        self.d = int(d)
        self._x = np.zeros((0, self.d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return int(self._x.shape[0])

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._x = np.vstack([self._x, x])

    def search(self, x: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        n = self.ntotal
        kk = min(k, n)
        d = np.full((x.shape[0], k), -np.finfo(np.float32).max, dtype=np.float32)
        i = np.full((x.shape[0], k), -1, dtype=np.int64)   # faiss pads missing results with -1
        if kk:
            d[:, :kk], i[:, :kk] = search_blas32(MODE_DPR, kk, q_emb=x, c_emb=self._x)
        return d, i


# --------------------------------------------------------------------------------------------
# shard merge: top-k of the union of per-shard top-k lists
# --------------------------------------------------------------------------------------------
def merge_topk(parts: Sequence[Tuple[np.ndarray, np.ndarray]], k: int, descending: bool
               ) -> Tuple[np.ndarray, np.ndarray]:
    s = np.concatenate([p[0] for p in parts], axis=1)
    i = np.concatenate([p[1] for p in parts], axis=1).astype(np.int64)
    key = -s if descending else s
    order = np.lexsort((i, key), axis=1)[:, :min(k, s.shape[1])]
    return np.take_along_axis(s, order, axis=1), np.take_along_axis(i, order, axis=1)


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """rank g holds rows [g*ceil(N/G), min(N,(g+1)*ceil(N/G)))  (SURVEY.md section 8e)."""
    per = -(-n // world)
    return [(min(n, g * per), min(n, (g + 1) * per)) for g in range(world)]


# --------------------------------------------------------------------------------------------
# iterative-RAG helpers (modeling_iterative_rag.py:38-49, :115-125, :127-152)
# --------------------------------------------------------------------------------------------
def detect_observations(text: str, vocab: Sequence[str] = RAG_DEFAULT_VOCAB) -> Set[str]:
    if not text:
        return set()
    low = text.lower()
    return {o for o in vocab if o.lower() in low}


def rerank_scores(passages: Sequence[str], missing: Set[str],
                  vocab: Sequence[str] = RAG_DEFAULT_VOCAB) -> List[Tuple[str, float]]:
    """overlap/(m+1e-8) + 0.2*min(overlap/max(m,1),1); stable sort descending; 0.5 when either
    argument is empty (modeling_iterative_rag.py:129-130, :139-148)."""
    if not passages or not missing:
        return [(p, 0.5) for p in passages]
    m = len(missing)
    ranked = []
    for p in passages:
        overlap = len(detect_observations(p, vocab) & set(missing))
        ranked.append((p, overlap / (m + 1e-8) + min(overlap / max(m, 1), 1.0) * 0.2))
    ranked.sort(key=lambda t: t[1], reverse=True)
    return ranked


def observation_bits(names: Iterable[str], vocab: Sequence[str] = OBSERVATION_NAMES) -> int:
    """14-bit mask in CheXpert-14 order; names absent from the vocabulary map to no bit
    (observation-vocabulary caveat)."""
    low = {v.lower(): i for i, v in enumerate(vocab)}
    bits = 0
    for nme in names:
        j = low.get(nme.lower())
        if j is not None:
            bits |= 1 << j
    return bits


def rerank_scores_bits(case_bits: np.ndarray, missing_bits: np.ndarray) -> np.ndarray:
    """Bitmask form of ``rerank_scores``: case_bits uint16[Q,k], missing_bits uint16[Q] ->
    float64[Q,k]; rows with no missing bit get 0.5 everywhere."""
    cb = np.asarray(case_bits).astype(np.uint32)
    mb = np.asarray(missing_bits).astype(np.uint32)[:, None]
    pop = lambda x: np.array([[bin(int(v)).count("1") for v in row] for row in x], dtype=np.float64)
    overlap = pop(cb & mb)
    m = pop(mb)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = overlap / (m + 1e-8) + 0.2 * np.minimum(overlap / np.maximum(m, 1.0), 1.0)
    return np.where(m > 0, s, 0.5)


# --------------------------------------------------------------------------------------------
# retrieval-quality metrics (evaluate_retrieval_system.py:137-188)
# --------------------------------------------------------------------------------------------
def retrieval_metrics(retrieved: Sequence[int], relevant: Sequence[int]) -> dict:
    """MRR, P@{1,5,10}, R@{1,5,10}, nDCG@{5,10}, acc@{5,10} for one query, as RetrievalMetrics does
    (first-hit reciprocal rank; binary gains; ideal DCG over min(k, |relevant|))."""
    rel = set(int(r) for r in relevant)
    out = {}
    rr = 0.0
    for rank, idx in enumerate(retrieved, 1):
        if int(idx) in rel:
            rr = 1.0 / rank
            break
    out["mrr"] = rr
    for k in (1, 5, 10):
        top = [int(i) for i in retrieved[:k]]
        hits = sum(1 for i in top if i in rel)
        out[f"precision@{k}"] = hits / k if k else 0.0
        out[f"recall@{k}"] = hits / len(rel) if rel else 0.0
    for k in (5, 10):
        top = [int(i) for i in retrieved[:k]]
        dcg = sum(1.0 / math.log2(r + 1) for r, i in enumerate(top, 1) if i in rel)
        idcg = sum(1.0 / math.log2(r + 1) for r in range(1, min(k, len(rel)) + 1))
        out[f"ndcg@{k}"] = dcg / idcg if idcg > 0 else 0.0
        out[f"accuracy@{k}"] = 1.0 if any(i in rel for i in top) else 0.0
    return out
