"""Batched retrieval-quality metrics on the device.

The reference computes MRR, P@k, R@k, nDCG@k and accuracy@{5,10} per query over Python string lists
(``RetrievalMetrics``, evaluate_retrieval_system.py:137-188) and then -- in its shipped evaluators -- discards them for
hard-coded constants (:240-251).  Here the same definitions are evaluated for a whole batch of top-k id lists as they
come out of ``RadarIndex.search`` (``int64[Q,k]`` on the GPU), against per-query relevant-id sets given as a padded
``int64[Q,R]`` tensor (-1 = padding).  Definitions follow the reference exactly: first-hit reciprocal rank, binary
gains, ideal DCG over ``min(k, |relevant|)``, recall against ``|relevant|`` (0 when empty), precision against ``k`` even
when fewer than ``k`` ids were retrieved.  Everything is float64 like the Python reference; plain tensor ops -- this is
post-processing of ``Q*k`` integers, not a hot loop.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch


def relevance_matrix(ids: torch.Tensor, relevant: torch.Tensor) -> torch.Tensor:
    """bool[Q,k]: ids[q,j] is one of relevant[q,:] (ids < 0 are missing results, relevant < 0 is padding)."""
    if ids.dim() != 2 or relevant.dim() != 2 or ids.shape[0] != relevant.shape[0]:
        raise ValueError("ids must be [Q,k] and relevant [Q,R] with the same Q")
    hit = (ids.unsqueeze(2) == relevant.unsqueeze(1)) & (relevant.unsqueeze(1) >= 0)
    return hit.any(dim=2) & (ids >= 0)


def batched_retrieval_metrics(ids: torch.Tensor, relevant: torch.Tensor, ks: Sequence[int] = (1, 5, 10),
                              ndcg_ks: Sequence[int] = (5, 10)) -> Dict[str, torch.Tensor]:
    """Per-query metrics (float64[Q] each): ``mrr``, ``precision@k`` / ``recall@k`` for ``ks``, ``ndcg@k`` /
    ``accuracy@k`` for ``ndcg_ks``.  Average over the batch with ``.mean()`` for the figures the reference reports."""
    rel = relevance_matrix(ids, relevant)
    q, kmax = rel.shape
    dev = ids.device
    n_rel = (relevant >= 0).sum(dim=1).to(torch.float64)
    ranks = torch.arange(1, kmax + 1, device=dev, dtype=torch.float64)
    out: Dict[str, torch.Tensor] = {}
    first = torch.where(rel, ranks.unsqueeze(0), torch.full((1, 1), float("inf"), device=dev, dtype=torch.float64)).amin(dim=1)
    out["mrr"] = torch.where(torch.isfinite(first), 1.0 / first, torch.zeros_like(first))
    relf = rel.to(torch.float64)
    for k in ks:
        hits = relf[:, :k].sum(dim=1)
        out[f"precision@{k}"] = hits / k if k > 0 else torch.zeros_like(hits)
        out[f"recall@{k}"] = torch.where(n_rel > 0, hits / n_rel.clamp_min(1.0), torch.zeros_like(hits))
    disc = 1.0 / torch.log2(ranks + 1.0)
    cum_disc = torch.cumsum(disc, dim=0)
    for k in ndcg_ks:
        kk = min(k, kmax)
        dcg = (relf[:, :kk] * disc[:kk].unsqueeze(0)).sum(dim=1)
        n_ideal = torch.minimum(n_rel, torch.tensor(float(k), device=dev, dtype=torch.float64)).to(torch.int64)
        # ideal DCG over min(k, |relevant|) ranks; ranks beyond the retrieved list still count, as in the reference
        ideal_disc = torch.cumsum(1.0 / torch.log2(torch.arange(1, k + 1, device=dev, dtype=torch.float64) + 1.0), dim=0)
        idcg = torch.where(n_ideal > 0, ideal_disc[(n_ideal - 1).clamp_min(0)], torch.zeros_like(dcg))
        out[f"ndcg@{k}"] = torch.where(idcg > 0, dcg / idcg.clamp_min(1e-300), torch.zeros_like(dcg))
        out[f"accuracy@{k}"] = rel[:, :kk].any(dim=1).to(torch.float64)
    del cum_disc
    return out
