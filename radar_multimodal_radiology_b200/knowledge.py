"""Observation-based KL-divergence retrieval -- the content of the reference's empty ``src/knowledge``.

``src/knowledge/__init__.py`` is a 0-byte file in the reference and ``README.md:64`` only mentions
"KL-divergence retrieval" in prose, so this module *defines* the package (SURVEY.md section 8a rows
K1-K3, definitions frozen in section 8c):

    KL(p_query || q_case) = sum_{j<14} p_j (log p_j - log q_j),  p, q clamped to [eps, 1], eps = 1e-8
    14 independent sigmoid probabilities in CheXpert-14 order (train_expert_models.py:50-65), not
    renormalised unless ``normalize=True``; cases ranked by ascending KL; an observation mask m zeroes
    the masked terms (p <- m*p).

The arithmetic runs in the sm_100a kernels behind ``RadarIndex``; nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch

# canonical observation order (train_expert_models.py:50-65; modeling_expert_model_gnn.py:135-149)
OBSERVATION_NAMES: List[str] = [
    "Enlarged Cardiomediastinum", "Cardiomegaly", "Lung Opacity", "Lung Lesion", "Edema",
    "Consolidation", "Pneumonia", "Atelectasis", "Pneumothorax", "Pleural Effusion",
    "Pleural Other", "Fracture", "Support Devices", "No Finding",
]
NUM_OBSERVATIONS = len(OBSERVATION_NAMES)
_NAME_TO_BIT = {n.lower(): i for i, n in enumerate(OBSERVATION_NAMES)}


def observation_bits(names: Iterable[str]) -> int:
    """14-bit set of the named observations in CheXpert-14 order.  Names outside the vocabulary (the
    iterative-RAG detector's default list has "Pulmonary Edema" and "Rib Fracture",
    modeling_iterative_rag.py:30-36) map to no bit."""
    bits = 0
    for n in names:
        j = _NAME_TO_BIT.get(str(n).lower())
        if j is not None:
            bits |= 1 << j
    return bits


def bits_to_mask(bits: Sequence[int], device=None) -> torch.Tensor:
    """uint8[Q,14] observation mask from 14-bit sets; an empty set keeps every observation (a query
    with nothing missing is not restricted)."""
    b = torch.as_tensor(np.asarray(bits, dtype=np.int64))
    m = ((b[:, None] >> torch.arange(NUM_OBSERVATIONS)[None, :]) & 1).to(torch.uint8)
    m[b == 0] = 1
    return m if device is None else m.to(device)


def missing_observation_mask(missing_sets: Sequence[Iterable[str]], device=None) -> torch.Tensor:
    """Observation masks for a batch of re-retrieval queries: the KL is restricted to the observations a
    case is still missing (the batched counterpart of the text query "Cases with <obs>, <obs>" built at
    modeling_iterative_rag.py:115-125)."""
    return bits_to_mask([observation_bits(s) for s in missing_sets], device)


class ObservationKLRetriever:
    """KL-divergence case retrieval over a corpus of 14-observation probability vectors."""

    def __init__(self, device="cuda", eps: float = 1e-8, normalize: bool = False, precision: str = "fp32",
                 algo: str = "auto", idx_offset: int = 0):
        from .index import RadarIndex
        self.index = RadarIndex(d=4, device=device, precision=precision, eps=eps, normalize=normalize,
                                algo=algo, idx_offset=idx_offset)

    @property
    def ntotal(self) -> int:
        return self.index.ntotal

    def build(self, corpus_probs) -> "ObservationKLRetriever":
        self.index.reset()
        self.index.add_observations(corpus_probs)
        return self

    def search(self, query_probs, k: int = 10, mask=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(KL float32[Q,k] ascending, ids int64[Q,k])."""
        k = min(k, self.ntotal)
        return self.index.search(None, k, query_probs=query_probs, mask=mask, mode="kl")

    def retrieve(self, query_probs, k: int = 10, mask=None) -> Tuple[List[int], List[float]]:
        """Single-query convenience: (ids, KL values) as Python lists, nearest first."""
        kl, ids = self.search(torch.as_tensor(query_probs).reshape(1, -1), k,
                              None if mask is None else torch.as_tensor(mask).reshape(1, -1))
        return ids[0].tolist(), [float(v) for v in kl[0].tolist()]
