"""Config surface of the retrieval path.

``RetrievalConfig`` keeps the reference's fields and defaults verbatim
(annotate_retrieve/modeling_dense_passage_retrieval.py:183-188); ``IterativeRAGConfig`` mirrors
annotate_retrieve/modeling_iterative_rag.py:12-20.  The reference's ``configs/knowledge.yaml`` is a
0-byte file, so ``KnowledgeConfig`` defines its content: the dataclass defaults plus the keys the
B200 path adds ("Config / flags").
"""
from __future__ import annotations

import os
from dataclasses import dataclass, fields
from typing import List, Optional

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_KNOWLEDGE_YAML = os.path.join(_REPO, "configs", "knowledge.yaml")


@dataclass
class RetrievalConfig:
    embedding_dim: int = 512
    num_retrieved: int = 5
    hybrid_alpha: float = 0.5
    device: str = 'cuda'


@dataclass
class IterativeRAGConfig:
    num_iterations: int = 3
    max_new_tokens: int = 100
    top_k: int = 5
    temperature: float = 0.7
    consistency_threshold: float = 0.7
    observation_vocab: Optional[List[str]] = None
    device: str = 'cuda'


@dataclass
class KnowledgeConfig:
    """Content of configs/knowledge.yaml."""
    # reference dataclass defaults (must stay as they are)
    embedding_dim: int = 512
    num_retrieved: int = 5
    hybrid_alpha: float = 0.5
    device: str = 'cuda'
    # keys added by the B200 path
    num_observations: int = 14
    kl_eps: float = 1e-8
    normalize: bool = False         # divide probability rows by their sum before the clamp (categorical KL)
    precision: str = 'bf16'         # 'bf16' (tensor-core filter + fp32 re-score) | 'fp32' (canonical-exact)
    score_mode: str = 'hybrid'      # 'dpr' | 'kl' | 'hybrid'
    algo: str = 'auto'              # 'auto' | 'simt' | 'tc'
    overfetch: int = 0              # candidates re-scored per query on the filter path; 0 = automatic
    rag_top_k: int = 5              # IterativeRAGConfig.top_k
    rag_num_iterations: int = 3     # IterativeRAGConfig.num_iterations

    def retrieval_config(self) -> RetrievalConfig:
        return RetrievalConfig(self.embedding_dim, self.num_retrieved, self.hybrid_alpha, self.device)

    def validate(self) -> "KnowledgeConfig":
        if self.score_mode not in ("dpr", "kl", "hybrid"):
            raise ValueError(f"score_mode must be dpr|kl|hybrid, got {self.score_mode!r}")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be bf16|fp32, got {self.precision!r}")
        if self.algo not in ("auto", "simt", "tc"):
            raise ValueError(f"algo must be auto|simt|tc, got {self.algo!r}")
        if not (0.0 <= self.hybrid_alpha <= 1.0):
            raise ValueError("hybrid_alpha must be in [0, 1]")
        if self.num_observations != 14:
            raise ValueError("num_observations is fixed at 14 (CheXpert-14)")
        if not (0.0 < self.kl_eps < 1.0):
            raise ValueError("kl_eps must be in (0, 1)")
        return self


def load_knowledge_config(path: Optional[str] = None) -> KnowledgeConfig:
    """Read configs/knowledge.yaml; unknown keys raise, missing keys keep their defaults."""
    import yaml
    path = path or DEFAULT_KNOWLEDGE_YAML
    with open(path, "r") as fh:
        raw = yaml.safe_load(fh) or {}
    if not isinstance(raw, dict):
        raise ValueError(f"{path}: expected a mapping at top level")
    known = {f.name for f in fields(KnowledgeConfig)}
    unknown = set(raw) - known
    if unknown:
        raise ValueError(f"{path}: unknown keys {sorted(unknown)}")
    return KnowledgeConfig(**raw).validate()
