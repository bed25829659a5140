"""Drop-in for ``annotate_retrieve/modeling_dense_passage_retrieval.py`` (reference lines 183-355).

Same public names, argument order, defaults and return types:
``RetrievalConfig``, ``CrossModalEmbedder``, ``HybridRetriever`` (``passages``, ``semantic_index``,
``build_indices``, ``retrieve``, ``retrieve_with_hard_negatives``), ``DensePassageRetrieval``
(``embedder``, ``retriever``, ``build_retrieval_database``, ``retrieve_for_text``,
``retrieve_for_image``), ``create_dpr_model``.

What changes underneath:
  * ``semantic_index`` is a :class:`RadarIndex` (sm_100a kernels) instead of a CPU ``faiss.IndexFlatIP``;
    the per-32-row ``.cpu().numpy()`` hop of ``build_indices`` (:292) and the per-query host copy of
    ``retrieve`` (:312) are gone -- tensors stay on the GPU.
  * ``hybrid_alpha`` (declared at :187, never read by the reference) is honoured: when the index holds
    observation probabilities and the caller passes ``query_probs``, scores are
    ``alpha*cos - (1-alpha)*KL``.
  * deliberate behavioural difference: the reference swallows every exception and degrades (first-k
    passages with score 0.5, :315-318).  Here retrieval errors raise -- there is no CPU fallback.
"""
from __future__ import annotations

import hashlib
import logging
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from ._nvtx import annotate as _nvtx_annotate
from .config import RetrievalConfig
from .index import RadarIndex, project_normalize
from .knowledge import OBSERVATION_NAMES, observation_bits  # noqa: F401 (re-exported)

logger = logging.getLogger(__name__)

BIOMEDCLIP_NAME = "microsoft/BiomedCLIP-PubMedBERT_256-vit_base_patch16_224"  # dpr.py:208
FEATURE_DIM = 768  # width of the BiomedCLIP features the two projections consume (dpr.py:202-203)


def _hashed_features(texts: Sequence[str], device: torch.device) -> torch.Tensor:
    """Deterministic stand-in for BiomedCLIP features when the backbone cannot be loaded.

    The reference's degraded mode draws ``torch.randn`` features per call (dpr.py:221-224, :244), so the same
    passage embeds differently every time; a seeded draw per text keeps that mode usable for smoke runs."""
    out = torch.empty((len(texts), FEATURE_DIM), dtype=torch.float32)
    for i, t in enumerate(texts):
        seed = int.from_bytes(hashlib.sha256(t.encode("utf-8")).digest()[:8], "little") % (2**63 - 1)
        g = torch.Generator().manual_seed(seed)
        out[i] = torch.randn(FEATURE_DIM, generator=g)
    return out.to(device)


class CrossModalEmbedder(nn.Module):
    """BiomedCLIP features -> ``nn.Linear(768, embedding_dim)`` -> L2 normalise (dpr.py:191-267).

    ``backbone`` may be injected (anything exposing ``get_text_features(**tokens)`` /
    ``get_image_features(images)``) together with ``tokenizer``; otherwise the HF checkpoint is looked up
    in the local cache only (this build never touches the network)."""

    def __init__(self, config: RetrievalConfig, backbone=None, tokenizer=None):
        super().__init__()
        self.config = config
        self.device = torch.device('cuda' if (torch.cuda.is_available() and config.device == 'cuda') else 'cpu')
        self.model, self.tokenizer = backbone, tokenizer
        self.model_loaded = backbone is not None
        if backbone is None:
            self._load_biomedclip_model()
        self.text_projection = nn.Linear(FEATURE_DIM, config.embedding_dim).to(self.device)
        self.image_projection = nn.Linear(FEATURE_DIM, config.embedding_dim).to(self.device)

    def _load_biomedclip_model(self) -> None:
        try:
            from transformers import AutoModel, AutoTokenizer
            self.model = AutoModel.from_pretrained(BIOMEDCLIP_NAME, local_files_only=True).to(self.device).eval()
            self.tokenizer = AutoTokenizer.from_pretrained(BIOMEDCLIP_NAME, local_files_only=True)
            self.model_loaded = True
        except Exception as e:  # no checkpoint in the local cache
            logger.warning("BiomedCLIP not available locally (%s); using seeded stand-in features", type(e).__name__)
            self.model_loaded = False

    def _project(self, feats: torch.Tensor, proj: nn.Linear) -> torch.Tensor:
        if self.device.type != 'cuda':
            raise RuntimeError("CrossModalEmbedder needs a CUDA device: the projection + normalise prologue has no CPU path")
        return project_normalize(feats.to(self.device, torch.float32), proj.weight, proj.bias)  # fused tcgen05 GEMM + normalise

    @torch.no_grad()
    def encode_text(self, texts: List[str]) -> torch.Tensor:
        if self.model_loaded and self.tokenizer is not None:
            tok = self.tokenizer(texts, padding=True, truncation=True, return_tensors='pt', max_length=512)
            feats = self.model.get_text_features(**{k: v.to(self.device) for k, v in tok.items()})
        else:
            feats = _hashed_features(texts, self.device)
        return self._project(feats, self.text_projection)

    @torch.no_grad()
    def encode_image(self, images: torch.Tensor) -> torch.Tensor:
        if self.model_loaded and hasattr(self.model, 'get_image_features'):
            feats = self.model.get_image_features(images.to(self.device))
        else:
            g = torch.Generator().manual_seed(int(images.float().abs().sum().item() * 1e3) % (2**31 - 1))
            feats = torch.randn(images.size(0), FEATURE_DIM, generator=g).to(self.device)
        return self._project(feats, self.image_projection)


class HybridRetriever(nn.Module):
    def __init__(self, config: RetrievalConfig, embedder: CrossModalEmbedder, precision: str = 'bf16',
                 encode_batch: int = 1024):
        super().__init__()
        self.config = config
        self.embedder = embedder
        self.passages: List[str] = []
        self.semantic_index: Optional[RadarIndex] = None
        self.precision = precision
        self.encode_batch = encode_batch
        self.case_bits: Optional[torch.Tensor] = None  # uint16[N] observation sets, CheXpert-14 bit order

    @_nvtx_annotate("retriever.build_indices")
    def build_indices(self, passages: List[str], observations: List[List[str]],
                      observation_probs=None, embeddings=None) -> None:
        """Encode and index ``passages`` (dpr.py:278-303).

        ``observations`` (lists of observation names) is ignored by the reference; here it becomes the
        14-bit observation set of each case (used by the iterative-RAG re-rank).  ``observation_probs``
        (float32[N,14]) enables KL / hybrid scoring; ``embeddings`` (float32[N,d]) skips the encoder."""
        self.passages = passages
        self.semantic_index = None
        if not passages:
            logger.warning("No passages to index")
            return
        device = self.embedder.device if self.embedder is not None else torch.device(self.config.device)
        if device.type != 'cuda':
            raise RuntimeError("HybridRetriever.build_indices needs a CUDA device (there is no CPU fallback)")
        index = RadarIndex(self.config.embedding_dim, device=device, precision=self.precision)
        if embeddings is not None:
            index.add(embeddings)
        else:
            for i in range(0, len(passages), self.encode_batch):  # stays on the GPU: no per-batch host copy
                index.add(self.embedder.encode_text(passages[i:i + self.encode_batch]))
        if index.ntotal != len(passages):
            raise RuntimeError(f"indexed {index.ntotal} rows for {len(passages)} passages")
        if observation_probs is not None:
            index.add_observations(observation_probs)
        if observations:
            bits = np.fromiter((observation_bits(o) for o in observations), dtype=np.uint16, count=len(observations))
            self.case_bits = torch.from_numpy(bits.astype(np.int16)).to(device)  # reinterpret as uint16 in C
        self.semantic_index = index
        logger.info("GPU index built: %d passages", index.ntotal)

    @_nvtx_annotate("retriever.retrieve")
    def retrieve(self, query_embed: torch.Tensor, k: int = None, query_probs=None, mask=None
                 ) -> Tuple[List[str], List[float]]:
        """Best-first (passages, scores) for ONE query embedding (dpr.py:305-318)."""
        if k is None:
            k = self.config.num_retrieved
        k = min(k, len(self.passages))
        if k <= 0 or not self.semantic_index:
            return [], []
        use_kl = query_probs is not None and self.semantic_index.logq16 is not None
        scores, ids = self.semantic_index.search(
            query_embed.reshape(1, -1), k,
            query_probs=None if not use_kl else torch.as_tensor(query_probs).reshape(1, -1),
            mask=None if (mask is None or not use_kl) else torch.as_tensor(mask).reshape(1, -1),
            alpha=self.config.hybrid_alpha, mode='hybrid' if use_kl else 'dpr')
        ids_h, scores_h = ids[0].tolist(), scores[0].tolist()  # the one device->host read of a query
        return [self.passages[i] for i in ids_h], [float(s) for s in scores_h]

    def retrieve_batch(self, query_embeds: torch.Tensor, k: int = None, query_probs=None, mask=None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Batched form: (scores float32[Q,k], ids int64[Q,k]) CUDA tensors; no host synchronisation."""
        if k is None:
            k = self.config.num_retrieved
        k = min(k, len(self.passages))
        if not self.semantic_index or k <= 0:
            raise RuntimeError("retrieve_batch on an empty index")
        use_kl = query_probs is not None and self.semantic_index.logq16 is not None
        return self.semantic_index.search(query_embeds, k, query_probs=query_probs if use_kl else None,
                                          mask=mask if use_kl else None, alpha=self.config.hybrid_alpha,
                                          mode='hybrid' if use_kl else 'dpr')

    def retrieve_with_hard_negatives(self, query_embed: torch.Tensor, k: int = None, num_negatives: int = 3) -> Dict:
        if k is None:
            k = self.config.num_retrieved
        retrieved, scores = self.retrieve(query_embed, k + num_negatives)
        return {
            'positives': retrieved[:k],
            'negatives': retrieved[k:k + num_negatives],
            'positive_scores': scores[:k],
            'negative_scores': scores[k:k + num_negatives],
        }


    def retrieve_batch_with_hard_negatives(self, query_embeds: torch.Tensor, k: int = None, num_negatives: int = 3,
                                           query_probs=None, mask=None) -> Dict[str, torch.Tensor]:
        """Hard-negative mining for a whole batch (dpr.py:320-331 asks for ``k + num_negatives`` per query and slices):
        ONE search for ``k + num_negatives`` results per query, split on the device -- ranks ``[0, k)`` are the
        positives, ranks ``[k, k + num_negatives)`` the hard negatives.  CUDA tensors: ``positives`` int64[Q,k],
        ``negatives`` int64[Q,n], ``positive_scores`` / ``negative_scores`` float32; no host synchronisation.  Like the
        reference, fewer than ``k + num_negatives`` passages shrink the negatives first."""
        if k is None:
            k = self.config.num_retrieved
        if not self.semantic_index or k <= 0:
            raise RuntimeError("retrieve_batch_with_hard_negatives on an empty index")
        scores, ids = self.retrieve_batch(query_embeds, k + num_negatives, query_probs=query_probs, mask=mask)
        return {
            'positives': ids[:, :k].contiguous(),
            'negatives': ids[:, k:k + num_negatives].contiguous(),
            'positive_scores': scores[:, :k].contiguous(),
            'negative_scores': scores[:, k:k + num_negatives].contiguous(),
        }


class DensePassageRetrieval(nn.Module):
    def __init__(self, config: RetrievalConfig, backbone=None, tokenizer=None, precision: str = 'bf16'):
        super().__init__()
        self.config = config
        self.embedder = CrossModalEmbedder(config, backbone=backbone, tokenizer=tokenizer)
        self.retriever = HybridRetriever(config, self.embedder, precision=precision)

    def build_retrieval_database(self, passages: List[str], observations: List[List[str]], **kw) -> None:
        self.retriever.build_indices(passages, observations, **kw)

    def retrieve_for_text(self, text: str, k: int = None, **kw) -> Tuple[List[str], List[float]]:
        query_embed = self.embedder.encode_text([text]).squeeze(0)
        return self.retriever.retrieve(query_embed, k, **kw)

    def retrieve_for_image(self, image: torch.Tensor, k: int = None, **kw) -> Tuple[List[str], List[float]]:
        query_embed = self.embedder.encode_image(image.unsqueeze(0)).squeeze(0)
        return self.retriever.retrieve(query_embed, k, **kw)


def create_dpr_model(device: str = 'cuda') -> DensePassageRetrieval:
    return DensePassageRetrieval(RetrievalConfig(device=device))


def make_retrieval_function(dpr: DensePassageRetrieval,
                            probs_provider: Optional[Callable[[str], Optional[Sequence[float]]]] = None
                            ) -> Callable[[str, int], Tuple[List[str], List[float]]]:
    """Adapter with the exact signature ``retrieval_function(query: str, k: int)`` that
    ``IterativeRetrievalAugmentedGeneration.generate_with_iterative_retrieval`` calls once per round
    (modeling_iterative_rag.py:199, :237).  ``probs_provider(query)`` may return the 14 observation
    probabilities of the case being re-queried to get hybrid scoring."""
    def retrieval_function(query: str, k: int) -> Tuple[List[str], List[float]]:
        probs = probs_provider(query) if probs_provider is not None else None
        return dpr.retrieve_for_text(query, k, query_probs=probs)
    return retrieval_function


__all__ = [
    "RetrievalConfig", "CrossModalEmbedder", "HybridRetriever", "DensePassageRetrieval", "create_dpr_model",
    "make_retrieval_function", "OBSERVATION_NAMES",
]
