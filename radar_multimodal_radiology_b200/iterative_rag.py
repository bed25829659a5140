"""Re-retrieval rounds of iterative RAG on top of the GPU index.

Mirrors the caller-side contract of ``annotate_retrieve/modeling_iterative_rag.py``:
  * ``IterativeRAGConfig`` (:12-20), ``ObservationDetector`` (:23-61), ``ConsistencyVerifier`` (:64-107),
    ``TargetedRetriever`` (:110-152), ``IterativeRetrieval`` (:155-185),
    ``IterativeRetrievalAugmentedGeneration.generate_with_iterative_retrieval`` (:198-275) and
    ``create_iterative_rag_model`` (:317-319) keep their names, arguments and return shapes, so code
    written against the reference module runs unchanged;
  * ``retrieval_function(query: str, k: int) -> (List[str], List[float])`` is the seam (:199, :237):
    pass ``dense_passage_retrieval.make_retrieval_function(dpr)``.
The generator (an MLLM) is external to this path and stays an injected callable.

What is new is the *batched* round driver the 10M-corpus configs need (BASELINE config 5: 3 rounds x
16k queries with observation masks): :func:`batched_retrieval_round` runs one round for a whole batch
of cases as one masked search plus the bitmask form of the overlap re-rank (:127-152) on the GPU.
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, List, Optional, Sequence, Set, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from ._nvtx import annotate as _nvtx_annotate
from .config import IterativeRAGConfig

logger = logging.getLogger(__name__)

# the detector's default vocabulary (modeling_iterative_rag.py:30-36) -- NOT the CheXpert-14 order
DEFAULT_RAG_VOCAB = [
    "Atelectasis", "Cardiomegaly", "Consolidation", "Edema", "Pleural Effusion", "Pneumonia",
    "Pneumothorax", "No Finding", "Fracture", "Support Devices", "Enlarged Cardiomediastinum",
    "Lung Opacity", "Pulmonary Edema", "Rib Fracture",
]


class ObservationDetector(nn.Module):
    def __init__(self, config: IterativeRAGConfig):
        super().__init__()
        self.config = config
        self.observation_vocab = config.observation_vocab or list(DEFAULT_RAG_VOCAB)

    def detect_observations(self, text: str) -> Set[str]:
        if not text:
            return set()
        low = text.lower()
        return {o for o in self.observation_vocab if o.lower() in low}

    def find_missing_observations(self, generated_text: str, reference_text: str) -> Set[str]:
        return self.detect_observations(reference_text) - self.detect_observations(generated_text)


class ConsistencyVerifier(nn.Module):
    def __init__(self, config: IterativeRAGConfig):
        super().__init__()
        self.config = config
        self.observation_detector = ObservationDetector(config)

    def _sets(self, generations: Sequence[str]) -> List[Set[str]]:
        return [self.observation_detector.detect_observations(g) for g in generations]

    def compute_consistency(self, generations: List[str]) -> float:
        if len(generations) < 2:
            return 1.0
        sets = self._sets(generations)
        union = set().union(*sets)
        if not union:
            return 1.0
        return float(len(set.intersection(*sets)) / len(union))

    def find_consistent_observations(self, generations: List[str]) -> Set[str]:
        sets = self._sets(generations)
        return set.intersection(*sets) if sets else set()


class TargetedRetriever(nn.Module):
    def __init__(self, config: IterativeRAGConfig):
        super().__init__()
        self.config = config

    def build_retrieval_query(self, missing_observations: Set[str], image_context: str = "") -> str:
        if not missing_observations:
            return "general findings"
        query = "Cases with " + ", ".join(list(missing_observations)[:5])
        return query + (f" in {image_context}" if image_context else "")

    def rank_retrieved_passages(self, passages: List[str], missing_observations: Set[str]
                                ) -> List[Tuple[str, float]]:
        """coverage + diversity bonus, best first (modeling_iterative_rag.py:127-152)."""
        if not passages or not missing_observations:
            return [(p, 0.5) for p in passages]
        det = ObservationDetector(self.config)
        m = len(missing_observations)
        ranked = []
        for p in passages:
            overlap = len(det.detect_observations(p) & missing_observations)
            ranked.append((p, overlap / (m + 1e-8) + min(overlap / max(m, 1), 1.0) * 0.2))
        ranked.sort(key=lambda t: t[1], reverse=True)
        return ranked


class IterativeRetrieval(nn.Module):
    def __init__(self, config: IterativeRAGConfig):
        super().__init__()
        self.config = config
        self.targeted_retriever = TargetedRetriever(config)
        self.observation_detector = ObservationDetector(config)

    def initialize_retrieval_state(self) -> Dict:
        return {'iteration': 0, 'retrieved_passages': [], 'retrieved_scores': [],
                'missing_observations': set(), 'cumulative_findings': set()}

    def update_retrieval_state(self, state: Dict, new_passages: List[str], new_scores: List[float],
                               generated_text: str) -> Dict:
        state['retrieved_passages'].extend(new_passages)
        state['retrieved_scores'].extend(new_scores)
        state['cumulative_findings'].update(self.observation_detector.detect_observations(generated_text))
        state['iteration'] += 1
        return state


class IterativeRetrievalAugmentedGeneration(nn.Module):
    def __init__(self, config: IterativeRAGConfig):
        super().__init__()
        self.config = config
        self.observation_detector = ObservationDetector(config)
        self.consistency_verifier = ConsistencyVerifier(config)
        self.targeted_retriever = TargetedRetriever(config)
        self.iterative_retrieval = IterativeRetrieval(config)

    def generate_with_iterative_retrieval(self, initial_findings: str, retrieval_function: Callable,
                                          generation_function: Callable,
                                          reference_text: Optional[str] = None) -> Dict:
        """Up to ``num_iterations`` rounds of generate -> find missing observations -> re-retrieve ->
        re-rank -> extend the context (modeling_iterative_rag.py:198-275).  As in the reference an
        exception inside a round is logged and ends the loop (:258-260)."""
        state = self.iterative_retrieval.initialize_retrieval_state()
        generations: List[str] = []
        context = initial_findings
        for it in range(self.config.num_iterations):
            try:
                text = generation_function(context)
                generations.append(text)
                if reference_text:
                    missing = self.observation_detector.find_missing_observations(text, reference_text)
                else:
                    if self.consistency_verifier.compute_consistency(generations) >= self.config.consistency_threshold:
                        break
                    missing = set()
                state['missing_observations'] = missing
                if not missing:
                    break
                query = self.targeted_retriever.build_retrieval_query(missing)
                passages, _scores = retrieval_function(query, self.config.top_k)
                if not passages:
                    break
                ranked = self.targeted_retriever.rank_retrieved_passages(passages, missing)
                state = self.iterative_retrieval.update_retrieval_state(
                    state, [p for p, _ in ranked], [s for _, s in ranked], text)
                top = [p for p, _ in ranked[:2]]
                if top:
                    context = text + "\n\nRetrieved Evidence:\n" + "\n".join(top)
            except Exception as e:  # noqa: BLE001 - reference behaviour: log and stop iterating
                logger.error("Error in iteration %d: %s", it, e)
                break
        return {
            'generations': generations,
            'all_generations': list(generations),
            'retrieved_passages': state['retrieved_passages'],
            'retrieved_scores': state['retrieved_scores'],
            'iterations': state['iteration'],
            'final_consistency': self.consistency_verifier.compute_consistency(generations),
            'consistent_findings': self.consistency_verifier.find_consistent_observations(generations),
            'cumulative_findings': state['cumulative_findings'],
            'final_text': generations[-1] if generations else initial_findings,
        }

    def generate_with_verification(self, input_text: str, generation_function: Callable,
                                   num_samples: int = 3) -> Dict:
        generations: List[str] = []
        try:
            for _ in range(num_samples):
                generations.append(generation_function(input_text))
            det = self.observation_detector
            return {
                'generations': generations,
                'best_generation': max(generations, key=lambda g: len(det.detect_observations(g))),
                'consistency_score': self.consistency_verifier.compute_consistency(generations),
                'consistent_observations': self.consistency_verifier.find_consistent_observations(generations),
                'all_observations': set().union(*[det.detect_observations(g) for g in generations]),
            }
        except Exception as e:  # noqa: BLE001 - reference behaviour (:305-314)
            logger.error("Error in verification: %s", e)
            return {'generations': generations, 'best_generation': input_text, 'consistency_score': 0.0,
                    'consistent_observations': set(), 'all_observations': set()}


def create_iterative_rag_model(num_observations: int = 14, device: str = 'cuda'
                               ) -> IterativeRetrievalAugmentedGeneration:
    return IterativeRetrievalAugmentedGeneration(IterativeRAGConfig(device=device))


# ---------------------------------------------------------------------------------------------------
# batched GPU round
# ---------------------------------------------------------------------------------------------------
def rerank_overlap(case_bits: torch.Tensor, missing_bits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Bitmask form of ``rank_retrieved_passages`` for a batch: ``case_bits`` int16/uint16 [Q,k] (observation
    set of every retrieved case), ``missing_bits`` [Q].  Returns (scores float64[Q,k] in the original
    passage order, order int32[Q,k] = stable best-first permutation).  ``radar_rerank_overlap``."""
    if case_bits.device.type != "cuda":
        raise RuntimeError("rerank_overlap runs on a CUDA device only")
    cb = case_bits.contiguous().to(torch.int16)
    mb = missing_bits.contiguous().to(torch.int16)
    q, k = cb.shape
    scores = torch.empty((q, k), dtype=torch.float64, device=cb.device)
    order = torch.empty((q, k), dtype=torch.int32, device=cb.device)
    with L.device_guard(cb.device):
        L.check(L.lib().radar_rerank_overlap(L.ptr(cb), L.ptr(mb), q, k, L.ptr(scores), L.ptr(order),
                                             L.current_stream_ptr(cb.device)), "radar_rerank_overlap")
    return scores, order


def gather_case_bits(table: torch.Tensor, ids: torch.Tensor, idx_offset: int = 0) -> torch.Tensor:
    """int16[Q,k] observation sets of the retrieved ids (``radar_gather_bits``)."""
    t = table.contiguous().to(torch.int16)
    ids = ids.contiguous()
    out = torch.empty(ids.shape, dtype=torch.int16, device=ids.device)
    with L.device_guard(ids.device):
        L.check(L.lib().radar_gather_bits(L.ptr(t), t.shape[0], L.ptr(ids), ids.shape[0], ids.shape[1], idx_offset,
                                          L.ptr(out), L.current_stream_ptr(ids.device)), "radar_gather_bits")
    return out


@_nvtx_annotate("rag.batched_retrieval_round")
def batched_retrieval_round(index, query_embeds: Optional[torch.Tensor], query_probs: torch.Tensor,
                            missing_bits: torch.Tensor, case_bits_table: Optional[torch.Tensor], k: int,
                            alpha: float = 0.5, mode: Optional[str] = None, search_fn: Optional[Callable] = None
                            ) -> Dict[str, torch.Tensor]:
    """One re-retrieval round for a batch of cases, entirely on the GPU.

    ``missing_bits`` int16[Q]: CheXpert-14 bit set of the observations each case is still missing; it is
    the observation mask of the KL term (an empty set leaves the query unmasked) and the target of the
    overlap re-rank.  ``search_fn`` defaults to ``index.search`` (pass ``ShardedRadarIndex.search`` for a
    row-sharded corpus).  Returns ids / scores (retrieval order), rerank_scores and rerank_order."""
    from .knowledge import NUM_OBSERVATIONS
    mb = missing_bits.to(torch.int64)
    mask = ((mb[:, None] >> torch.arange(NUM_OBSERVATIONS, device=mb.device)[None, :]) & 1).to(torch.uint8)
    mask[mb == 0] = 1
    fn = search_fn or index.search
    scores, ids = fn(query_embeds, k, query_probs=query_probs, mask=mask, alpha=alpha, mode=mode)
    out = {"scores": scores, "ids": ids}
    if case_bits_table is not None:
        cb = gather_case_bits(case_bits_table, ids, 0)
        out["rerank_scores"], out["rerank_order"] = rerank_overlap(cb, missing_bits)
    return out
