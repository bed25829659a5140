"""Shim: the reference's import path -> radar_multimodal_radiology_b200.iterative_rag."""
from radar_multimodal_radiology_b200.config import IterativeRAGConfig  # noqa: F401
from radar_multimodal_radiology_b200.iterative_rag import (  # noqa: F401
    ConsistencyVerifier, IterativeRetrieval, IterativeRetrievalAugmentedGeneration, ObservationDetector,
    TargetedRetriever, batched_retrieval_round, create_iterative_rag_model,
)
