"""Shim: the reference's import path -> radar_multimodal_radiology_b200.dense_passage_retrieval."""
from radar_multimodal_radiology_b200.dense_passage_retrieval import (  # noqa: F401
    CrossModalEmbedder, DensePassageRetrieval, HybridRetriever, RetrievalConfig, create_dpr_model,
    make_retrieval_function,
)
