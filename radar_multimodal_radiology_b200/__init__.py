"""radar_multimodal_radiology_b200 -- B200-native (sm_100a) implementation of RADAR's case-retrieval hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + C ABI), the host-side mirror of the
reference's retrieval interface, and the multi-GPU shard/merge plumbing.  See DESIGN.md.
"""
from .config import IterativeRAGConfig, KnowledgeConfig, RetrievalConfig, load_knowledge_config

__all__ = ["RetrievalConfig", "IterativeRAGConfig", "KnowledgeConfig", "load_knowledge_config"]
__version__ = "0.1.0"
