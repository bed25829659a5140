"""``src.knowledge`` -- the package north_star names for KL-divergence observation retrieval.  The
reference's file is 0 bytes; the content is radar_multimodal_radiology_b200.knowledge."""
from radar_multimodal_radiology_b200.knowledge import (  # noqa: F401
    NUM_OBSERVATIONS, OBSERVATION_NAMES, ObservationKLRetriever, bits_to_mask, missing_observation_mask,
    observation_bits,
)
from radar_multimodal_radiology_b200.config import KnowledgeConfig, load_knowledge_config  # noqa: F401
