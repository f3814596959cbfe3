"""gcpnet_b200 -- B200-native GCPNet message-passing layer (GCPInteractions / GCPMessagePassing).

Python host code over hand-written sm_100a kernels behind the C ABI of include/gcpnet_b200.h.
"""
from .scalar_vector import ScalarVector  # noqa: F401
from .interactions import GCPInteractions, GCPMessagePassing, GCP2Params, localize, graph_views, clear_graph_cache  # noqa: F401
from .graphs import GraphedStep, prepack  # noqa: F401
from .interactions import centralize, decentralize  # noqa: F401
from .modules import (GCP2, GCP3, GCPDropout, GCPEmbedding, GCPInteractions2, GCPLayerNorm, GCPMLPDecoder, GCPNetCPD,  # noqa: F401
                      GCPNetLBA, GCPNetNMS, GCPNetPSR, GCPNetRS)
from . import ddp, bucketing  # noqa: F401
from .bucketing import BatchSampler, BucketedSteps, pad_batch, bucket_shape  # noqa: F401
from .ddp import FlatGradients  # noqa: F401

__all__ = ["GCPInteractions", "GCPMessagePassing", "GCP2Params", "ScalarVector", "localize", "graph_views", "clear_graph_cache", "GraphedStep", "prepack", "FlatGradients", "ddp", "bucketing", "BatchSampler", "BucketedSteps", "pad_batch", "bucket_shape", "centralize", "decentralize", "GCP2", "GCPLayerNorm", "GCPEmbedding", "GCPNetNMS", "GCPNetCPD", "GCPMLPDecoder", "GCP3", "GCPInteractions2", "GCPDropout", "GCPNetLBA", "GCPNetPSR", "GCPNetRS"]
