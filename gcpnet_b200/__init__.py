"""gcpnet_b200 -- B200-native GCPNet message-passing layer (GCPInteractions / GCPMessagePassing).

Python host code over hand-written sm_100a kernels behind the C ABI of include/gcpnet_b200.h.
"""
from .scalar_vector import ScalarVector  # noqa: F401
from .interactions import GCPInteractions, GCP2Params, localize, graph_views, clear_graph_cache  # noqa: F401
from .graphs import GraphedStep, prepack  # noqa: F401

__all__ = ["GCPInteractions", "GCP2Params", "ScalarVector", "localize", "graph_views", "clear_graph_cache", "GraphedStep", "prepack"]
