"""GNN surrogate model for wing stress prediction (same exports as the reference's
``src/deep_fem_uav_wing/gnn/__init__.py:3-6``), backed by ``libdfw_b200.so``."""

from deep_fem_uav_wing.gnn.dataset import WingStressDataset, build_graph_data
from deep_fem_uav_wing.gnn.model import GraphSAGEModel

__all__ = ["WingStressDataset", "build_graph_data", "GraphSAGEModel"]
