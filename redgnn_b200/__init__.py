"""redgnn_b200 -- B200-native (sm_100a) implementation of RED-GNN's query-conditioned
relational-digraph propagation path, behind the reference's Python surface.

Importing this package loads libredgnn_b200.so (ctypes); there is no CPU / PyTorch fallback.
"""
from . import _lib
from .graph import DeviceGraph, Frontier
from .ops import Segments, edge_aggregate
from .layers import GNNLayer, RedGNN
from .data import TransductiveLoader, InductiveLoader
from .transductive.models import RED_GNN_trans
from .inductive.models import RED_GNN_induc
from . import metrics, dist

__all__ = ["DeviceGraph", "Frontier", "Segments", "edge_aggregate", "GNNLayer", "RedGNN", "TransductiveLoader",
           "InductiveLoader", "RED_GNN_trans", "RED_GNN_induc"]
__version__ = "0.1.0"
