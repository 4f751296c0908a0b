"""dp_gsat_b200 -- B200-native (sm_100a) implementation of GSAT's per-step stochastic-attention
message-passing path, exposed through the reference's own Python surfaces.

(The repo-root entry ``dp-gsat_b200`` is a symlink to this directory: a Python package name cannot hold '-'.)

There is no CPU path and no eager fallback: importing the package loads libgsat_b200.so and fails loudly if it
has not been built (``python -m dp_gsat_b200.build``).
"""
import sys as _sys

from ._lib import lib as _lib

if 'dp_gsat_b200.build' not in getattr(_sys, 'orig_argv', []):     # `python -m dp_gsat_b200.build` creates the library
    _lib()  # fail at import time if the CUDA library is missing or does not export the declared C ABI

from .data import Batch  # noqa: E402
from .index import GraphIndex, get_graph_index, prefetch_graph_index, evict_graph_index, clear_index_cache, set_index_cache_capacity  # noqa: E402
from .nn import (GIN, GINConv, GINEConv, LEConv, SPMotifNet, ExtractorMLP, MLP, BatchSequential, InstanceNorm, Criterion, get_model,  # noqa: E402
                 get_preds, AtomEncoder, BondEncoder)
from .gsat import (GSAT, DualGSAT, is_undirected, transpose, reorder_like, get_r, concrete_sample,  # noqa: E402
                   lift_node_att_to_edge_att, gumbel_sigmoid, f1_sparsity_loss, info_loss)
from .pna import PNA, PNAConvSimple  # noqa: E402
from .dual import line_graph_dual, line_graph_dual_dense, dense_dual_node_features  # noqa: E402
from .metrics import get_precision_at_k, get_delta_kl  # noqa: E402
from .loader import Graph, PackedDataset, DeviceLoader  # noqa: E402
from . import ops, dense  # noqa: E402
from .dense import Linear  # noqa: E402

__all__ = ['Batch', 'GraphIndex', 'get_graph_index', 'prefetch_graph_index', 'evict_graph_index', 'clear_index_cache', 'set_index_cache_capacity', 'GIN', 'GINConv', 'GINEConv', 'LEConv', 'SPMotifNet', 'ExtractorMLP', 'MLP',
           'BatchSequential', 'InstanceNorm', 'Criterion', 'get_model', 'get_preds', 'AtomEncoder', 'BondEncoder',
           'GSAT', 'DualGSAT', 'is_undirected', 'transpose', 'reorder_like', 'get_r', 'concrete_sample',
           'lift_node_att_to_edge_att', 'gumbel_sigmoid', 'f1_sparsity_loss', 'info_loss', 'ops', 'PNA', 'PNAConvSimple', 'line_graph_dual', 'line_graph_dual_dense', 'dense_dual_node_features', 'get_precision_at_k', 'get_delta_kl',
           'Graph', 'PackedDataset', 'DeviceLoader', 'Linear', 'dense']
