"""Line-graph ("dual") construction of the DP-GSAT fork on the GPU.

  reference src/datasets/mutag_dual.py:342-378   dual nodes = directed primal edges, dual edges between primal edges
                                                 that share their FIRST endpoint (group_by_first), both directions
  reference src/datasets/mutag_dual.py:536-548   optional relabelling: the two directions of a primal edge (consecutive
                                                 rows of the edge list) share one dual node id   (``halve=True``)

  reference src/datasets/ba_2motifs_dual.py:33-73  the dense-matrix variant: dual nodes = UNDIRECTED primal edges numbered
                                                 in order of first appearance in the adjacency matrix, dual edges between
                                                 edges that share a node, in dense_to_sparse (row-major) order
                                                 (``line_graph_dual_dense``)

The reference builds these with Python dict / matrix loops over every edge at dataset-processing time; here it is two
kernels on top of the CSC that K0 already built for the primal batch (plus, for the dense variant, a rank and a sort of
the dual edge keys on the device).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from ._lib import lib, ptr, stream
from .index import get_graph_index


def line_graph_dual(edge_index: torch.Tensor, batch: torch.Tensor, num_graphs: Optional[int] = None,
                    halve: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dual_edge_index int64 [2, E_d], dual_batch int64 [E] or [E/2]) of the primal graph batch.

    Dual node e is primal edge e (``halve``: edge pair e >> 1, 0-based); dual_batch[e] = batch[src(e)]."""
    gi = get_graph_index(edge_index, batch, num_graphs)
    E, N = gi.E, gi.N
    dev = edge_index.device
    L = lib()
    if halve and E % 2:
        raise ValueError('halve=True needs both directions of every primal edge as consecutive rows (E even)')
    offs = torch.empty(max(E, 1), dtype=torch.int64, device=dev)[:E]
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = int(L.cdll.gsatb_line_graph_workspace(N, E))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    # members of each source group in primal edge order (K0's eid_by_src orders a group by destination instead)
    members = torch.empty(max(E, 1), dtype=torch.int32, device=dev)[:E]
    so_bytes = int(L.cdll.gsatb_stable_order_workspace(E))
    so_ws = torch.empty(so_bytes, dtype=torch.uint8, device=dev)
    L.call('gsatb_stable_order', ptr(gi.src), E, N, ptr(members), ptr(so_ws), ctypes.c_size_t(so_bytes), stream())
    L.call('gsatb_line_graph_count', ptr(gi.src), ptr(gi.rowptr_src), ptr(members), N, E, ptr(offs), ptr(total),
           ptr(ws), ctypes.c_size_t(ws_bytes), stream())
    Ed = int(total.item())                 # the one host sync: the output size is data dependent
    dual_ei = torch.empty((2, Ed), dtype=torch.int64, device=dev)
    nd = E // 2 if halve else E
    dual_batch = torch.empty(nd, dtype=torch.int64, device=dev)
    batch_c = batch.contiguous()
    L.call('gsatb_line_graph_fill', ptr(gi.src), ptr(gi.rowptr_src), ptr(members), ptr(offs),
           ptr(batch_c), N, E, int(halve), ptr(dual_ei), Ed, ptr(dual_batch), stream())
    return dual_ei, dual_batch


def line_graph_dual_dense(edge_index: torch.Tensor, batch: torch.Tensor, num_graphs: Optional[int] = None
                          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The BA-2Motifs dual of the fork (reference src/datasets/ba_2motifs_dual.py:33-73) for a whole batch on the device.

    Dual node = undirected primal edge {u, v}, numbered in the order the reference's scan of the adjacency matrix meets
    it (rows ascending, columns ascending == ascending (min, max)), graph after graph; dual edge (i, j), i != j, for every
    two edges that share a node, listed in ``dense_to_sparse`` order (ascending (i, j), no duplicates).  Needs a
    symmetric, duplicate-free primal edge set without self loops (what a 0/1 adjacency matrix holds).

    Returns (dual_edge_index int64 [2, E_d], dual_batch int64 [E/2], und_id int64 [E]: the dual node of every directed
    primal edge)."""
    gi = get_graph_index(edge_index, batch, num_graphs)
    if not gi.symmetric or gi.has_duplicates:
        raise ValueError('the dense dual needs a symmetric, duplicate-free primal edge set (a 0/1 adjacency matrix)')
    src, dst = gi.src.long(), gi.dst.long()
    if bool((src == dst).any()):
        raise ValueError('the dense dual is defined for graphs without self loops (ba_2motifs_dual.py:44)')
    E = gi.E
    M = E // 2
    order = gi.eid_by_src.long()                       # directed edges in ascending (src, dst): the matrix scan order
    first = (src < dst)[order]                         # the scan numbers an edge when it meets its (min, max) entry
    rank = torch.cumsum(first.long(), 0) - 1
    und = torch.empty(E, dtype=torch.int64, device=edge_index.device)
    und[order[first]] = rank[first]
    rev = gi.rev.long()
    und[rev[order[first]]] = rank[first]
    pairs, _ = line_graph_dual(edge_index, batch, num_graphs)              # directed edges sharing their source node
    key = torch.unique(und[pairs[0]] * max(M, 1) + und[pairs[1]])          # sorted: dense_to_sparse order
    dual_ei = torch.stack([key // max(M, 1), key % max(M, 1)])
    dual_batch = torch.empty(M, dtype=torch.int64, device=edge_index.device)
    dual_batch[rank[first]] = batch[src[order[first]]]
    return dual_ei, dual_batch, und


def dense_dual_node_features(x: torch.Tensor, edge_index: torch.Tensor, und_id: torch.Tensor) -> torch.Tensor:
    """Features of the dense dual's nodes (reference src/datasets/ba_2motifs_dual.py:48): for the undirected primal edge
    {u, v}, u < v, ``cat(x[u], x[v])``.  ``und_id`` as returned by line_graph_dual_dense."""
    src, dst = edge_index[0], edge_index[1]
    first = src < dst
    out = torch.empty((int(und_id.numel()) // 2, 2 * x.shape[1]), dtype=x.dtype, device=x.device)
    out[und_id[first]] = torch.cat([x[src[first]], x[dst[first]]], dim=1)
    return out
