"""Line-graph ("dual") construction of the DP-GSAT fork on the GPU.

  reference src/datasets/mutag_dual.py:342-378   dual nodes = directed primal edges, dual edges between primal edges
                                                 that share their FIRST endpoint (group_by_first), both directions
  reference src/datasets/mutag_dual.py:536-548   optional relabelling: the two directions of a primal edge (consecutive
                                                 rows of the edge list) share one dual node id   (``halve=True``)

The reference builds this with Python dict loops over every edge at dataset-processing time; here it is two kernels on
top of the CSC that K0 already built for the primal batch.
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
