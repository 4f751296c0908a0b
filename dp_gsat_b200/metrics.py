"""Per-batch explanation metrics of the trainer on the device.

  reference src/run_gsat.py:783-791  GSAT.get_precision_at_k   (Python loop over graphs, numpy argsort, .cpu() inputs)
  reference src/run_gsat.py:793-800  GSAT.get_delta_kl

Both return DEVICE tensors (no host sync); call ``.tolist()`` / ``.item()`` once per epoch, not per batch.
"""
from __future__ import annotations

import torch

from ._lib import lib, ptr, stream
from .index import get_graph_index


def get_precision_at_k(att: torch.Tensor, exp_labels: torch.Tensor, k: int, batch: torch.Tensor,
                       edge_index: torch.Tensor, num_graphs=None) -> torch.Tensor:
    """[G] float32: fraction of ground-truth explanation edges among the k highest-attention edges of each graph."""
    gi = get_graph_index(edge_index, batch, num_graphs)
    gi.require_graph_contiguous()
    a = att.detach().reshape(-1).float().contiguous()
    y = exp_labels.detach().reshape(-1).float().contiguous()
    if a.numel() != gi.E or y.numel() != gi.E:
        raise ValueError('att / exp_labels must hold one value per edge')
    out = torch.empty(gi.G, dtype=torch.float32, device=a.device)
    lib().call('gsatb_precision_at_k', ptr(a), ptr(y), ptr(gi.edge_ptr), gi.G, int(k), ptr(out), stream())
    return out


def get_delta_kl(exp_labels: torch.Tensor, att: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Device scalar; elementwise torch on the device (not a hot path: two reductions over E values)."""
    p = exp_labels.detach().reshape(-1).float().clamp(min=eps, max=1 - eps)
    r_uv = att.detach().reshape(-1).float().clamp(min=eps, max=1 - eps)
    r = r_uv.mean().clamp(min=eps, max=1 - eps)
    return (p * torch.log(r_uv / r) + (1 - p) * torch.log((1 - r_uv) / (1 - r))).sum()
