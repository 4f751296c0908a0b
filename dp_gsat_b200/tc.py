"""Python side of the tensor-core (tcgen05) ops: weight preparation, fused extractor MLP, fused GIN node MLP.

Precision: operands are rounded to bf16, accumulation is fp32 (TMEM).  These ops are selected with
``precision='bf16'`` on the modules of nn.py; ``precision='fp32'`` keeps the strict fp32 path (library sgemm +
segment-norm kernels) that the rtol-1e-5 parity tests use.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ._lib import lib, ptr, stream
from .index import GraphIndex


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def prep_weight(w: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """fp32 [OUT, K] -> zero-padded bf16 [pad128(OUT), pad64(K)] (or the same for W^T), the TMA-ready layout."""
    w = w.detach().contiguous()
    OUT, K = w.shape
    rows, cols = (K, OUT) if transpose else (OUT, K)
    wp = torch.empty((_pad(rows, 128), _pad(cols, 64)), dtype=torch.bfloat16, device=w.device)
    lib().call('gsatb_tc_prep_weight', ptr(w), OUT, K, int(transpose), ptr(wp), stream())
    return wp


def linear(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], out_features: int,
           in_scale: Optional[torch.Tensor] = None, in_shift: Optional[torch.Tensor] = None, relu_out: bool = False,
           want_stats: bool = False):
    """out = act(pro(x) W^T + bias) on the tensor cores; optional per-channel (sum z, sum z^2) in fp64."""
    x = x.contiguous()
    rows, K = x.shape
    out = torch.empty((rows, out_features), dtype=torch.float32, device=x.device)
    part = stats = None
    if want_stats:
        part = torch.empty(int(lib().cdll.gsatb_tc_stat_partials_elems(out_features)), dtype=torch.float32,
                           device=x.device)
        stats = torch.empty(2 * out_features, dtype=torch.float64, device=x.device)
    lib().call('gsatb_tc_linear_fwd', ptr(x), K, ptr(in_scale), ptr(in_shift), ptr(wp), ptr(bias), ptr(out),
               out_features, int(relu_out), ptr(part), ptr(stats), rows, K, out_features, stream())
    return (out, stats) if want_stats else out


def extractor_forward(emb: torch.Tensor, gi: GraphIndex, w1: torch.Tensor, w2: torch.Tensor, w3: torch.Tensor,
                      b3: Optional[torch.Tensor], *, edge_mode: bool, pdrop: float, training: bool, seed: int,
                      mask1: Optional[torch.Tensor] = None, mask2: Optional[torch.Tensor] = None, eps: float = 1e-5):
    """Fused forward of the extractor MLP.  Returns (logit [rows, 1], saved) or None when the batch cannot be tiled
    (a graph with more than 128 rows)."""
    plan = gi.tile_plan('edge' if edge_mode else 'node')
    if plan is None:
        return None
    tile_row, tile_seg, T = plan
    emb = emb.contiguous()
    N, H = emb.shape
    C1 = w1.shape[0]
    rows = gi.E if edge_mode else gi.N
    seg_ptr = gi.edge_ptr if edge_mode else gi.node_ptr
    dev = emb.device
    w1p, w2p = prep_weight(w1), prep_weight(w2)
    xhat1 = torch.empty((rows, C1), dtype=torch.bfloat16, device=dev)
    rstd1 = torch.empty((gi.G, C1), dtype=torch.float32, device=dev)
    xhat2 = torch.empty((rows, H), dtype=torch.bfloat16, device=dev)
    rstd2 = torch.empty((gi.G, H), dtype=torch.float32, device=dev)
    logit = torch.empty((rows, 1), dtype=torch.float32, device=dev)
    L = lib()
    L.call('gsatb_tc_ext_fwd1', ptr(emb), ptr(gi.src) if edge_mode else None, ptr(gi.dst) if edge_mode else None,
           ptr(w1p), ptr(tile_row), ptr(tile_seg), ptr(seg_ptr), T, ptr(xhat1), ptr(rstd1), rows, H, C1,
           ctypes.c_float(eps), stream())
    w3f = w3.detach().reshape(-1).contiguous()
    L.call('gsatb_tc_ext_fwd2', ptr(xhat1), ptr(w2p), ptr(w3f), ptr(b3), ptr(mask1), ptr(mask2),
           ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training), ptr(tile_row), ptr(tile_seg), ptr(seg_ptr), T,
           ptr(xhat2), ptr(rstd2), ptr(logit), rows, C1, H, ctypes.c_float(eps), stream())
    return logit, dict(xhat1=xhat1, rstd1=rstd1, xhat2=xhat2, rstd2=rstd2, plan=plan)
