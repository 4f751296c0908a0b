"""Graph-sharded data parallelism: one process per GPU, replicas of the ~70-200 k parameters, each
rank owns a contiguous range of graphs chosen on the prefix sum of edges per graph, and ONE all-reduce per step
over a single flat fp32 gradient bucket (NCCL over NVLink 5 / NVSwitch; payload < 1 MB, latency bound).

The reference has no distributed code at all (single device, src/run_gsat.py:1069); this is new capability.
Loss terms are means over GLOBAL counts: the local info loss is weighted by E_local/E_global and the local
prediction loss by G_local/G_global, so the SUM of the ranks' gradients equals the single-device gradient (up to
BatchNorm, whose batch statistics are shard-local here, as in torch DDP).

The module is device-agnostic torch code, so the N>1 logic is covered on CPU with the gloo backend.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    """All parameter gradients as views into one contiguous fp32 buffer: backward writes straight into the buffer
    that the all-reduce sends, no pack/unpack copies."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


def global_counts(num_edges: int, num_graphs: int, device, group=None):
    """(E_global, G_global) summed over ranks (one tiny all-reduce, done once per batch, not per step)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return int(num_edges), int(num_graphs)
    t = torch.tensor([float(num_edges), float(num_graphs)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0].item()), int(t[1].item())


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


class TrainStep:
    """forward_pass -> backward -> (all-reduce) -> Adam.step for one (sharded) batch.  ``gsat`` is any object with
    the reference's ``forward_pass(data, epoch, training)`` and the ``pred_scale`` / ``info_scale`` attributes."""

    def __init__(self, gsat, lr: float = 1e-3, weight_decay: float = 0.0, group=None, fused_adam: Optional[bool] = None):
        self.gsat = gsat
        self.group = group
        params = list(gsat.extractor.parameters()) + list(gsat.clf.parameters())   # order of src/run_gsat.py:1007
        self.bucket = FlatGradBucket(params)
        if fused_adam is None:
            fused_adam = self.bucket.flat.is_cuda
        self.optimizer = torch.optim.Adam(self.bucket.params, lr=lr, weight_decay=weight_decay, fused=fused_adam)
        self._counts = {}

    def set_shard_weights(self, data):
        key = (int(data.edge_index.shape[1]), int(data.batch.numel()), int(data.y.shape[0]))
        if key not in self._counts:
            learn = getattr(self.gsat, 'learn_edge_att', True)
            n_att = data.edge_index.shape[1] if learn or getattr(self.gsat, 'info_on', 'att') == 'edge_att' \
                else data.batch.numel()
            n_g = int(data.y.shape[0])
            eg, gg = global_counts(n_att, n_g, data.x.device, self.group)
            self._counts[key] = (n_att / max(eg, 1), n_g / max(gg, 1))
        self.gsat.info_scale, self.gsat.pred_scale = self._counts[key]

    def __call__(self, data, epoch: int, noise_u=None):
        self.set_shard_weights(data)
        edge_att, loss, loss_dict, clf_logits = self.gsat.forward_pass(data, epoch, True, noise_u=noise_u)
        self.bucket.zero()
        loss.backward()
        self.bucket.all_reduce(self.group)
        self.optimizer.step()
        return edge_att, loss, loss_dict, clf_logits
