"""The GSAT step and the functional surfaces the reference's step calls, backed by libgsat_b200.so.

  reference                                                  -> here
  example/gsat.py:12-117  GSAT (canonical single-graph step) -> GSAT
  src/run_gsat.py:860-885 get_r / sampling / concrete_sample / lift_node_att_to_edge_att -> same names
  src/run_gsat.py:182-187 gumbel_sigmoid, :151-180 f1_sparsity_loss (fork glue, elementwise) -> same names
  torch_geometric.utils.is_undirected, torch_sparse.transpose, utils.reorder_like          -> same names
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as tnn

from . import ops
from ._lib import lib, ptr, stream
from .index import GraphIndex, get_graph_index


# ------------------------------------------------------------------------------------------------------------
# functional drop-ins
# ------------------------------------------------------------------------------------------------------------
def is_undirected(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> bool:
    """torch_geometric.utils.is_undirected (src/run_gsat.py:232,242).  The answer is a flag K0 computed when the
    index of this edge_index was built; only the first call per batch reads it back from the device."""
    n = num_nodes if num_nodes is not None else (int(edge_index.max().item()) + 1 if edge_index.numel() else 0)
    gi = _index_for_edges(edge_index, n)
    return gi.symmetric


def _index_for_edges(edge_index: torch.Tensor, num_nodes: int) -> GraphIndex:
    from .index import _CACHE
    for key, (gi, ei, _) in _CACHE.items():           # prefer an index already built with the real batch vector
        if ei is edge_index or (key[0] == edge_index.data_ptr() and key[1] == tuple(edge_index.shape)
                                and key[2] == edge_index._version):
            return gi
    return get_graph_index(edge_index, None, num_nodes=num_nodes)


def transpose(index, value, m=None, n=None, coalesced: bool = False):
    """torch_sparse.transpose(index, value, m, n, coalesced=False): a pure row swap (src/run_gsat.py:243).  The
    returned index remembers what it is the transpose of, so reorder_like can use the cached reverse-edge map."""
    if coalesced:
        raise NotImplementedError('the reference only calls transpose(..., coalesced=False)')
    row, col = index[0], index[1]
    t = torch.stack([col, row], dim=0)
    t._gsatb_transpose_of = index
    return t, value


def reorder_like(from_edge_index, to_edge_index, values):
    """src/utils/utils.py:19-25.  Fast path (from is transpose(to)): one gather through the cached reverse-edge map.
    General path: both edge lists are ranked by K0 and matched position by position.  Raises the reference's
    ValueError when the two edge sets differ."""
    msg = 'Edges in from_edge_index and to_edge_index are different, impossible to match both.'
    if getattr(from_edge_index, '_gsatb_transpose_of', None) is to_edge_index:
        n = int(to_edge_index.max().item()) + 1 if to_edge_index.numel() else 0
        gi = _index_for_edges(to_edge_index, n)
        if not gi.symmetric:
            raise ValueError(msg)
        return ops.gather_reverse(values, gi.rev)
    n = int(max(from_edge_index.max().item(), to_edge_index.max().item())) + 1 if to_edge_index.numel() else 0
    if from_edge_index.shape != to_edge_index.shape:
        raise ValueError(msg)
    g_to = get_graph_index(to_edge_index, None, num_nodes=n)
    g_from = get_graph_index(from_edge_index, None, num_nodes=n)
    E = g_to.E
    mp = torch.empty(max(E, 1), dtype=torch.int32, device=values.device)[:E]
    mism = torch.zeros(1, dtype=torch.int32, device=values.device)
    lib().call('gsatb_match_orders', ptr(g_to.eid_by_src), ptr(g_from.eid_by_src), ptr(g_to.src), ptr(g_to.dst),
               ptr(g_from.src), ptr(g_from.dst), ptr(mp), ptr(mism), E, stream())
    if int(mism.item()) != 0:
        raise ValueError(msg)
    return ops.gather_reverse(values, mp)


def get_r(decay_interval, decay_r, current_epoch, init_r=0.9, final_r=0.5):
    """src/run_gsat.py:860-864 (host scalar)."""
    r = init_r - current_epoch // decay_interval * decay_r
    if r < final_r:
        r = final_r
    return r


def concrete_sample(att_log_logit, temp=1, training=True, noise_u: Optional[torch.Tensor] = None, seed: int = 0,
                    offset: int = 0):
    """src/run_gsat.py:877-885.  ``noise_u`` injects the uniform draw; otherwise Philox(seed, offset + e)."""
    att, _, _ = ops.sample_avg_info(att_log_logit, training=training, rev=None, average=False, noise_u=noise_u,
                                    temp=float(temp), want_info=False, seed=seed, offset=offset)
    return att


def lift_node_att_to_edge_att(node_att, edge_index, batch: Optional[torch.Tensor] = None):
    """src/run_gsat.py:870-875."""
    gi = get_graph_index(edge_index, batch, num_nodes=node_att.shape[0]) if batch is not None \
        else _index_for_edges(edge_index, node_att.shape[0])
    return ops.lift_node_att(node_att, gi)


def gumbel_sigmoid(logits, tau=1.0, eps=1e-10, noise_u: Optional[torch.Tensor] = None):
    """src/run_gsat.py:182-187 (fork glue; plain elementwise PyTorch, composes with the autograd ops above)."""
    U = torch.rand_like(logits) if noise_u is None else noise_u
    g = -torch.log(-torch.log(U + eps) + eps)
    return torch.sigmoid((logits + g) / tau)


def f1_sparsity_loss(p_uv, y_uv, eps=1e-6):
    """src/run_gsat.py:151-180 (fork glue)."""
    TP = (p_uv.view(-1) * y_uv.view(-1)).sum()
    P, G = p_uv.sum(), y_uv.sum()
    precision, recall = TP / (P + eps), TP / (G + eps)
    f1 = 2 * precision * recall / (precision + recall + eps)
    return (1 - f1) + p_uv.abs().mean()


def info_loss(att, r):
    """KL-to-Bernoulli(r) regulariser as an op of its own (fork: per-edge tensor ``r``, src/run_gsat.py:129-132)."""
    flat = att.reshape(-1, 1)
    # logit-free entry: feed att through the fused kernel in eval mode on logit(att) would lose precision, so the
    # standalone form is composed from elementwise torch ops; the fused form is used by GSAT.forward_pass.
    return (flat * torch.log(flat / r + 1e-6) + (1 - flat) * torch.log((1 - flat) / (1 - r + 1e-6) + 1e-6)).mean()


# ------------------------------------------------------------------------------------------------------------
# the step
# ------------------------------------------------------------------------------------------------------------
class GSAT(tnn.Module):
    """example/gsat.py:12-117.  forward_pass(data, epoch, training) -> (edge_att, loss, loss_dict, clf_logits).

    ``info_on='att'`` is upstream (loss on the pre-average attention, example/gsat.py:91); ``'edge_att'`` is the
    fork (src/run_gsat.py:276).  ``lazy_metrics=True`` keeps loss_dict values as device scalars instead of paying the
    reference's three .item() host syncs per step (example/gsat.py:34)."""

    def __init__(self, clf, extractor, criterion, optimizer=None, learn_edge_att=True, final_r=0.7, decay_interval=10,
                 decay_r=0.1, init_r=0.9, info_on: str = 'att', pred_loss_coef=1.0, info_loss_coef=1.0,
                 lazy_metrics: bool = False, seed: int = 0):
        super().__init__()
        self.clf, self.extractor, self.criterion, self.optimizer = clf, extractor, criterion, optimizer
        self.learn_edge_att = learn_edge_att
        self.final_r, self.decay_interval, self.decay_r, self.init_r = final_r, decay_interval, decay_r, init_r
        self.info_on = info_on
        self.pred_loss_coef, self.info_loss_coef = pred_loss_coef, info_loss_coef
        self.lazy_metrics = lazy_metrics
        self.seed = seed
        self.pred_scale = self.info_scale = 1.0   # data-parallel shard weights (G_local/G_global, E_local/E_global)
        self._step = 0
        try:
            self.device = next(self.parameters()).device
        except StopIteration:
            self.device = torch.device('cuda')
        # device-resident step counter (gsatb_set_step_counter): bumped once per training step ON THE DEVICE, so a
        # CUDA graph of the step (parallel.TrainStep.enable_cuda_graph) draws fresh noise / dropout on every replay
        self.step_counter = lib().step_counter(self.device) if self.device.type == 'cuda' else None

    def __loss__(self, info_mean, clf_logits, clf_labels, epoch):
        pred_loss = self.criterion(clf_logits, clf_labels) * (self.pred_loss_coef * self.pred_scale)
        il = info_mean * (self.info_loss_coef * self.info_scale)
        loss = pred_loss + il
        if self.lazy_metrics:
            loss_dict = {'loss': loss.detach(), 'pred': pred_loss.detach(), 'info': il.detach()}
        else:
            loss_dict = {'loss': loss.item(), 'pred': pred_loss.item(), 'info': il.item()}
        return loss, loss_dict

    def forward_pass(self, data, epoch, training, noise_u: Optional[torch.Tensor] = None, r=None):
        self.clf._enc_scope = {}          # the two GNN passes of this step share node_encoder(x) (nn._encode_once)
        if self.step_counter is not None and training:
            self.step_counter.add_(1)
        try:
            return self._forward_pass(data, epoch, training, noise_u, r)
        finally:
            self.clf._enc_scope = None

    def _forward_pass(self, data, epoch, training, noise_u, r):
        gi = get_graph_index(data.edge_index, data.batch, getattr(data, 'num_graphs', None) or None)
        emb = self.clf.get_emb(data.x, data.edge_index, batch=data.batch, edge_attr=data.edge_attr)
        att_log_logits = self.extractor(emb, data.edge_index, data.batch)
        if r is None:
            r = get_r(self.decay_interval, self.decay_r, epoch, init_r=self.init_r, final_r=self.final_r)
        self._step += 1
        if self.learn_edge_att:
            average = gi.symmetric                      # == is_undirected(edge_index), read once per batch
            att, edge_att, info_mean = ops.sample_avg_info(
                att_log_logits, training=training, rev=gi.rev, average=average, r=r, noise_u=noise_u, temp=1.0,
                info_on_edge_att=(self.info_on == 'edge_att'), seed=self.seed, offset=self._step * (1 << 32))
        else:
            att, _, info_att = ops.sample_avg_info(
                att_log_logits, training=training, rev=None, average=False, r=r, noise_u=noise_u, temp=1.0,
                info_on_edge_att=False, want_info=(self.info_on == 'att'), seed=self.seed,
                offset=self._step * (1 << 32))
            edge_att = ops.lift_node_att(att, gi)
            info_mean = info_att if self.info_on == 'att' else info_loss(edge_att, r)
        clf_logits = self.clf(data.x, data.edge_index, data.batch, edge_attr=data.edge_attr, edge_atten=edge_att)
        loss, loss_dict = self.__loss__(info_mean, clf_logits, data.y, epoch)
        return edge_att, loss, loss_dict, clf_logits

    @staticmethod
    def sampling(att_log_logit, training, noise_u=None):
        return concrete_sample(att_log_logit, 1, training, noise_u)

    get_r = staticmethod(get_r)
    lift_node_att_to_edge_att = staticmethod(lift_node_att_to_edge_att)
    concrete_sample = staticmethod(concrete_sample)
