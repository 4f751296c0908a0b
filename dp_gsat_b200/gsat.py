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
    return ops.gather_reverse(values, mp, involution=False)


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


class DualGSAT(tnn.Module):
    """The fork's GSAT class (src/run_gsat.py:33-149, 189-283, 610-637): a primal model and a dual (line-graph) model
    trained together by ``dual_forward_pass(primal_data, dual_data, epoch, training)`` ->
    ``(primal_edge_att, loss, loss_dict, primal_clf_logits)``.  Same attribute names (``primal_clf``,
    ``dual_extractor``, ``primal_learn_edge_att``, ``dual_final_r`` ...) and the same loss composition: primal / dual
    prediction losses, dual info loss against the scalar schedule r, primal info loss against the per-edge prior
    ``sigmoid(dual_logits).detach()``, f1 sparsity loss of the dual attention against the primal edge labels, and the
    0.3 / 0.7 mix of the dual attention into the primal edge attention after epoch 50.

    Every gather / scatter / sampling step is a kernel of this library (K0 index, K1 extractor, fused sampler +
    reverse-average + info loss, lift, K3 / K4 message passing, K5 pooling); the elementwise fork glue
    (gumbel_sigmoid, f1 loss, the mix) is plain device-side PyTorch composed through autograd.  Not carried over (SURVEY
    App. C): the per-batch ``.cpu().numpy()`` copies and plotting of :262-274 and the dead ``comb_att`` line that
    raises NameError when ``primal_learn_edge_att`` is set.  ``from_reference_args`` takes the reference constructor's
    28 positional arguments unchanged."""

    def __init__(self, primal_clf, primal_extractor, dual_clf, dual_extractor, primal_num_class, primal_multi_label,
                 dual_num_class, dual_multi_label, primal_method_config, primal_shared_config, dual_method_config,
                 dual_shared_config, primal_optimizer=None, dual_optimizer=None, lazy_metrics: bool = False, seed: int = 0):
        super().__init__()
        from .nn import Criterion
        self.primal_clf, self.primal_extractor, self.primal_optimizer = primal_clf, primal_extractor, primal_optimizer
        self.dual_clf, self.dual_extractor, self.dual_optimizer = dual_clf, dual_extractor, dual_optimizer
        self.primal_multi_label, self.dual_multi_label = primal_multi_label, dual_multi_label
        self.primal_criterion = Criterion(primal_num_class, primal_multi_label)
        self.dual_criterion = Criterion(dual_num_class, dual_multi_label)
        for side, mc, sc in (('primal', primal_method_config, primal_shared_config),
                             ('dual', dual_method_config, dual_shared_config)):
            setattr(self, f'{side}_method_name', mc.get('method_name', 'GSAT'))
            setattr(self, f'{side}_learn_edge_att', sc['learn_edge_att'])
            setattr(self, f'{side}_k', sc.get('precision_k', 5))
            setattr(self, f'{side}_epochs', mc.get('epochs', 0))
            setattr(self, f'{side}_pred_loss_coef', mc['pred_loss_coef'])
            setattr(self, f'{side}_info_loss_coef', mc['info_loss_coef'])
            setattr(self, f'{side}_fix_r', mc.get('fix_r', None))
            setattr(self, f'{side}_decay_interval', mc.get('decay_interval', None))
            setattr(self, f'{side}_decay_r', mc.get('decay_r', None))
            setattr(self, f'{side}_final_r', mc.get('final_r', 0.1))
            setattr(self, f'{side}_init_r', mc.get('init_r', 0.9))
        self.lazy_metrics, self.seed, self._step = lazy_metrics, seed, 0

    @classmethod
    def from_reference_args(cls, primal_clf, primal_extractor, primal_optimizer, primal_scheduler, primal_writer,
                            primal_device, primal_model_dir, primal_dataset_name, primal_num_class, primal_multi_label,
                            primal_random_state, primal_method_config, primal_shared_config, primal_model_config, dual_clf,
                            dual_extractor, dual_optimizer, dual_scheduler, dual_writer, dual_device, dual_model_dir,
                            dual_dataset_name, dual_num_class, dual_multi_label, dual_random_state, dual_method_config,
                            dual_shared_config, dual_model_config):
        """Argument order of the reference constructor (src/run_gsat.py:35-37); trainer-side objects are kept as
        attributes but unused by the step."""
        self = cls(primal_clf, primal_extractor, dual_clf, dual_extractor, primal_num_class, primal_multi_label,
                   dual_num_class, dual_multi_label, primal_method_config, primal_shared_config, dual_method_config,
                   dual_shared_config, primal_optimizer, dual_optimizer)
        for side, vals in (('primal', (primal_scheduler, primal_writer, primal_device, primal_model_dir,
                                       primal_dataset_name, primal_random_state, primal_model_config)),
                           ('dual', (dual_scheduler, dual_writer, dual_device, dual_model_dir, dual_dataset_name,
                                     dual_random_state, dual_model_config))):
            for name, v in zip(('scheduler', 'writer', 'device', 'model_dir', 'dataset_name', 'random_state',
                                'model_config'), vals):
                setattr(self, f'{side}_{name}', v)
        return self

    def __loss__(self, primal_info, dual_att, primal_clf_logits, dual_clf_logits, primal_clf_labels, dual_clf_labels,
                 epoch):
        """src/run_gsat.py:121-149; ``primal_info`` is the (already reduced) primal info loss against the per-edge prior."""
        primal_pred_loss = self.primal_criterion(primal_clf_logits, primal_clf_labels) * self.primal_pred_loss_coef
        dual_pred_loss = self.dual_criterion(dual_clf_logits, dual_clf_labels) * self.dual_pred_loss_coef
        dual_r = self.dual_fix_r if self.dual_fix_r else get_r(self.dual_decay_interval, self.dual_decay_r, epoch,
                                                               final_r=self.dual_final_r, init_r=self.dual_init_r)
        dual_info_loss = info_loss(dual_att, dual_r) * self.dual_info_loss_coef
        primal_info_loss = primal_info * self.primal_info_loss_coef
        loss = primal_pred_loss + dual_pred_loss + primal_info_loss + dual_info_loss
        if self.lazy_metrics:
            loss_dict = {'loss': loss.detach(), 'pred': dual_pred_loss.detach(), 'info': dual_info_loss.detach()}
        else:       # the reference's dict: the dual entries overwrite the primal ones (:144-147)
            loss_dict = {'loss': loss.item(), 'pred': dual_pred_loss.item(), 'info': dual_info_loss.item()}
        return loss, loss_dict

    def dual_forward_pass(self, primal_data, dual_data, epoch, training, noise=None):
        for clf in (self.primal_clf, self.dual_clf):
            clf._enc_scope = {}
        try:
            return self._dual_forward_pass(primal_data, dual_data, epoch, training, noise or {})
        finally:
            for clf in (self.primal_clf, self.dual_clf):
                clf._enc_scope = None

    def _dual_forward_pass(self, p, d, epoch, training, noise):
        gi_p = get_graph_index(p.edge_index, p.batch, getattr(p, 'num_graphs', None) or None)
        gi_d = get_graph_index(d.edge_index, d.batch, getattr(d, 'num_graphs', None) or None)
        self._step += 1
        off = self._step * (1 << 32)
        # dual side first: its logits are the prior of the primal info loss
        dual_emb = self.dual_clf.get_emb(d.x, d.edge_index, batch=d.batch, edge_attr=d.edge_attr)
        dual_att_log_logits = self.dual_extractor(dual_emb, d.edge_index, d.batch, 'dual')
        dual_node_att = gumbel_sigmoid(dual_att_log_logits, tau=0.1, noise_u=noise.get('dual_U'))[:, 0].unsqueeze(-1)
        f1_loss = f1_sparsity_loss(dual_node_att, p.edge_label.float().to(dual_node_att.device))       # :226-227
        if self.dual_learn_edge_att:
            dual_edge_att = ops.gather_reverse(dual_node_att, gi_d.rev).add(dual_node_att).div(2) if gi_d.symmetric \
                else dual_node_att                                                                     # :232-238
        else:
            dual_edge_att = ops.lift_node_att(dual_node_att, gi_d)                                     # :239-240
        # primal side
        primal_emb = self.primal_clf.get_emb(p.x, p.edge_index, batch=p.batch, edge_attr=p.edge_attr)
        primal_att_log_logits = self.primal_extractor(primal_emb, p.edge_index, p.batch, 'primal')
        primal_r = dual_att_log_logits.sigmoid().detach()                                              # :129
        mix = epoch > 50
        if self.primal_learn_edge_att:
            # sampling + reverse average (+ the info loss on the averaged attention when nothing is mixed in) in ONE kernel
            _, primal_edge_att, primal_info = ops.sample_avg_info(
                primal_att_log_logits, training=training, rev=gi_p.rev, average=gi_p.symmetric, r=primal_r,
                noise_u=noise.get('primal_u'), temp=1.0, info_on_edge_att=True, want_info=not mix, seed=self.seed,
                offset=off)
        else:
            primal_node_att, _, _ = ops.sample_avg_info(
                primal_att_log_logits, training=training, rev=None, average=False, noise_u=noise.get('primal_u'),
                temp=1.0, want_info=False, seed=self.seed, offset=off)
            primal_edge_att = ops.lift_node_att(primal_node_att, gi_p)                                 # :249
            primal_info = None
        if mix:                                                                                        # :252-253
            primal_edge_att = 0.3 * dual_node_att + (1 - 0.3) * primal_edge_att
            primal_info = None
        if primal_info is None:
            primal_info = info_loss(primal_edge_att, primal_r)                                         # :132
        primal_clf_logits = self.primal_clf(p.x, p.edge_index, p.batch, edge_attr=p.edge_attr, edge_atten=primal_edge_att)
        dual_clf_logits = self.dual_clf(d.x, d.edge_index, d.batch, edge_attr=d.edge_attr, edge_atten=dual_edge_att)
        loss, loss_dict = self.__loss__(primal_info, dual_edge_att, primal_clf_logits, dual_clf_logits, p.y, d.y, epoch)
        loss = loss + f1_loss                                                                          # :281
        return primal_edge_att, loss, loss_dict, primal_clf_logits

    def dual_eval_one_batch(self, primal_data, dual_data, epoch):
        """src/run_gsat.py:610-618 (outputs stay on the device: no per-batch .cpu())."""
        for m in (self.primal_extractor, self.primal_clf, self.dual_extractor, self.dual_clf):
            m.eval()
        with torch.no_grad():
            att, loss, loss_dict, clf_logits = self.dual_forward_pass(primal_data, dual_data, epoch, training=False)
        return att.detach().reshape(-1), loss_dict, clf_logits.detach()

    def dual_train_one_batch(self, primal_data, dual_data, epoch):
        """src/run_gsat.py:620-637."""
        for m in (self.primal_extractor, self.primal_clf, self.dual_extractor, self.dual_clf):
            m.train()
        att, loss, loss_dict, clf_logits = self.dual_forward_pass(primal_data, dual_data, epoch, training=True)
        self.primal_optimizer.zero_grad()
        self.dual_optimizer.zero_grad()
        loss.backward()
        self.primal_optimizer.step()
        self.dual_optimizer.step()
        return att.detach().reshape(-1), loss_dict, clf_logits.detach()

    @staticmethod
    def sampling(att_log_logits, epoch, training, noise_u=None):
        return concrete_sample(att_log_logits, 1, training, noise_u)

    get_r = staticmethod(get_r)
    gumbel_sigmoid = staticmethod(gumbel_sigmoid)
    f1_sparsity_loss = staticmethod(f1_sparsity_loss)
    lift_node_att_to_edge_att = staticmethod(lift_node_att_to_edge_att)
    concrete_sample = staticmethod(concrete_sample)
