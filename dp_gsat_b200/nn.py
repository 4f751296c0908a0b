"""Host-side mirror of the reference's model surfaces (same names, constructor arguments, forward signatures and
state_dict keys), backed by the sm_100a kernels of libgsat_b200.so.

  reference file                         -> here
  src/models/conv_layers.py:14-34  GINConv            -> GINConv
  src/models/gin.py:12-81          GIN                -> GIN
  src/models/conv_layers.py:37-92  GINEConv, LEConv   -> GINEConv, LEConv
  src/models/spmotif_gnn.py:9-87   SPMotifNet         -> SPMotifNet
  src/utils/get_model.py:7-68      get_model/Criterion/BatchSequential/MLP -> same names
  torch_geometric InstanceNorm / global_add_pool / global_mean_pool        -> InstanceNorm / ops.global_*_pool
  src/run_gsat.py:888-927, example/gsat.py:120-139  ExtractorMLP           -> ExtractorMLP

Dense Linear / BatchNorm1d layers run on this repo's tcgen05 GEMM and streaming kernels in both precision modes
(dense.py: 'fp32' = split-bf16 x3 strict mode, 'bf16' = single pass); every gather / scatter / segment / sampling op is a
kernel of this repo.  Dropout masks can be injected (``masks``) for parity tests; otherwise F.dropout.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as tnn
import torch.nn.functional as F

from . import dense, ops
from .dense import Linear, PrecisionMixin
from .index import GraphIndex, get_graph_index


def _dropout(x, p: float, training: bool, masks=None, key: str = ''):
    if not training or p == 0.0:
        return x
    if masks is None:
        return F.dropout(x, p, True)
    m = masks.get(key, x.shape, p).to(device=x.device, dtype=x.dtype)
    return x * m / (1.0 - p)


class InstanceNorm(tnn.Module):
    """torch_geometric.nn.InstanceNorm(C) with its defaults (eps 1e-5, affine False, no running stats), applied
    per graph over contiguous row segments.  ``batch`` may be the reference's int64 segment-id vector (then the
    segment pointers come from the cached GraphIndex passed as ``seg``) or a prepared (seg_ptr, num_segments)."""

    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps

    def forward(self, x, batch, seg=None):
        if seg is None:
            seg = segments_from_batch(batch)
        seg_ptr, G = seg
        return ops.segment_instance_norm(x, seg_ptr, G, self.eps)


def segments_from_batch(batch: torch.Tensor):
    """(seg_ptr int32 [G+1], G) from a non-decreasing int64 segment-id vector, through K0."""
    dummy = _EMPTY_EDGES.get(batch.device)
    if dummy is None:
        dummy = torch.zeros((2, 0), dtype=torch.int64, device=batch.device)
        _EMPTY_EDGES[batch.device] = dummy
    gi = get_graph_index(dummy, batch)
    gi.require_graph_contiguous()
    return gi.node_ptr, gi.G


_EMPTY_EDGES = {}


class BatchSequential(tnn.Sequential):
    """src/utils/get_model.py:47-54."""

    def forward(self, inputs, batch, seg=None, masks=None, key: str = 'ext'):
        li = 0
        for module in self._modules.values():
            if isinstance(module, InstanceNorm):
                inputs = module(inputs, batch, seg)
            elif isinstance(module, tnn.Dropout):
                inputs = _dropout(inputs, module.p, self.training, masks, f'{key}.{li}')
                li += 1
            else:
                inputs = module(inputs)
        return inputs


class MLP(BatchSequential):
    """src/utils/get_model.py:57-68 (state_dict keys '0','4','8' as in the reference Sequential)."""

    def __init__(self, channels: Sequence[int], dropout: float, bias: bool = True):
        m = []
        for i in range(1, len(channels)):
            m.append(Linear(channels[i - 1], channels[i], bias))
            if i < len(channels) - 1:
                m.append(InstanceNorm(channels[i]))
                m.append(tnn.ReLU())
                m.append(tnn.Dropout(dropout))
        super().__init__(*m)


class Criterion(tnn.Module):
    """src/utils/get_model.py:19-34."""

    def __init__(self, num_class, multi_label):
        super().__init__()
        self.num_class, self.multi_label = num_class, multi_label

    def forward(self, logits, targets):
        if self.num_class == 2 and not self.multi_label:
            return F.binary_cross_entropy_with_logits(logits, targets.float())
        elif self.num_class > 2 and not self.multi_label:
            return F.cross_entropy(logits, targets.long())
        is_labeled = targets == targets
        return F.binary_cross_entropy_with_logits(logits[is_labeled], targets[is_labeled].float())


def get_preds(logits, multi_label):
    """src/utils/get_model.py:37-44."""
    if multi_label:
        return (logits.sigmoid() > 0.5).float()
    if logits.shape[1] > 1:
        return logits.argmax(dim=1).float()
    return (logits.sigmoid() > 0.5).float()


class _SyncBatchNormFn(torch.autograd.Function):
    """BatchNorm1d over the rows of ALL ranks of a data-parallel group (the one cross-shard coupling
    of the step).  Two-pass statistics (sum -> mean, then centred sum of squares), each one small all-reduce; backward
    all-reduces (sum dy, sum dy * xhat).  The parameter gradients returned are the LOCAL sums: the step's gradient
    all-reduce adds them up like every other parameter gradient."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, momentum, eps, group):
        import torch.distributed as dist
        C = x.shape[1]
        buf = torch.empty(C + 1, dtype=torch.float64, device=x.device)
        buf[:C] = x.sum(0, dtype=torch.float64)
        buf[C] = x.shape[0]
        dist.all_reduce(buf, group=group)
        n = buf[C].clone()
        mean = (buf[:C] / n)
        xc = x - mean.to(x.dtype)
        sq = xc.square().sum(0, dtype=torch.float64)
        dist.all_reduce(sq, group=group)
        var = sq / n                                           # biased, as F.batch_norm normalises in training mode
        rstd = torch.rsqrt(var + eps).to(x.dtype)
        xhat = xc * rstd
        with torch.no_grad():
            running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
            running_var.mul_(1 - momentum).add_((var * (n / (n - 1).clamp_min(1))).to(running_var.dtype), alpha=momentum)
            nbt.add_(1)
        ctx.save_for_backward(xhat, gamma, rstd, n)
        ctx.group = group
        return xhat * gamma + beta

    @staticmethod
    def backward(ctx, dy):
        import torch.distributed as dist
        xhat, gamma, rstd, n = ctx.saved_tensors
        C = xhat.shape[1]
        local = torch.empty(2 * C, dtype=torch.float64, device=dy.device)
        local[:C] = dy.sum(0, dtype=torch.float64)
        local[C:] = (dy * xhat).sum(0, dtype=torch.float64)
        glob = local.clone()
        dist.all_reduce(glob, group=ctx.group)
        m_dy, m_dyx = (glob[:C] / n).to(dy.dtype), (glob[C:] / n).to(dy.dtype)
        dx = (gamma * rstd) * (dy - m_dy - xhat * m_dyx)
        return dx, local[C:].to(gamma.dtype), local[:C].to(gamma.dtype), None, None, None, None, None, None


class BatchNorm1d(tnn.BatchNorm1d):
    """torch.nn.BatchNorm1d as the reference builds it (gin.py:59; pna.py:45 through PyG BatchNorm: eps 1e-5, momentum
    0.1, affine, running statistics -- same parameters, buffers and state_dict keys), plus an optional data-parallel
    group: with ``sync_group`` set (parallel.enable_sync_batchnorm) the TRAINING-mode batch statistics span the rows of
    every rank, which makes the graph-sharded N-GPU step compute exactly the single-device step;
    ``sync_group = None`` (default) keeps shard-local statistics, the semantics of torch DDP."""
    sync_group = None

    def forward(self, x):
        if self.sync_group is None or not self.training:
            training = self.training or self.running_mean is None
            if training and self.momentum is None:
                raise NotImplementedError('BatchNorm1d(momentum=None) (cumulative average) is not used by the reference')
            if training and self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
            return dense.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, training,
                                    self.momentum if self.momentum is not None else 0.1, self.eps)
        return _SyncBatchNormFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var,
                                      self.num_batches_tracked, self.momentum if self.momentum is not None else 0.1,
                                      self.eps, self.sync_group)


class BatchNorm(tnn.Module):
    """torch_geometric.nn.BatchNorm as the reference's PNA uses it (pna.py:8,45): a thin wrapper holding the
    BatchNorm1d as ``self.module``, so that a PNA state_dict has the reference's ``batch_norms.{i}.module.*`` keys."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class GINConv(tnn.Module):
    """src/models/conv_layers.py:14-34 over torch_geometric GINConv(nn, eps=0., train_eps=False): ``eps`` is a
    buffer, present in the state_dict."""

    def __init__(self, nn: tnn.Module, eps: float = 0.0, train_eps: bool = False):
        super().__init__()
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            raise NotImplementedError('train_eps=True is never used by the reference (gin.py:40)')
        self.register_buffer('eps', torch.tensor([eps]))

    def forward(self, x, edge_index, edge_attr=None, edge_atten=None, size=None, _index: Optional[GraphIndex] = None):
        gi = _index if _index is not None else get_graph_index(edge_index, None, num_nodes=x.shape[0])
        out = ops.gin_aggregate(x, edge_atten, gi, self.initial_eps)
        return self.nn(out)


class GINEConv(tnn.Module):
    """src/models/conv_layers.py:37-66 over torch_geometric GINEConv(nn, eps=0., train_eps=False, edge_dim): message
    relu(x_j + lin(edge_attr)) * edge_atten; ``lin`` = Linear(edge_dim, in_channels) when edge_dim is given; ``eps`` is
    a buffer, present in the state_dict (keys: nn.*, eps, lin.weight, lin.bias)."""

    def __init__(self, nn: tnn.Module, eps: float = 0.0, train_eps: bool = False, edge_dim: Optional[int] = None):
        super().__init__()
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            raise NotImplementedError('train_eps=True is never used by the reference (gin.py:38)')
        self.register_buffer('eps', torch.tensor([eps]))
        if edge_dim is not None:
            first = nn[0] if isinstance(nn, tnn.Sequential) else nn
            in_channels = first.in_features if hasattr(first, 'in_features') else first.in_channels
            self.lin = Linear(edge_dim, in_channels)
        else:
            self.lin = None

    def forward(self, x, edge_index, edge_attr=None, edge_atten=None, size=None, _index: Optional[GraphIndex] = None):
        gi = _index if _index is not None else get_graph_index(edge_index, None, num_nodes=x.shape[0])
        if self.lin is None and x.size(-1) != edge_attr.size(-1):
            raise ValueError("Node and edge feature dimensionalities do not match. Consider setting the 'edge_dim' "
                             "attribute of 'GINEConv'")
        ef = self.lin(edge_attr) if self.lin is not None else edge_attr
        out = ops.gine_aggregate(x, ef, edge_atten, gi, self.initial_eps)
        return self.nn(out)


class LEConv(tnn.Module):
    """src/models/conv_layers.py:69-92 over torch_geometric 2.0.3 LEConv(in_channels, out_channels, bias=True):
    lin1 (bias), lin2 (no bias), lin3 (bias) -- the same state_dict keys; message (a_j - b_i) * edge_weight * edge_atten
    summed over incoming edges, plus lin3(x).  The three Linears are library GEMMs; the message passing, the root
    term and their backward are kernels of csrc/leconv.cu."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin1 = Linear(in_channels, out_channels, bias=bias)
        self.lin2 = Linear(in_channels, out_channels, bias=False)
        self.lin3 = Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, edge_index, edge_weight=None, edge_atten=None, _index: Optional[GraphIndex] = None):
        gi = _index if _index is not None else get_graph_index(edge_index, None, num_nodes=x.shape[0])
        return ops.le_aggregate(self.lin1(x), self.lin2(x), edge_weight, edge_atten, gi, add=self.lin3(x))


class SPMotifNet(PrecisionMixin, tnn.Module):
    """src/models/spmotif_gnn.py:9-87 (same attribute names, state_dict keys and method signatures)."""

    def __init__(self, x_dim, edge_attr_dim, num_class, multi_label, model_config):
        super().__init__()
        self.n_layers = model_config['n_layers']
        hidden_size = model_config['hidden_size']
        self.edge_attr_dim = edge_attr_dim
        self.node_emb = Linear(x_dim, hidden_size)
        self.convs = tnn.ModuleList()
        self.relus = tnn.ModuleList()
        for _ in range(self.n_layers):
            self.convs.append(LEConv(in_channels=hidden_size, out_channels=hidden_size))
            self.relus.append(tnn.ReLU())
        self.fc_out = tnn.Sequential(Linear(hidden_size, 2 * hidden_size), tnn.ReLU(),
                                     Linear(2 * hidden_size, num_class))
        self.conf_mlp = tnn.Sequential(Linear(hidden_size, 2 * hidden_size), tnn.ReLU(),
                                       Linear(2 * hidden_size, 3))
        self.cq = Linear(3, 3)
        self.conf_fw = tnn.Sequential(self.conf_mlp, self.cq)

    def pool(self, x, batch, _index: Optional[GraphIndex] = None):
        gi = _index if _index is not None else get_graph_index(_no_edges(batch.device), batch)
        return ops.global_mean_pool(x, gi)

    def forward(self, x, edge_index, batch, edge_attr, edge_atten=None):
        node_x = self.get_node_reps(x, edge_index, edge_attr, batch, edge_atten=edge_atten)
        return self.get_causal_pred(self.pool(node_x, batch, get_graph_index(edge_index, batch)))

    def get_emb(self, x, edge_index, batch, edge_attr, edge_atten=None):
        return self.get_node_reps(x, edge_index, edge_attr, batch, edge_atten=edge_atten)

    def get_pred_from_emb(self, emb, batch):
        return self.fc_out(self.pool(emb, batch))

    def get_node_reps(self, x, edge_index, edge_attr, batch, edge_atten):
        gi = get_graph_index(edge_index, batch)
        x = _encode_once(self, x, 'node_emb')
        for conv, relu in zip(self.convs, self.relus):
            x = relu(conv(x=x, edge_index=edge_index, edge_weight=edge_attr, edge_atten=edge_atten, _index=gi))
        return x

    def get_graph_rep(self, x, edge_index, edge_attr, batch, edge_atten):
        node_x = self.get_node_reps(x, edge_index, edge_attr, batch, edge_atten=edge_atten)
        return self.pool(node_x, batch, get_graph_index(edge_index, batch))

    def get_causal_pred(self, causal_graph_x):
        return self.fc_out(causal_graph_x)

    def get_conf_pred(self, conf_graph_x):
        return self.conf_fw(conf_graph_x)

    def get_comb_pred(self, causal_graph_x, conf_graph_x):
        causal_pred = self.fc_out(causal_graph_x)
        conf_pred = self.conf_mlp(conf_graph_x).detach()
        return torch.sigmoid(conf_pred) * causal_pred

    def reset_parameters(self):
        with torch.no_grad():
            for param in self.parameters():
                param.uniform_(-1.0, 1.0)


class GIN(PrecisionMixin, tnn.Module):
    """src/models/gin.py:12-81."""

    def __init__(self, x_dim, edge_attr_dim, num_class, multi_label, model_config):
        super().__init__()
        self.n_layers = model_config['n_layers']
        hidden_size = model_config['hidden_size']
        self.edge_attr_dim = edge_attr_dim
        self.dropout_p = model_config['dropout_p']
        self.use_edge_attr = model_config.get('use_edge_attr', True)
        self.with_edges = edge_attr_dim != 0 and self.use_edge_attr
        if model_config.get('atom_encoder', False):
            self.node_encoder = AtomEncoder(emb_dim=hidden_size)
            if self.with_edges:
                self.edge_encoder = BondEncoder(emb_dim=hidden_size)
        else:
            self.node_encoder = Linear(x_dim, hidden_size)
            if self.with_edges:
                self.edge_encoder = Linear(edge_attr_dim, hidden_size)
        self.convs = tnn.ModuleList()
        self.relu = tnn.ReLU()
        for _ in range(self.n_layers):
            if self.with_edges:      # gin.py:36-38
                self.convs.append(GINEConv(GIN.MLP(hidden_size, hidden_size), edge_dim=hidden_size))
            else:
                self.convs.append(GINConv(GIN.MLP(hidden_size, hidden_size)))
        self.fc_out = tnn.Sequential(Linear(hidden_size, 1 if num_class == 2 and not multi_label else num_class))
        self.masks = None     # parity tests inject dropout masks here
        self.precision = 'fp32'   # 'bf16': fused node MLPs on tcgen05 (tc.gin_layer); 'fp32': strict split-bf16 x3 path
        self.seed = 0
        self._calls = 0

    @staticmethod
    def MLP(in_channels: int, out_channels: int):
        return tnn.Sequential(Linear(in_channels, out_channels), BatchNorm1d(out_channels),
                              tnn.ReLU(inplace=True), Linear(out_channels, out_channels))

    def pool(self, x, batch, _index: Optional[GraphIndex] = None):
        gi = _index if _index is not None else get_graph_index(_no_edges(batch.device), batch)
        return ops.global_add_pool(x, gi)

    def get_emb(self, x, edge_index, batch, edge_attr=None, edge_atten=None, mask_key: str = 'gin'):
        gi = get_graph_index(edge_index, batch)
        x = _encode_once(self, x)
        if edge_attr is not None and self.use_edge_attr and self.with_edges:
            edge_attr = self.edge_encoder(edge_attr)          # gin.py:46-47, 66-67
        else:
            edge_attr = None
        fused = self.precision == 'bf16' and x.shape[1] % 8 == 0 and x.shape[1] <= 128 and not self.with_edges
        for i in range(self.n_layers):
            if fused:
                # K3 aggregation (bf16 out) chained into the node MLP + ReLU + dropout on the tensor cores
                from . import tc
                dm = None
                if self.masks is not None and self.training and self.dropout_p > 0:
                    dm = self.masks.get(f'{mask_key}.{i}', (x.shape[0], self.convs[i].nn[3].weight.shape[0]),
                                        self.dropout_p).to(device=x.device, dtype=torch.uint8)
                self._calls += 1
                x = tc.gin_layer(x, edge_atten, gi, self.convs[i], self.training, self.dropout_p,
                                 self.seed * 7919 + self._calls, dm)         # ReLU + dropout fused in the epilogue
                continue
            x = self.convs[i](x, edge_index, edge_attr=edge_attr, edge_atten=edge_atten, _index=gi)
            x = self.relu(x)
            x = _dropout(x, self.dropout_p, self.training, self.masks, f'{mask_key}.{i}')
        return x

    def forward(self, x, edge_index, batch, edge_attr=None, edge_atten=None, mask_key: str = 'gin.clf'):
        gi = get_graph_index(edge_index, batch)
        x = self.get_emb(x, edge_index, batch, edge_attr, edge_atten, mask_key=mask_key)
        return self.fc_out(self.pool(x, batch, gi))

    def get_graph_emb(self, x, edge_index, batch, edge_attr=None, edge_atten=None):
        return self.pool(self.get_emb(x, edge_index, batch, edge_attr, edge_atten), batch)

    def get_pred_from_emb(self, emb, batch):
        return self.fc_out(self.pool(emb, batch))


def _encode_once(model, x, attr: str = 'node_encoder'):
    """node_encoder(x), shared between the two GNN passes of ONE GSAT.forward_pass (get_emb, then clf: same input,
    same weights, no dropout in front of it -- example/gsat.py:75,86), so the encoder GEMM and its weight-gradient
    GEMM run once per step instead of twice.  GSAT.forward_pass opens / closes the scope (``_enc_scope``); outside
    of it every call encodes afresh."""
    scope = getattr(model, '_enc_scope', None)
    if scope is None:
        return _encode(model, x, attr)
    key = (x.data_ptr(), tuple(x.shape), x._version, torch.is_grad_enabled())
    if scope.get('key') != key:
        scope['key'], scope['out'] = key, _encode(model, x, attr)
    return scope['out']


def _encode(model, x, attr: str = 'node_encoder'):
    enc = getattr(model, attr)
    if isinstance(enc, tnn.Linear) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 \
            and x.shape[1] < 16 and enc.out_features % 4 == 0:
        return ops.small_linear(x, enc.weight, enc.bias)      # own kernel for the K = N weight-gradient reduction
    return enc(x)


def _no_edges(device):
    d = _EMPTY_EDGES.get(device)
    if d is None:
        d = torch.zeros((2, 0), dtype=torch.int64, device=device)
        _EMPTY_EDGES[device] = d
    return d


ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_FEATURE_DIMS = [5, 6, 2]


class _SumEmbeddingEncoder(tnn.Module):
    """Sum of one embedding table per integer feature column (ogb 1.3.2 AtomEncoder / BondEncoder).

    ``fused = True`` (default) runs the whole encoder as ONE gather-sum kernel forward and one deterministic backward into
    the tables (csrc/encoders.cu) instead of K embedding lookups + K-1 adds and K scatter-adds
    (sorts + atomics) backward; results are bit-identical forward (same addition order).  ``fused = False`` keeps the library
    embedding lookups."""
    fused = True
    _list_name = ''

    def _tables(self):
        return [emb.weight for emb in getattr(self, self._list_name)]

    def forward(self, x):
        if self.fused:
            return ops.embedding_sum(x, self._tables(), getattr(self, 'oob_flag', None))
        out = 0
        for i, emb in enumerate(getattr(self, self._list_name)):
            out = out + emb(x[:, i])
        return out


class AtomEncoder(_SumEmbeddingEncoder):
    """ogb 1.3.2 AtomEncoder: sum of 9 embedding tables (state_dict keys atom_embedding_list.{k}.weight)."""
    _list_name = 'atom_embedding_list'

    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = tnn.ModuleList()
        for d in ATOM_FEATURE_DIMS:
            emb = tnn.Embedding(d, emb_dim)
            tnn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)


class BondEncoder(_SumEmbeddingEncoder):
    """ogb 1.3.2 BondEncoder: sum of 3 embedding tables (state_dict keys bond_embedding_list.{k}.weight)."""
    _list_name = 'bond_embedding_list'

    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = tnn.ModuleList()
        for d in BOND_FEATURE_DIMS:
            emb = tnn.Embedding(d, emb_dim)
            tnn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)


class ExtractorMLP(PrecisionMixin, tnn.Module):
    """Upstream form  ExtractorMLP(hidden_size, learn_edge_att).forward(emb, edge_index, batch)
    (example/gsat.py:120-139) and fork form  ExtractorMLP(hidden_size, shared_config, type).forward(emb, edge_index,
    batch, type)  (src/run_gsat.py:888-927; parameters live under '<type>_feature_extractor')."""

    def __init__(self, hidden_size, shared_config, type: Optional[str] = None):
        super().__init__()
        if isinstance(shared_config, bool):
            shared_config = {'learn_edge_att': shared_config, 'extractor_dropout_p': 0.5}
        self.learn_edge_att = shared_config['learn_edge_att']
        dropout_p = shared_config['extractor_dropout_p']
        self.kind = type
        self._name = 'feature_extractor' if type is None else f'{type}_feature_extractor'
        if type is not None:
            setattr(self, f'{type}_learn_edge_att', self.learn_edge_att)
        if self.learn_edge_att:
            mlp = MLP([hidden_size * 2, hidden_size * 4, hidden_size, 1], dropout=dropout_p)
        else:
            mlp = MLP([hidden_size * 1, hidden_size * 2, hidden_size, 1], dropout=dropout_p)
        setattr(self, self._name, mlp)
        self.masks = None
        self.chunk_rows = 1 << 21      # rows per graph-aligned chunk (0 = never chunk)
        self.precision = 'fp32'        # 'fp32': strict path (split-bf16 x3 tcgen05 GEMMs + segment-norm kernels, rtol 1e-5
        #                                parity); 'bf16': fused tcgen05 kernels (tc.fused_extractor_v2), bf16 tolerance
        self.seed = 0
        self._calls = 0

    def forward(self, emb, edge_index, batch, type: Optional[str] = None):
        if type is not None and type != self.kind:
            return None      # the reference falls through and returns None for an unknown type
        mlp = getattr(self, self._name)
        gi = get_graph_index(edge_index, batch)
        gi.require_graph_contiguous()
        from . import tc
        if self.precision == 'bf16' and tc.fused_extractor_supported(emb, gi, self.learn_edge_att):
            # K1 of the design: the whole MLP as one persistent tcgen05 kernel per direction (bf16 operands).  Batches it
            # cannot tile -- a graph with more rows than one accumulator tile (mutag-dual: up to 406 dual edges), or
            # hidden_size > 128 / not a multiple of 8 (the H = 300 sweep) -- take the layer-by-layer path below in the
            # SAME precision mode and on the same tensor-core GEMM kernels (dense.Linear + the segment-norm kernels).
            lin = [m for m in mlp if isinstance(m, tnn.Linear)]
            p = next(m.p for m in mlp if isinstance(m, tnn.Dropout))
            rows_n = gi.E if self.learn_edge_att else gi.N
            m1 = m2 = None
            if self.masks is not None and self.training and p > 0:
                m1 = self.masks.get('ext.0', (rows_n, lin[0].weight.shape[0]), p).to(device=emb.device, dtype=torch.uint8)
                m2 = self.masks.get('ext.1', (rows_n, lin[1].weight.shape[0]), p).to(device=emb.device, dtype=torch.uint8)
            self._calls += 1
            return tc.fused_extractor_v2(emb, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight,
                                         lin[2].bias, gi, edge_mode=self.learn_edge_att, pdrop=p, training=self.training,
                                         seed=self.seed * 1000003 + self._calls, mask1=m1, mask2=m2)
        if self.learn_edge_att:
            f12 = ops.gather_concat(emb, gi)      # cat(emb[col], emb[row]) with col, row = edge_index
            rows, seg_all = f12, (gi.edge_ptr, gi.G)
        else:
            rows, seg_all = emb, (gi.node_ptr, gi.G)
        if self.chunk_rows and rows.shape[0] > self.chunk_rows and self.masks is None:
            # InstanceNorm is per graph, so graph-aligned row ranges are independent: run them one after the other
            # and recompute each range in backward, which bounds the [rows, 4H] activations held at any time
            from torch.utils.checkpoint import checkpoint
            outs = []
            for r0, r1, seg_ptr, ng in gi.chunk_plan(self.chunk_rows, 'edge' if self.learn_edge_att else 'node'):
                fn = lambda t, sp=seg_ptr, n=ng: mlp(t, None, seg=(sp, n))
                outs.append(checkpoint(fn, rows[r0:r1], use_reentrant=False) if torch.is_grad_enabled()
                            else fn(rows[r0:r1]))
            return torch.cat(outs, dim=0)
        return mlp(rows, None, seg=seg_all, masks=self.masks)


def get_model(x_dim, edge_attr_dim, num_class, multi_label, model_config, device):
    """src/utils/get_model.py:7-16."""
    if model_config['model_name'] == 'GIN':
        model = GIN(x_dim, edge_attr_dim, num_class, multi_label, model_config)
    elif model_config['model_name'] == 'PNA':
        from .pna import PNA
        model = PNA(x_dim, edge_attr_dim, num_class, multi_label, model_config)
    elif model_config['model_name'] == 'SPMotifNet':
        model = SPMotifNet(x_dim, edge_attr_dim, num_class, multi_label, model_config)
    else:
        raise ValueError('[ERROR] Unknown model name!')
    return model.to(device)
