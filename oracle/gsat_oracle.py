"""CPU oracle for the GSAT stochastic-attention message-passing path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``dp_gsat_b200``)
may import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference (mihikamd/DP-GSAT) ships no tests, golden
vectors or fixtures for this path, and its arithmetic lives in wheels that are
absent from this image (torch-geometric 2.0.3, torch-scatter 2.0.9,
torch-sparse 0.6.12, ogb 1.3.2 -- pins from reference README.md:50,66-68 and
requirements.txt:1), so the reference itself cannot be imported to generate
vectors.  This file restates the published algorithms of those wheels
 and the reference's own call sites in plain PyTorch,
dtype-generic (fp32 / fp64), with injectable noise and dropout masks.  The
only reference-side pins available are (1) the ``reorder_like`` runtime
invariant (src/utils/utils.py:23), (2) the in-tree Mutagenicity topology
(data/mutag_dual/raw/Mutagenicity_A.txt, a slice of which is committed under
tests/golden/ by tests/golden/make_golden.py) whose consecutive rows are
mutual reverses, (3) analytic known answers (tests/test_oracle.py), and (4) the
outputs of the reference's own first-party function / class bodies, extracted
with ``ast`` and executed by tests/golden/make_golden.py and
make_golden_fork.py (reorder_like, sampler, losses, MLP / ExtractorMLP, the
conv layers' forward + message, SPMotifNet, the fork's __loss__ and
dual_forward_pass) -- first-party code is pinned, the third-party wheels'
semantics are not.

Every public symbol cites the reference file:line it follows (paths relative
to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------
# torch_scatter 2.0.9 / torch_geometric.utils restatements
# --------------------------------------------------------------------------


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter.scatter(src, index, 0, None, dim_size, 'sum') == index_add in edge order."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def degree(index: torch.Tensor, num_nodes: int, dtype=torch.float32) -> torch.Tensor:
    """torch_geometric.utils.degree: scatter-sum of ones."""
    out = torch.zeros(num_nodes, dtype=dtype, device=index.device)
    return out.index_add_(0, index, torch.ones(index.numel(), dtype=dtype, device=index.device))


def scatter_mean(src, index, dim_size):
    """sum / count.clamp(min=1) (torch_scatter 2.0.9 scatter_mean)."""
    s = scatter_sum(src, index, dim_size)
    cnt = degree(index, dim_size, dtype=src.dtype).clamp_(min=1)
    return s / cnt.view(-1, *([1] * (src.dim() - 1)))


def _scatter_minmax(src, index, dim_size, is_max: bool):
    """torch_scatter scatter_min / scatter_max: empty segments -> 0, arg = first occurrence on ties (CPU)."""
    big = torch.finfo(src.dtype).max
    fill = -big if is_max else big
    out = torch.full((dim_size,) + tuple(src.shape[1:]), fill, dtype=src.dtype, device=src.device)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    out = out.scatter_reduce(0, idx, src.detach(), reduce='amax' if is_max else 'amin', include_self=True)
    # arg element: smallest edge id attaining the extremum (gradient flows only there)
    E = src.shape[0]
    hit = src.detach() == out.index_select(0, index)
    eid = torch.arange(E, device=src.device).view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    cand = torch.where(hit, eid, torch.full_like(eid, E))
    arg = torch.full(out.shape, E, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(0, idx, cand, reduce='amin', include_self=True)
    has = arg < E
    gathered = torch.gather(src, 0, arg.clamp(max=max(E - 1, 0))) if E > 0 else torch.zeros_like(out)
    return torch.where(has, gathered, torch.zeros_like(out)), arg


def scatter_min(src, index, dim_size):
    return _scatter_minmax(src, index, dim_size, False)[0]


def scatter_max(src, index, dim_size):
    return _scatter_minmax(src, index, dim_size, True)[0]


def global_add_pool(x, batch, num_graphs: Optional[int] = None):
    """PyG global_add_pool = scatter(x, batch, dim_size=batch.max()+1, 'add') (gin.py:34,53)."""
    G = int(batch.max()) + 1 if num_graphs is None else num_graphs
    return scatter_sum(x, batch, G)


def global_mean_pool(x, batch, num_graphs: Optional[int] = None):
    """PyG global_mean_pool (pna.py:47,62)."""
    G = int(batch.max()) + 1 if num_graphs is None else num_graphs
    return scatter_mean(x, batch, G)


def sort_edge_index(edge_index, edge_attr=None, num_nodes: Optional[int] = None):
    """PyG 2.0.3 sort_edge_index: perm = (row*num_nodes+col).argsort(). Stable here so that
    duplicate edges have a defined order (the reference's argsort leaves ties unordered)."""
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    perm = torch.argsort(edge_index[0] * n + edge_index[1], stable=True)
    ei = edge_index[:, perm]
    return ei, (None if edge_attr is None else edge_attr[perm])


def transpose(index, value, m=None, n=None, coalesced: bool = False):
    """torch_sparse.transpose(..., coalesced=False): pure row swap (run_gsat.py:243)."""
    assert not coalesced
    row, col = index[0], index[1]
    return torch.stack([col, row], dim=0), value


def is_undirected(edge_index) -> bool:
    """PyG is_undirected: the edge (multi)set equals its transpose (run_gsat.py:242)."""
    if edge_index.numel() == 0:
        return True
    n = int(edge_index.max()) + 1
    k1 = torch.sort(edge_index[0] * n + edge_index[1]).values
    k2 = torch.sort(edge_index[1] * n + edge_index[0]).values
    return bool((k1 == k2).all())


def reorder_like(from_edge_index, to_edge_index, values):
    """src/utils/utils.py:19-25, restated line for line (stable argsorts)."""
    from_edge_index, values = sort_edge_index(from_edge_index, values)
    ranking_score = to_edge_index[0] * (to_edge_index.max() + 1) + to_edge_index[1]
    ranking = ranking_score.argsort(stable=True).argsort(stable=True)
    if not (from_edge_index[:, ranking] == to_edge_index).all():
        raise ValueError("Edges in from_edge_index and to_edge_index are different, impossible to match both.")
    return values[ranking]


def reverse_edge_permutation(edge_index) -> torch.Tensor:
    """rev[e] such that reorder_like(transpose(ei, v), ei, v) == v[rev]  (utils.py:19-25 composed with
    run_gsat.py:243): the edge at rank k of the stable (dst,src) order is matched to the edge at rank k of the
    stable (src,dst) order."""
    E = edge_index.shape[1]
    ar = torch.arange(E, dtype=torch.long)
    t_idx, t_val = transpose(edge_index, ar)
    return reorder_like(t_idx, edge_index, t_val)


def build_index_oracle(edge_index: torch.Tensor, batch: torch.Tensor, num_graphs: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Bit-exact specification of K0 (index builder): CSR-by-dst, CSC-by-src, reverse map, graph segment
    pointers and flags, all int32.  Canonical orders: stable ascending (dst, src) and stable ascending
    (src, dst); these are the two argsorts reorder_like performs (utils.py:20-22)."""
    src, dst = edge_index[0].long(), edge_index[1].long()
    N = batch.numel()
    E = src.numel()
    G = (int(batch.max()) + 1 if N > 0 else 0) if num_graphs is None else num_graphs
    by_dst = torch.argsort(dst * max(N, 1) + src, stable=True)
    by_src = torch.argsort(src * max(N, 1) + dst, stable=True)
    rowptr_dst = torch.zeros(N + 1, dtype=torch.long)
    rowptr_dst[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0)
    rowptr_src = torch.zeros(N + 1, dtype=torch.long)
    rowptr_src[1:] = torch.cumsum(torch.bincount(src, minlength=N), 0)
    rev = torch.full((E,), -1, dtype=torch.long)
    ok = (src[by_src] == dst[by_dst]) & (dst[by_src] == src[by_dst])
    rev[by_src[ok]] = by_dst[ok]
    symmetric = bool(ok.all())
    key_sorted = (src * max(N, 1) + dst)[by_src]
    has_dup = bool((key_sorted[1:] == key_sorted[:-1]).any()) if E > 1 else False
    node_ptr = torch.zeros(G + 1, dtype=torch.long)
    node_ptr[1:] = torch.cumsum(torch.bincount(batch, minlength=G), 0)
    eb = batch[src] if E > 0 else torch.zeros(0, dtype=torch.long)
    edge_ptr = torch.zeros(G + 1, dtype=torch.long)
    edge_ptr[1:] = torch.cumsum(torch.bincount(eb, minlength=G), 0)
    nodes_sorted = bool((batch[1:] >= batch[:-1]).all()) if N > 1 else True
    edges_sorted = bool((eb[1:] >= eb[:-1]).all()) if E > 1 else True
    same_graph = bool((batch[src] == batch[dst]).all()) if E > 0 else True
    i32 = lambda t: t.to(torch.int32)
    return {
        'src': i32(src), 'dst': i32(dst),
        'rev': i32(rev),
        'rowptr_dst': i32(rowptr_dst), 'eid_by_dst': i32(by_dst), 'src_by_dst': i32(src[by_dst]),
        'rowptr_src': i32(rowptr_src), 'eid_by_src': i32(by_src), 'dst_by_src': i32(dst[by_src]),
        'node_ptr': i32(node_ptr), 'edge_ptr': i32(edge_ptr), 'edge_graph': i32(eb),
        'symmetric': symmetric, 'has_dup': has_dup,
        'graph_contiguous': nodes_sorted and edges_sorted and same_graph,
    }


# --------------------------------------------------------------------------
# dropout with injectable masks
# --------------------------------------------------------------------------


class MaskSource:
    """Deterministic keep-masks shared between the oracle and the CUDA path in parity tests.  ``get`` returns a
    float mask of 0/1 for the given site key; masks are created on CPU from a seeded generator the first time a
    (key, shape) is requested and replayed afterwards, so two models asking for the same sites see the same masks."""

    def __init__(self, seed: int = 2):
        self.seed = seed
        self.store: Dict[str, torch.Tensor] = {}

    def get(self, key: str, shape, p: float) -> torch.Tensor:
        k = f"{key}:{tuple(shape)}:{p}"
        if k not in self.store:
            g = torch.Generator().manual_seed(self.seed + (hash_str(k) % 100003))
            self.store[k] = (torch.rand(tuple(shape), generator=g) >= p).to(torch.float32)
        return self.store[k]


def hash_str(s: str) -> int:
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def dropout(x, p: float, training: bool, masks: Optional[MaskSource] = None, key: str = ''):
    """F.dropout semantics: kept values scaled by 1/(1-p).  With ``masks`` the keep-mask is injected."""
    if not training or p == 0.0:
        return x
    if masks is None:
        return F.dropout(x, p, True)
    m = masks.get(key, x.shape, p).to(dtype=x.dtype, device=x.device)
    return x * m / (1.0 - p)


# --------------------------------------------------------------------------
# norms / MLP  (src/utils/get_model.py:47-68, PyG InstanceNorm -- SURVEY A.3)
# --------------------------------------------------------------------------


class InstanceNorm(nn.Module):
    """PyG 2.0.3 InstanceNorm(C) defaults: eps=1e-5, affine=False, track_running_stats=False."""

    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.channels = channels
        self.eps = eps

    def forward(self, x, batch, num_graphs: Optional[int] = None):
        G = int(batch.max()) + 1 if num_graphs is None else num_graphs
        cnt = degree(batch, G, dtype=x.dtype).clamp_(min=1).view(-1, 1)
        mean = scatter_sum(x, batch, G) / cnt
        x = x - mean.index_select(0, batch)
        var = scatter_sum(x * x, batch, G) / cnt
        return x / (var + self.eps).sqrt().index_select(0, batch)


class MLP(nn.Module):
    """get_model.py:57-68: Linear -> InstanceNorm -> ReLU -> Dropout for every non-last layer, Linear last.
    Module indices (0,4,8) match the reference's nn.Sequential so state_dict keys are identical."""

    def __init__(self, channels: Sequence[int], dropout: float, bias: bool = True):
        super().__init__()
        self.p = dropout
        self.channels = list(channels)
        idx = 0
        self._lin_ids: List[int] = []
        for i in range(1, len(channels)):
            self.add_module(str(idx), nn.Linear(channels[i - 1], channels[i], bias))
            self._lin_ids.append(idx)
            idx += 1
            if i < len(channels) - 1:
                self.add_module(str(idx), InstanceNorm(channels[i]))
                idx += 3  # InstanceNorm, ReLU, Dropout occupy three slots in the reference Sequential

    def forward(self, inputs, batch, masks: Optional[MaskSource] = None, key: str = 'ext'):
        x = inputs
        n = len(self._lin_ids)
        for li, mid in enumerate(self._lin_ids):
            x = getattr(self, str(mid))(x)
            if li < n - 1:
                x = getattr(self, str(mid + 1))(x, batch)
                x = F.relu(x)
                x = dropout(x, self.p, self.training, masks, f'{key}.{li}')
        return x


class Criterion(nn.Module):
    """get_model.py:19-34."""

    def __init__(self, num_class, multi_label):
        super().__init__()
        self.num_class = num_class
        self.multi_label = multi_label

    def forward(self, logits, targets):
        if self.num_class == 2 and not self.multi_label:
            return F.binary_cross_entropy_with_logits(logits, targets.to(logits.dtype))
        elif self.num_class > 2 and not self.multi_label:
            return F.cross_entropy(logits, targets.long())
        is_labeled = targets == targets
        return F.binary_cross_entropy_with_logits(logits[is_labeled], targets[is_labeled].to(logits.dtype))


# --------------------------------------------------------------------------
# conv layers (src/models/conv_layers.py)
# --------------------------------------------------------------------------


class GINConv(nn.Module):
    """conv_layers.py:14-34 on top of PyG GINConv(nn, eps=0, train_eps=False): eps is a buffer."""

    def __init__(self, mlp: nn.Module, eps: float = 0.0):
        super().__init__()
        self.nn = mlp
        self.register_buffer('eps', torch.tensor([eps]))

    def forward(self, x, edge_index, edge_attr=None, edge_atten=None, size=None):
        x_j = x.index_select(0, edge_index[0])                     # propagate/__collect__
        msg = x_j * edge_atten if edge_atten is not None else x_j   # message :29-34
        out = scatter_sum(msg, edge_index[1], x.shape[0])           # aggregate (aggr='add')
        out = out + (1 + self.eps.to(x.dtype)) * x                  # :23-25
        return self.nn(out)


class GINEConv(nn.Module):
    """conv_layers.py:37-66 on top of PyG GINEConv(nn, eps=0, train_eps=False, edge_dim): lin = Linear(edge_dim,
    in_channels) when edge_dim is given; message (x_j + lin(edge_attr)).relu() * edge_atten; aggr 'add'."""

    def __init__(self, mlp: nn.Module, eps: float = 0.0, edge_dim: Optional[int] = None):
        super().__init__()
        self.nn = mlp
        self.register_buffer('eps', torch.tensor([eps]))
        in_channels = mlp[0].in_features if isinstance(mlp, nn.Sequential) else getattr(mlp, 'in_features', None)
        self.lin = nn.Linear(edge_dim, in_channels) if edge_dim is not None else None

    def forward(self, x, edge_index, edge_attr=None, edge_atten=None, size=None):
        x_j = x.index_select(0, edge_index[0])
        if self.lin is None and x_j.size(-1) != edge_attr.size(-1):
            raise ValueError("Node and edge feature dimensionalities do not match. Consider setting the 'edge_dim' "
                             "attribute of 'GINEConv'")
        ea = self.lin(edge_attr) if self.lin is not None else edge_attr            # :57-58
        m = (x_j + ea).relu()                                                       # :59
        msg = m * edge_atten if edge_atten is not None else m                       # :61-64
        out = scatter_sum(msg, edge_index[1], x.shape[0])
        out = out + (1 + self.eps.to(x.dtype)) * x                                  # :46-48
        return self.nn(out)


class LEConv(nn.Module):
    """conv_layers.py:69-92 on top of PyG 2.0.3 LEConv(in_channels, out_channels, bias=True): lin1 = Linear(in, out,
    bias), lin2 = Linear(in, out, bias=False), lin3 = Linear(in, out, bias); aggr 'add'; message
    (a_j - b_i) [* edge_weight.view(-1, 1)] [* edge_atten]; out = propagate(...) + lin3(x)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin1 = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin2 = nn.Linear(in_channels, out_channels, bias=False)
        self.lin3 = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, edge_index, edge_weight=None, edge_atten=None):
        a, b = self.lin1(x), self.lin2(x)                                           # :76-77
        out = a.index_select(0, edge_index[0]) - b.index_select(0, edge_index[1])   # a_j - b_i, :86
        m = out if edge_weight is None else out * edge_weight.view(-1, 1)           # :87
        if edge_atten is not None:                                                  # :89-92
            m = m * edge_atten
        return scatter_sum(m, edge_index[1], x.shape[0]) + self.lin3(x)             # :80-82


def gin_mlp(in_channels: int, out_channels: int) -> nn.Sequential:
    """gin.py:55-62."""
    return nn.Sequential(nn.Linear(in_channels, out_channels), nn.BatchNorm1d(out_channels),
                         nn.ReLU(inplace=False), nn.Linear(out_channels, out_channels))


def aggregate_var(src, index, dim_size):
    mean = scatter_mean(src, index, dim_size)
    mean_squares = scatter_mean(src * src, index, dim_size)
    return mean_squares - mean * mean


def aggregate_std(src, index, dim_size):
    return torch.sqrt(torch.relu(aggregate_var(src, index, dim_size)) + 1e-5)


AGGREGATORS = {'sum': scatter_sum, 'mean': scatter_mean, 'min': scatter_min, 'max': scatter_max,
               'var': aggregate_var, 'std': aggregate_std}  # conv_layers.py:193-226


def _scale_identity(src, deg, avg):
    return src


def _scale_amplification(src, deg, avg):
    return src * (torch.log(deg + 1) / avg['log'])


def _scale_attenuation(src, deg, avg):
    scale = avg['log'] / torch.log(deg + 1)
    scale[deg == 0] = 1
    return src * scale


def _scale_linear(src, deg, avg):
    return src * (deg / avg['lin'])


def _scale_inverse_linear(src, deg, avg):
    scale = avg['lin'] / deg
    scale[deg == 0] = 1
    return src * scale


SCALERS = {'identity': _scale_identity, 'amplification': _scale_amplification, 'attenuation': _scale_attenuation,
           'linear': _scale_linear, 'inverse_linear': _scale_inverse_linear}  # conv_layers.py:229-259


class PNAConvSimple(nn.Module):
    """conv_layers.py:96-190."""

    def __init__(self, in_channels, out_channels, aggregators, scalers, deg, post_layers: int = 1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.aggregator_names = list(aggregators)
        self.scaler_names = list(scalers)
        degf = deg.to(torch.float)
        self.avg_deg = {'lin': degf.mean().item(), 'log': (degf + 1).log().mean().item(),
                        'exp': degf.exp().mean().item()}
        mods: List[nn.Module] = [nn.Linear(len(aggregators) * len(scalers) * in_channels, out_channels)]
        for _ in range(post_layers - 1):
            mods += [nn.ReLU(), nn.Linear(out_channels, out_channels)]
        self.post_nn = nn.Sequential(*mods)

    def forward(self, x, edge_index, edge_attr=None, edge_atten=None):
        src, dst = edge_index[0], edge_index[1]
        x_i, x_j = x.index_select(0, dst), x.index_select(0, src)
        m = torch.cat([x_i, x_j] + ([edge_attr] if edge_attr is not None else []), dim=-1)   # :166-171
        if edge_atten is not None:
            m = m * edge_atten
        N = x.shape[0]
        out = torch.cat([AGGREGATORS[a](m, dst, N) for a in self.aggregator_names], dim=-1)   # :178-181
        deg = degree(dst, N, dtype=m.dtype).view(-1, 1)
        out = torch.cat([SCALERS[s](out, deg, self.avg_deg) for s in self.scaler_names], dim=-1)
        return self.post_nn(out)


# --------------------------------------------------------------------------
# encoders (ogb 1.3.2)
# --------------------------------------------------------------------------

ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_FEATURE_DIMS = [5, 6, 2]


class _SumEmbedding(nn.Module):
    def __init__(self, dims, emb_dim, attr):
        super().__init__()
        lst = nn.ModuleList()
        for d in dims:
            emb = nn.Embedding(d, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            lst.append(emb)
        setattr(self, attr, lst)
        self._attr = attr

    def forward(self, x):
        out = 0
        for i, emb in enumerate(getattr(self, self._attr)):
            out = out + emb(x[:, i])
        return out


class AtomEncoder(_SumEmbedding):
    def __init__(self, emb_dim):
        super().__init__(ATOM_FEATURE_DIMS, emb_dim, 'atom_embedding_list')


class BondEncoder(_SumEmbedding):
    def __init__(self, emb_dim):
        super().__init__(BOND_FEATURE_DIMS, emb_dim, 'bond_embedding_list')


# --------------------------------------------------------------------------
# backbones (src/models/gin.py, src/models/pna.py)
# --------------------------------------------------------------------------


class GIN(nn.Module):
    """gin.py:12-81 (GINConv, or GINEConv when edge_attr_dim != 0 and use_edge_attr)."""

    def __init__(self, x_dim, edge_attr_dim, num_class, multi_label, model_config):
        super().__init__()
        self.n_layers = model_config['n_layers']
        hidden = model_config['hidden_size']
        self.edge_attr_dim = edge_attr_dim
        self.dropout_p = model_config['dropout_p']
        self.use_edge_attr = model_config.get('use_edge_attr', True)
        self.with_edges = edge_attr_dim != 0 and self.use_edge_attr
        if model_config.get('atom_encoder', False):
            self.node_encoder = AtomEncoder(hidden)
            if self.with_edges:
                self.edge_encoder = BondEncoder(hidden)
        else:
            self.node_encoder = nn.Linear(x_dim, hidden)
            if self.with_edges:
                self.edge_encoder = nn.Linear(edge_attr_dim, hidden)
        if self.with_edges:                      # gin.py:36-38
            self.convs = nn.ModuleList([GINEConv(gin_mlp(hidden, hidden), edge_dim=hidden) for _ in range(self.n_layers)])
        else:
            self.convs = nn.ModuleList([GINConv(gin_mlp(hidden, hidden)) for _ in range(self.n_layers)])
        self.fc_out = nn.Sequential(nn.Linear(hidden, 1 if num_class == 2 and not multi_label else num_class))
        self.masks: Optional[MaskSource] = None
        self._pass = 0

    def get_emb(self, x, edge_index, batch, edge_attr=None, edge_atten=None, mask_key: str = 'gin'):
        x = self.node_encoder(x)
        edge_attr = self.edge_encoder(edge_attr) if (edge_attr is not None and self.with_edges) else None   # gin.py:66-67
        for i in range(self.n_layers):
            x = self.convs[i](x, edge_index, edge_attr=edge_attr, edge_atten=edge_atten)
            x = F.relu(x)
            x = dropout(x, self.dropout_p, self.training, self.masks, f'{mask_key}.{i}')
        return x

    def forward(self, x, edge_index, batch, edge_attr=None, edge_atten=None, mask_key: str = 'gin.clf'):
        x = self.get_emb(x, edge_index, batch, edge_attr, edge_atten, mask_key=mask_key)
        return self.fc_out(global_add_pool(x, batch))

    def get_pred_from_emb(self, emb, batch):
        return self.fc_out(global_add_pool(emb, batch))


class BatchNorm(nn.Module):
    """torch_geometric.nn.BatchNorm (2.0.3; pna.py:8,45): a wrapper that holds ``torch.nn.BatchNorm1d(C, eps=1e-5,
    momentum=0.1, affine=True, track_running_stats=True)`` as ``self.module`` -- hence the ``batch_norms.{i}.module.*``
    keys of a reference PNA checkpoint (from the published source, unpinned)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class PNA(nn.Module):
    """pna.py:12-78."""

    def __init__(self, x_dim, edge_attr_dim, num_class, multi_label, model_config):
        super().__init__()
        hidden = model_config['hidden_size']
        self.n_layers = model_config['n_layers']
        self.dropout_p = model_config['dropout_p']
        self.edge_attr_dim = edge_attr_dim
        use_ea = model_config.get('use_edge_attr', True)
        if model_config.get('atom_encoder', False):
            self.node_encoder = AtomEncoder(hidden)
            if edge_attr_dim != 0 and use_ea:
                self.edge_encoder = BondEncoder(hidden)
        else:
            self.node_encoder = nn.Linear(x_dim, hidden)
            if edge_attr_dim != 0 and use_ea:
                self.edge_encoder = nn.Linear(edge_attr_dim, hidden)
        aggregators = model_config['aggregators']
        scalers = ['identity', 'amplification', 'attenuation'] if model_config['scalers'] else ['identity']
        deg = model_config['deg']
        in_channels = (hidden * 2 if edge_attr_dim == 0 else hidden * 3) if use_ea else hidden * 2
        self.convs = nn.ModuleList()
        self.batch_norms = nn.ModuleList()
        for _ in range(self.n_layers):
            self.convs.append(PNAConvSimple(in_channels, hidden, aggregators, scalers, deg, post_layers=1))
            self.batch_norms.append(BatchNorm(hidden))
        self.fc_out = nn.Sequential(nn.Linear(hidden, hidden // 2), nn.ReLU(),
                                    nn.Linear(hidden // 2, hidden // 4), nn.ReLU(),
                                    nn.Linear(hidden // 4, 1 if num_class == 2 and not multi_label else num_class))
        self.masks: Optional[MaskSource] = None

    def get_emb(self, x, edge_index, batch, edge_attr, edge_atten=None, mask_key: str = 'pna'):
        x = self.node_encoder(x)
        if edge_attr is not None:
            edge_attr = self.edge_encoder(edge_attr)
        for i, (conv, bn) in enumerate(zip(self.convs, self.batch_norms)):
            h = F.relu(bn(conv(x, edge_index, edge_attr, edge_atten=edge_atten)))
            x = h + x
            x = dropout(x, self.dropout_p, self.training, self.masks, f'{mask_key}.{i}')
        return x

    def forward(self, x, edge_index, batch, edge_attr, edge_atten=None, mask_key: str = 'pna.clf'):
        x = self.get_emb(x, edge_index, batch, edge_attr, edge_atten, mask_key=mask_key)
        return self.fc_out(global_mean_pool(x, batch))

    def get_pred_from_emb(self, emb, batch):
        return self.fc_out(global_mean_pool(emb, batch))


class SPMotifNet(nn.Module):
    """spmotif_gnn.py:9-87 (LEConv backbone of "Discovering Invariant Rationales"; global_mean_pool readout)."""

    def __init__(self, x_dim, edge_attr_dim, num_class, multi_label, model_config):
        super().__init__()
        self.n_layers = model_config['n_layers']
        hidden = model_config['hidden_size']
        self.edge_attr_dim = edge_attr_dim
        self.node_emb = nn.Linear(x_dim, hidden)
        self.convs = nn.ModuleList([LEConv(hidden, hidden) for _ in range(self.n_layers)])
        self.relus = nn.ModuleList([nn.ReLU() for _ in range(self.n_layers)])
        self.fc_out = nn.Sequential(nn.Linear(hidden, 2 * hidden), nn.ReLU(), nn.Linear(2 * hidden, num_class))
        self.conf_mlp = nn.Sequential(nn.Linear(hidden, 2 * hidden), nn.ReLU(), nn.Linear(2 * hidden, 3))
        self.cq = nn.Linear(3, 3)
        self.conf_fw = nn.Sequential(self.conf_mlp, self.cq)

    def get_node_reps(self, x, edge_index, edge_attr, batch, edge_atten):
        x = self.node_emb(x)
        for conv, relu in zip(self.convs, self.relus):
            x = relu(conv(x, edge_index, edge_weight=edge_attr, edge_atten=edge_atten))      # :60-62
        return x

    def get_emb(self, x, edge_index, batch, edge_attr, edge_atten=None):
        return self.get_node_reps(x, edge_index, edge_attr, batch, edge_atten=edge_atten)

    def forward(self, x, edge_index, batch, edge_attr, edge_atten=None):
        return self.fc_out(global_mean_pool(self.get_node_reps(x, edge_index, edge_attr, batch, edge_atten), batch))

    def get_pred_from_emb(self, emb, batch):
        return self.fc_out(global_mean_pool(emb, batch))

    def get_graph_rep(self, x, edge_index, edge_attr, batch, edge_atten):
        return global_mean_pool(self.get_node_reps(x, edge_index, edge_attr, batch, edge_atten), batch)

    def get_causal_pred(self, causal_graph_x):
        return self.fc_out(causal_graph_x)

    def get_conf_pred(self, conf_graph_x):
        return self.conf_fw(conf_graph_x)

    def get_comb_pred(self, causal_graph_x, conf_graph_x):
        return torch.sigmoid(self.conf_mlp(conf_graph_x).detach()) * self.fc_out(causal_graph_x)


def get_model(x_dim, edge_attr_dim, num_class, multi_label, model_config, device='cpu'):
    """get_model.py:7-16."""
    if model_config['model_name'] == 'GIN':
        model = GIN(x_dim, edge_attr_dim, num_class, multi_label, model_config)
    elif model_config['model_name'] == 'PNA':
        model = PNA(x_dim, edge_attr_dim, num_class, multi_label, model_config)
    elif model_config['model_name'] == 'SPMotifNet':
        model = SPMotifNet(x_dim, edge_attr_dim, num_class, multi_label, model_config)
    else:
        raise ValueError('[ERROR] Unknown model name!')
    return model.to(device)


# --------------------------------------------------------------------------
# extractor + GSAT step (example/gsat.py:27-139, src/run_gsat.py:121-149,182-187,860-927)
# --------------------------------------------------------------------------


class ExtractorMLP(nn.Module):
    """example/gsat.py:120-139 (3-arg forward) and run_gsat.py:888-927 (4-arg forward with ``type``)."""

    def __init__(self, hidden_size, shared_config, type: Optional[str] = None):
        super().__init__()
        if isinstance(shared_config, bool):   # upstream ctor: ExtractorMLP(hidden_size, learn_edge_att)
            shared_config = {'learn_edge_att': shared_config, 'extractor_dropout_p': 0.5}
        self.learn_edge_att = shared_config['learn_edge_att']
        p = shared_config['extractor_dropout_p']
        self.kind = type
        name = 'feature_extractor' if type is None else f'{type}_feature_extractor'
        chans = [hidden_size * 2, hidden_size * 4, hidden_size, 1] if self.learn_edge_att \
            else [hidden_size, hidden_size * 2, hidden_size, 1]
        setattr(self, name, MLP(chans, dropout=p))
        self._name = name
        self.masks: Optional[MaskSource] = None

    def forward(self, emb, edge_index, batch, type: Optional[str] = None):
        if type is not None and type != self.kind:
            return None   # run_gsat.py:909-927 falls through for an unknown type
        mlp = getattr(self, self._name)
        if self.learn_edge_att:
            col, row = edge_index[0], edge_index[1]
            f12 = torch.cat([emb[col], emb[row]], dim=-1)
            return mlp(f12, batch[col], self.masks)
        return mlp(emb, batch, self.masks)


def get_r(decay_interval, decay_r, current_epoch, init_r=0.9, final_r=0.5):
    """run_gsat.py:860-864 / example/gsat.py:105-110."""
    r = init_r - current_epoch // decay_interval * decay_r
    if r < final_r:
        r = final_r
    return r


def concrete_sample(att_log_logit, temp=1, training=True, noise_u: Optional[torch.Tensor] = None):
    """run_gsat.py:877-885; ``noise_u`` injects the uniform draw (already in [1e-10, 1-1e-10])."""
    if training:
        if noise_u is None:
            noise_u = torch.empty_like(att_log_logit).uniform_(1e-10, 1 - 1e-10)
        rn = torch.log(noise_u) - torch.log(1.0 - noise_u)
        return ((att_log_logit + rn) / temp).sigmoid()
    return att_log_logit.sigmoid()


def gumbel_sigmoid(logits, tau=1.0, eps=1e-10, noise_u: Optional[torch.Tensor] = None):
    """run_gsat.py:182-187."""
    U = torch.rand_like(logits) if noise_u is None else noise_u
    g = -torch.log(-torch.log(U + eps) + eps)
    return torch.sigmoid((logits + g) / tau)


def lift_node_att_to_edge_att(node_att, edge_index):
    """run_gsat.py:870-875."""
    return node_att[edge_index[0]] * node_att[edge_index[1]]


def info_loss(att, r):
    """example/gsat.py:31 / run_gsat.py:127,132; ``r`` scalar or per-edge tensor."""
    return (att * torch.log(att / r + 1e-6) + (1 - att) * torch.log((1 - att) / (1 - r + 1e-6) + 1e-6)).mean()


def f1_sparsity_loss(p_uv, y_uv, eps=1e-6):
    """run_gsat.py:151-180."""
    TP = (p_uv.view(-1) * y_uv.view(-1)).sum()
    P, G = p_uv.sum(), y_uv.sum()
    precision, recall = TP / (P + eps), TP / (G + eps)
    f1 = 2 * precision * recall / (precision + recall + eps)
    return (1 - f1) + p_uv.abs().mean()


def undirected_average(att, edge_index):
    """run_gsat.py:241-247 / example/gsat.py:79-85."""
    if is_undirected(edge_index):
        trans_idx, trans_val = transpose(edge_index, att, None, None, coalesced=False)
        trans_val_perm = reorder_like(trans_idx, edge_index, trans_val)
        return (att + trans_val_perm) / 2
    return att


class GSAT(nn.Module):
    """Canonical single-graph step, example/gsat.py:12-117.  ``info_on='att'`` is upstream (loss on the pre-average
    attention, :91); ``info_on='edge_att'`` is the fork (run_gsat.py:276)."""

    def __init__(self, clf, extractor, criterion, optimizer=None, learn_edge_att=True, final_r=0.7,
                 decay_interval=10, decay_r=0.1, init_r=0.9, info_on: str = 'att',
                 pred_loss_coef=1.0, info_loss_coef=1.0):
        super().__init__()
        self.clf, self.extractor, self.criterion, self.optimizer = clf, extractor, criterion, optimizer
        self.learn_edge_att = learn_edge_att
        self.final_r, self.decay_interval, self.decay_r, self.init_r = final_r, decay_interval, decay_r, init_r
        self.info_on = info_on
        self.pred_loss_coef, self.info_loss_coef = pred_loss_coef, info_loss_coef
        self.pred_scale = self.info_scale = 1.0   # data-parallel shard weights (G_local/G_global, E_local/E_global)

    def __loss__(self, att, clf_logits, clf_labels, epoch, r=None):
        pred_loss = self.criterion(clf_logits, clf_labels) * (self.pred_loss_coef * self.pred_scale)
        if r is None:
            r = get_r(self.decay_interval, self.decay_r, epoch, init_r=self.init_r, final_r=self.final_r)
        il = info_loss(att, r) * (self.info_loss_coef * self.info_scale)
        loss = pred_loss + il
        return loss, {'loss': loss.item(), 'pred': pred_loss.item(), 'info': il.item()}

    def forward_pass(self, data, epoch, training, noise_u=None, r=None):
        emb = self.clf.get_emb(data.x, data.edge_index, batch=data.batch, edge_attr=data.edge_attr)
        att_log_logits = self.extractor(emb, data.edge_index, data.batch)
        att = concrete_sample(att_log_logits, 1, training, noise_u)
        if self.learn_edge_att:
            edge_att = undirected_average(att, data.edge_index)
        else:
            edge_att = lift_node_att_to_edge_att(att, data.edge_index)
        clf_logits = self.clf(data.x, data.edge_index, data.batch, edge_attr=data.edge_attr, edge_atten=edge_att)
        loss, loss_dict = self.__loss__(att if self.info_on == 'att' else edge_att, clf_logits, data.y, epoch, r)
        return edge_att, loss, loss_dict, clf_logits

    sampling = staticmethod(lambda logits, training, noise_u=None: concrete_sample(logits, 1, training, noise_u))
    get_r = staticmethod(get_r)
    lift_node_att_to_edge_att = staticmethod(lift_node_att_to_edge_att)


class DualGSAT(nn.Module):
    """The fork's two-model step, run_gsat.py:33-149 (constructor fields, ``__loss__``) and :189-283
    (``dual_forward_pass``): a primal GSAT and a dual (line-graph) GSAT trained together.  Dual nodes are primal
    edges, so the dual "node attention" (one value per dual node, gumbel_sigmoid at tau 0.1) has one entry per primal
    edge: it feeds the f1 loss against the primal edge labels (:226-227), the per-edge prior r of the primal info
    loss (``primal_r = sigmoid(dual_logits).detach()``, :129) and, after epoch 50, the 0.3 / 0.7 mix into the primal
    edge attention (:252-253).  Not restated: the numpy / plotting code of :262-274 and the dead ``comb_att`` line
    (:270; it reads ``old_primal_edge_att``, which does not exist when primal_learn_edge_att is set -- SURVEY App. C).
    ``noise``: optional dict of injected uniforms {'primal_u', 'dual_U'}."""

    def __init__(self, primal_clf, primal_extractor, dual_clf, dual_extractor, primal_criterion, dual_criterion,
                 primal_method_config, primal_shared_config, dual_method_config, dual_shared_config):
        super().__init__()
        self.primal_clf, self.primal_extractor = primal_clf, primal_extractor
        self.dual_clf, self.dual_extractor = dual_clf, dual_extractor
        self.primal_criterion, self.dual_criterion = primal_criterion, dual_criterion
        for side, mc, sc in (('primal', primal_method_config, primal_shared_config),
                             ('dual', dual_method_config, dual_shared_config)):
            setattr(self, f'{side}_learn_edge_att', sc['learn_edge_att'])
            setattr(self, f'{side}_pred_loss_coef', mc['pred_loss_coef'])
            setattr(self, f'{side}_info_loss_coef', mc['info_loss_coef'])
            setattr(self, f'{side}_fix_r', mc.get('fix_r', None))
            setattr(self, f'{side}_decay_interval', mc.get('decay_interval', None))
            setattr(self, f'{side}_decay_r', mc.get('decay_r', None))
            setattr(self, f'{side}_final_r', mc.get('final_r', 0.1))
            setattr(self, f'{side}_init_r', mc.get('init_r', 0.9))

    def __loss__(self, primal_att, dual_att, primal_clf_logits, dual_clf_logits, primal_clf_labels, dual_clf_labels,
                 dual_att_log_logits, epoch):
        primal_pred_loss = self.primal_criterion(primal_clf_logits, primal_clf_labels)           # :122
        dual_pred_loss = self.dual_criterion(dual_clf_logits, dual_clf_labels)                   # :124
        dual_r = self.dual_fix_r if self.dual_fix_r else get_r(self.dual_decay_interval, self.dual_decay_r, epoch,
                                                               final_r=self.dual_final_r, init_r=self.dual_init_r)
        dual_info_loss = info_loss(dual_att, dual_r)                                             # :127
        primal_r = dual_att_log_logits.sigmoid().detach()                                        # :129
        primal_info_loss = info_loss(primal_att, primal_r)                                       # :132
        primal_pred_loss = primal_pred_loss * self.primal_pred_loss_coef
        primal_info_loss = primal_info_loss * self.primal_info_loss_coef
        dual_pred_loss = dual_pred_loss * self.dual_pred_loss_coef
        dual_info_loss = dual_info_loss * self.dual_info_loss_coef
        loss = primal_pred_loss + dual_pred_loss + primal_info_loss + dual_info_loss             # :142
        loss_dict = {'loss': loss.item(), 'pred': primal_pred_loss.item(), 'info': primal_info_loss.item()}
        loss_dict.update({'loss': loss.item(), 'pred': dual_pred_loss.item(), 'info': dual_info_loss.item()})
        return loss, loss_dict

    def dual_forward_pass(self, primal_data, dual_data, epoch, training, noise=None):
        noise = noise or {}
        p, d = primal_data, dual_data
        primal_emb = self.primal_clf.get_emb(p.x, p.edge_index, batch=p.batch, edge_attr=p.edge_attr)
        primal_att_log_logits = self.primal_extractor(primal_emb, p.edge_index, p.batch, 'primal')
        primal_node_att = concrete_sample(primal_att_log_logits, 1, training, noise.get('primal_u'))     # :204
        dual_emb = self.dual_clf.get_emb(d.x, d.edge_index, batch=d.batch, edge_attr=d.edge_attr)
        dual_att_log_logits = self.dual_extractor(dual_emb, d.edge_index, d.batch, 'dual')
        dual_node_att = gumbel_sigmoid(dual_att_log_logits, tau=0.1, noise_u=noise.get('dual_U'))[:, 0].unsqueeze(-1)
        f1_loss = f1_sparsity_loss(dual_node_att, p.edge_label.float())                                   # :226-227
        if self.dual_learn_edge_att:
            dual_edge_att = undirected_average(dual_node_att, d.edge_index)
        else:
            dual_edge_att = lift_node_att_to_edge_att(dual_node_att, d.edge_index)
        if self.primal_learn_edge_att:
            primal_edge_att = undirected_average(primal_node_att, p.edge_index)
        else:
            primal_edge_att = lift_node_att_to_edge_att(primal_node_att, p.edge_index)
        if epoch > 50:                                                                                     # :252-253
            primal_edge_att = 0.3 * dual_node_att + (1 - 0.3) * primal_edge_att
        primal_clf_logits = self.primal_clf(p.x, p.edge_index, p.batch, edge_attr=p.edge_attr, edge_atten=primal_edge_att)
        dual_clf_logits = self.dual_clf(d.x, d.edge_index, d.batch, edge_attr=d.edge_attr, edge_atten=dual_edge_att)
        loss, loss_dict = self.__loss__(primal_edge_att, dual_edge_att, primal_clf_logits, dual_clf_logits, p.y, d.y,
                                        dual_att_log_logits, epoch)
        loss = loss + f1_loss                                                                              # :281
        return primal_edge_att, loss, loss_dict, primal_clf_logits


# --------------------------------------------------------------------------
# line-graph ("dual") construction of the fork
# --------------------------------------------------------------------------


def line_graph_dual(edge_index: torch.Tensor, batch: torch.Tensor, halve: bool = False):
    """Restatement of reference src/datasets/mutag_dual.py:342-378 (``group_by_first`` + ``add_pairs_from_group``),
    with dual nodes identified by their index in the primal edge list, and -- ``halve`` -- of the relabelling of
    :536-548 (``dual_dict = {pair: idx // 2 + 1}``: consecutive rows, the two directions of a primal edge, share one
    id; 0-based here).  Plain Python loops on purpose: small cases only.

    Returns (dual_edge_index int64 [2, E_d], dual_batch int64 [E] or [E/2])."""
    src = edge_index[0].tolist()
    E = len(src)
    group_by_first = {}                      # dict insertion order == order of first appearance (mutag_dual.py:351-353)
    for idx, a in enumerate(src):
        group_by_first.setdefault(a, []).append(idx)
    dual_edges = []
    for group in group_by_first.values():    # mutag_dual.py:374-376
        if len(group) > 1:
            for i in range(len(group)):      # add_pairs_from_group, mutag_dual.py:363-372
                for j in range(i + 1, len(group)):
                    e1, e2 = group[i], group[j]
                    if halve:
                        e1, e2 = e1 // 2, e2 // 2
                    dual_edges.append([e1, e2])
                    dual_edges.append([e2, e1])
    dual_ei = torch.tensor(dual_edges, dtype=torch.int64).t().contiguous() if dual_edges \
        else torch.zeros((2, 0), dtype=torch.int64)
    node_of_dual = edge_index[0][::2] if halve else edge_index[0]
    return dual_ei, batch[node_of_dual].clone()


# --------------------------------------------------------------------------
# per-batch explanation metrics of the trainer
# --------------------------------------------------------------------------


def get_precision_at_k(att, exp_labels, k, batch, edge_index):
    """src/run_gsat.py:783-791, verbatim structure (loop over graphs, boolean edge masks, argsort of -att); the
    argsort is made stable (kind='stable') so that ties are defined: the smaller edge index wins."""
    import numpy as np
    att = att.detach().reshape(-1).cpu()
    exp_labels = exp_labels.detach().reshape(-1).cpu()
    precision_at_k = []
    for i in range(int(batch.max()) + 1):
        nodes_for_graph_i = batch == i
        edges_for_graph_i = nodes_for_graph_i[edge_index[0]] & nodes_for_graph_i[edge_index[1]]
        labels_for_graph_i = exp_labels[edges_for_graph_i]
        mask_log_logits_for_graph_i = att[edges_for_graph_i]
        top = np.argsort(-mask_log_logits_for_graph_i.numpy(), kind='stable')[:k]
        precision_at_k.append(labels_for_graph_i[top].sum().item() / k)
    return precision_at_k


def get_delta_kl(exp_labels, att, eps=1e-6):
    """src/run_gsat.py:793-800."""
    p = exp_labels.float().clamp(min=eps, max=1 - eps)
    r_uv = att.clamp(min=eps, max=1 - eps)
    r = r_uv.mean().clamp(min=eps, max=1 - eps)
    delta_kl = p * torch.log(r_uv / r) + (1 - p) * torch.log((1 - r_uv) / (1 - r))
    return delta_kl.sum().item()


# --------------------------------------------------------------------------
# batch collate of the reference's loaders
# --------------------------------------------------------------------------


def collate_data_list(graphs):
    """torch_geometric 2.0.3 ``Batch.from_data_list`` as the reference's DataLoaders apply it
    (src/utils/get_data_loaders.py:130-145): every per-node / per-edge / per-graph tensor concatenated along dim 0,
    ``edge_index`` concatenated along dim 1 after adding the cumulative node count of the preceding graphs, ``batch`` =
    index of the owning graph for every node.  ``graphs``: objects with x, edge_index (graph-local ids), y and optional
    edge_attr / edge_label / node_label.  Returns a dict of tensors."""
    xs, eis, batch, ys, eas, els, nls = [], [], [], [], [], [], []
    offset = 0
    for i, g in enumerate(graphs):
        n = g.x.shape[0]
        xs.append(g.x)
        eis.append(g.edge_index + offset)                       # __inc__('edge_index') = num_nodes
        batch.append(torch.full((n,), i, dtype=torch.long))
        ys.append(g.y)
        eas.append(getattr(g, 'edge_attr', None))
        els.append(getattr(g, 'edge_label', None))
        nls.append(getattr(g, 'node_label', None))
        offset += n
    opt = lambda parts: None if any(p is None for p in parts) else torch.cat(parts, dim=0)
    return {'x': torch.cat(xs, 0), 'edge_index': torch.cat(eis, 1), 'batch': torch.cat(batch, 0), 'y': torch.cat(ys, 0),
            'edge_attr': opt(eas), 'edge_label': opt(els), 'node_label': opt(nls), 'num_graphs': len(graphs)}
