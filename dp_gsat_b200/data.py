"""Synthetic batch generators and the minimal batch container.

The reference's data ingest (src/utils/get_data_loaders.py, src/datasets/*) is out of scope; only the *shape and
topology* of its batches matter for the hot path.  Every generator is seeded and vectorised in numpy so that the
10 M-edge configuration (cfg4) builds in seconds on the host.

Layout of a batch follows PyG ``Batch`` collate: per-graph tensors concatenated, ``edge_index``
offset by cumulative node counts, so nodes and edges are graph-contiguous and ``batch`` is non-decreasing.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch


@dataclass
class Batch:
    """Stand-in for ``torch_geometric.data.Batch`` with the attributes the step reads
    (reference src/run_gsat.py:191-259: x, edge_index, batch, edge_attr, y, edge_label)."""
    x: torch.Tensor
    edge_index: torch.Tensor            # int64 [2, E]  (row 0 = source j, row 1 = target i)
    batch: torch.Tensor                 # int64 [N]
    y: torch.Tensor                     # [G, 1]
    edge_attr: Optional[torch.Tensor] = None
    edge_label: Optional[torch.Tensor] = None
    num_graphs: int = 0
    _cache: dict = field(default_factory=dict, repr=False, compare=False)
    node_label: Optional[torch.Tensor] = None      # per-node explanation labels (mutag.py / spmotif.py Data.node_label)

    def to(self, device, non_blocking: bool = False) -> "Batch":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return Batch(mv(self.x), mv(self.edge_index), mv(self.batch), mv(self.y), mv(self.edge_attr),
                     mv(self.edge_label), self.num_graphs, node_label=mv(self.node_label))

    def pin_memory(self) -> "Batch":
        pm = lambda t: None if t is None else t.pin_memory()
        return Batch(pm(self.x), pm(self.edge_index), pm(self.batch), pm(self.y), pm(self.edge_attr),
                     pm(self.edge_label), self.num_graphs, node_label=pm(self.node_label))

    @property
    def num_nodes(self) -> int:
        return int(self.batch.numel())

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    def nbytes(self) -> int:
        tot = 0
        for t in (self.x, self.edge_index, self.batch, self.y, self.edge_attr, self.edge_label, self.node_label):
            if t is not None:
                tot += t.numel() * t.element_size()
        return tot


# ---------------------------------------------------------------------------------------------------------
# cfg1 / cfg4 / cfg5: BA-2Motifs-shaped graphs (reference src/datasets/ba_2motifs.py:19-45)
# ---------------------------------------------------------------------------------------------------------

_CYCLE = np.array([[20, 21], [21, 22], [22, 23], [23, 24], [24, 20]], dtype=np.int64)
_HOUSE_EXTRA = np.array([[21, 24]], dtype=np.int64)   # house = 5-cycle + one chord (6 motif edges)


def ba2motifs_batch(num_graphs: int, seed: int = 0, x_dim: int = 10, first_label: int = 0) -> Batch:
    """``num_graphs`` graphs of 25 nodes: a Barabasi-Albert tree (m=1) on nodes 0..19, a motif on nodes 20..24
    (label 0: 5-cycle, label 1: house) and one edge attaching the motif to the tree; labels alternate.  Edges are
    emitted in ``dense_to_sparse`` order (ascending (src, dst) inside each graph), 50 / 52 directed edges per graph.
    ``x`` is the constant 0.1 (the reference pickle holds constant features), ``edge_label`` marks motif edges."""
    rng = np.random.default_rng(seed)
    G = num_graphs
    # BA tree, m = 1: node k attaches to the endpoint of a uniformly chosen existing half-edge
    ends = np.zeros((G, 38), dtype=np.int64)
    und = np.zeros((G, 26, 2), dtype=np.int64)
    ends[:, 0], ends[:, 1] = 0, 1
    und[:, 0, 0], und[:, 0, 1] = 0, 1
    for k in range(2, 20):
        pick = rng.integers(0, 2 * (k - 1), size=G)
        tgt = ends[np.arange(G), pick]
        ends[:, 2 * (k - 1)] = k
        ends[:, 2 * (k - 1) + 1] = tgt
        und[:, k - 1, 0] = tgt
        und[:, k - 1, 1] = k
    und[:, 19:24, :] = _CYCLE[None]
    und[:, 24, 0] = rng.integers(0, 20, size=G)
    und[:, 24, 1] = 20
    label = (np.arange(G) + first_label) % 2
    und[:, 25, :] = _HOUSE_EXTRA[0][None]
    valid = np.ones((G, 26), dtype=bool)
    valid[:, 25] = label == 1
    # both directions, sorted row-major inside each graph; invalid slots pushed to the end
    s = np.concatenate([und[:, :, 0], und[:, :, 1]], axis=1)
    d = np.concatenate([und[:, :, 1], und[:, :, 0]], axis=1)
    v = np.concatenate([valid, valid], axis=1)
    key = np.where(v, s * 25 + d, 25 * 25)
    key.sort(axis=1)
    keep = key < 25 * 25
    off = (np.arange(G, dtype=np.int64) * 25)[:, None]
    src = (key // 25 + off)[keep]
    dst = (key % 25 + off)[keep]
    edge_index = torch.from_numpy(np.stack([src, dst], axis=0))
    ls, ld = (key // 25)[keep], (key % 25)[keep]
    edge_label = torch.from_numpy(((ls >= 20) & (ld >= 20)).astype(np.float32))
    N = G * 25
    x = torch.full((N, x_dim), 0.1, dtype=torch.float32)
    batch = torch.arange(G, dtype=torch.int64).repeat_interleave(25)
    y = torch.from_numpy(label.astype(np.float32)).view(-1, 1)
    return Batch(x, edge_index, batch, y, None, edge_label, G)


# ---------------------------------------------------------------------------------------------------------
# cfg2: Mutagenicity topology (reference data/mutag_dual/raw/*) and its line-graph ("dual") variant
# ---------------------------------------------------------------------------------------------------------


def batch_from_edge_list(src: np.ndarray, dst: np.ndarray, node_graph: np.ndarray, x_dim: int, seed: int = 0,
                         y: Optional[np.ndarray] = None) -> Batch:
    """Wrap a global edge list (0-based node ids, edge order kept as given) into a Batch with uniform features."""
    g = torch.Generator().manual_seed(seed)
    N = int(node_graph.shape[0])
    G = int(node_graph.max()) + 1 if N else 0
    x = torch.rand((N, x_dim), generator=g, dtype=torch.float32)
    if y is None:
        y = (np.arange(G) % 2).astype(np.float32)
    ei = torch.from_numpy(np.stack([src, dst], axis=0).astype(np.int64))
    return Batch(x, ei, torch.from_numpy(node_graph.astype(np.int64)), torch.from_numpy(y.astype(np.float32)).view(-1, 1),
                 None, torch.zeros(ei.shape[1]), G)


def line_graph_dual(src: np.ndarray, dst: np.ndarray, node_graph: np.ndarray):
    """Line-graph construction of the fork (reference src/datasets/mutag_dual.py:342-378): one dual node per directed
    primal edge (a, b); dual nodes sharing the same first endpoint ``a`` are pairwise connected, both directions
    emitted back to back ((p, q) then (q, p), p < q in primal-edge order), groups in order of first appearance.
    Returns (dual_src, dual_dst, dual_node_graph).  E_dual = sum_v d(v) (d(v) - 1)."""
    E = src.shape[0]
    order = np.argsort(src, kind='stable')              # dual nodes grouped by first endpoint, primal order kept
    s_sorted = src[order]
    starts = np.flatnonzero(np.r_[True, s_sorted[1:] != s_sorted[:-1]])
    counts = np.diff(np.r_[starts, E])
    first_seen = order[starts]                           # group order = order of first appearance
    gorder = np.argsort(first_seen, kind='stable')
    ds, dd = [], []
    for gi in gorder:                                    # groups are tiny (degree <= ~10); pairs built vectorised
        c = counts[gi]
        if c < 2:
            continue
        members = order[starts[gi]:starts[gi] + c]
        iu, ju = np.triu_indices(c, k=1)
        p, q = members[iu], members[ju]
        ds.append(np.stack([p, q], axis=1).reshape(-1))
        dd.append(np.stack([q, p], axis=1).reshape(-1))
    if ds:
        dsrc, ddst = np.concatenate(ds), np.concatenate(dd)
    else:
        dsrc = ddst = np.zeros(0, dtype=np.int64)
    return dsrc.astype(np.int64), ddst.astype(np.int64), node_graph[src].astype(np.int64)


def load_mutag_fixture(path: str):
    """Load the committed slice of the Mutagenicity topology (tests/golden/mutag_slice.npz, written by
    tests/golden/make_golden.py from reference data/mutag_dual/raw/Mutagenicity_A.txt)."""
    z = np.load(path)
    return z['src'].astype(np.int64), z['dst'].astype(np.int64), z['node_graph'].astype(np.int64)


def graph_contiguous_relabel(dsrc, ddst, dnode_graph):
    """Dual nodes are primal edges, which are already graph-contiguous in the Mutagenicity file; dual edges are
    emitted group by group, so they must be bucketed by graph to follow PyG collate order (App. A.9)."""
    eg = dnode_graph[dsrc]
    order = np.argsort(eg, kind='stable')
    return dsrc[order], ddst[order]


# ---------------------------------------------------------------------------------------------------------
# cfg3: ogbg-molhiv-shaped batches for PNA
# ---------------------------------------------------------------------------------------------------------

ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_FEATURE_DIMS = [5, 6, 2]


def molhiv_like_batch(num_graphs: int = 256, seed: int = 0, with_edge_attr: bool = True) -> Batch:
    """Molecule-shaped graphs: nodes/graph ~ clipped N(25.5, 12) in [2, 222]; a random spanning tree plus ring
    closures up to ~27.5 undirected edges on average; integer atom features [N, 9] and bond features [E, 3] in the
    ogb ranges; both directions emitted back to back (ogb smiles2graph order)."""
    rng = np.random.default_rng(seed)
    n = np.clip(np.rint(rng.normal(25.5, 12.0, size=num_graphs)), 2, 222).astype(np.int64)
    srcs, dsts, node_graph = [], [], []
    off = 0
    for g in range(num_graphs):
        k = int(n[g])
        parent = np.array([rng.integers(max(0, i - 3), i) for i in range(1, k)], dtype=np.int64)
        u = np.arange(1, k, dtype=np.int64)
        und = np.stack([parent, u], axis=1)
        n_ring = int(round(k * 2.0 / 25.5))
        extra = []
        have = set(map(tuple, und.tolist()))
        tries = 0
        while len(extra) < n_ring and tries < 10 * n_ring and k > 4:
            a = int(rng.integers(0, k - 3))
            b = a + int(rng.integers(3, min(6, k - a)))
            tries += 1
            if b < k and (a, b) not in have:
                have.add((a, b))
                extra.append((a, b))
        if extra:
            und = np.concatenate([und, np.array(extra, dtype=np.int64)], axis=0)
        s = np.stack([und[:, 0], und[:, 1]], axis=1).reshape(-1) + off
        d = np.stack([und[:, 1], und[:, 0]], axis=1).reshape(-1) + off
        srcs.append(s)
        dsts.append(d)
        node_graph.append(np.full(k, g, dtype=np.int64))
        off += k
    src, dst = np.concatenate(srcs), np.concatenate(dsts)
    node_graph = np.concatenate(node_graph)
    N, E = node_graph.shape[0], src.shape[0]
    x = np.stack([rng.integers(0, d, size=N) for d in ATOM_FEATURE_DIMS], axis=1).astype(np.int64)
    ea_und = np.stack([rng.integers(0, d, size=E // 2) for d in BOND_FEATURE_DIMS], axis=1).astype(np.int64)
    ea = np.repeat(ea_und, 2, axis=0)
    y = (rng.random(num_graphs) < 0.5).astype(np.float32)
    return Batch(torch.from_numpy(x), torch.from_numpy(np.stack([src, dst], 0)), torch.from_numpy(node_graph),
                 torch.from_numpy(y).view(-1, 1), torch.from_numpy(ea) if with_edge_attr else None,
                 torch.zeros(E), num_graphs)


def in_degree_histogram(batch: Batch, minlength: int = 10) -> torch.Tensor:
    """``deg`` for PNA (reference src/utils/get_data_loaders.py:99-101): histogram of node in-degrees."""
    d = torch.bincount(batch.edge_index[1], minlength=batch.num_nodes)
    return torch.bincount(d, minlength=minlength)


# ---------------------------------------------------------------------------------------------------------
# graph sharding for data parallelism
# ---------------------------------------------------------------------------------------------------------


def shard_bounds_by_edges(edge_ptr: np.ndarray, world_size: int) -> np.ndarray:
    """Contiguous graph ranges [g_k, g_{k+1}) whose edge counts are as equal as a prefix-sum split allows."""
    G = edge_ptr.shape[0] - 1
    E = int(edge_ptr[-1])
    targets = (np.arange(1, world_size, dtype=np.float64) * E / world_size)
    cuts = np.searchsorted(edge_ptr, targets, side='left')
    cuts = np.clip(cuts, 0, G)
    return np.concatenate([[0], cuts, [G]]).astype(np.int64)


def shard_batch(b: Batch, rank: int, world_size: int) -> Batch:
    """Slice graphs [g0, g1) of a graph-contiguous batch and rebase node ids (one shard per rank)."""
    if world_size == 1:
        return b
    G = b.num_graphs
    src = b.edge_index[0]
    eg = b.batch[src]
    edge_ptr = torch.zeros(G + 1, dtype=torch.int64)
    edge_ptr[1:] = torch.cumsum(torch.bincount(eg, minlength=G), 0)
    node_ptr = torch.zeros(G + 1, dtype=torch.int64)
    node_ptr[1:] = torch.cumsum(torch.bincount(b.batch, minlength=G), 0)
    bounds = shard_bounds_by_edges(edge_ptr.numpy(), world_size)
    g0, g1 = int(bounds[rank]), int(bounds[rank + 1])
    n0, n1 = int(node_ptr[g0]), int(node_ptr[g1])
    e0, e1 = int(edge_ptr[g0]), int(edge_ptr[g1])
    sl = lambda t, a, c: None if t is None else t[a:c].clone()
    return Batch(sl(b.x, n0, n1), (b.edge_index[:, e0:e1] - n0).clone(), (b.batch[n0:n1] - g0).clone(),
                 sl(b.y, g0, g1), sl(b.edge_attr, e0, e1), sl(b.edge_label, e0, e1), g1 - g0)
