"""Device-resident dataset and batch collate (the step before the hot path).

The reference builds every batch on the host: ``DataLoader(dataset[split], batch_size, shuffle=False)``
(src/utils/get_data_loaders.py:130-145) runs torch_geometric's collate -- per-graph tensors concatenated, ``edge_index``
shifted by cumulative node counts, ``batch`` vector built -- and the trainer then copies the batch to
the GPU (``data.to(device)``, src/run_gsat.py:296).  GSAT's datasets are a few MB to a few GB: on a 180 GB B200 the
whole packed dataset lives in HBM, and a batch is gathered from the list of graph ids by two kernels
(csrc/collate.cu); the only per-batch H2D traffic is the id list and two (B+1)-entry pointer arrays.

  PackedDataset.from_data_list(graphs)      pack once (host), upload once
  PackedDataset.collate(ids) -> Batch       == Batch.from_data_list([graphs[i] for i in ids]) on the device, bit-exact
  DeviceLoader(ds, ids, batch_size)         iterates like the reference loaders (shuffle=False by default)
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from ._lib import lib, ptr, stream
from .data import Batch


class Graph:
    """One sample, the fields of ``torch_geometric.data.Data`` the reference datasets fill (x, edge_index with
    graph-local node ids, y, edge_attr, edge_label, node_label)."""

    def __init__(self, x, edge_index, y, edge_attr=None, edge_label=None, node_label=None):
        self.x, self.edge_index, self.y = x, edge_index, y
        self.edge_attr, self.edge_label, self.node_label = edge_attr, edge_label, node_label


def _rows(t: torch.Tensor) -> torch.Tensor:
    """[rows, ...] -> contiguous [rows, row_bytes / itemsize]."""
    width = 1
    for d in t.shape[1:]:
        width *= int(d)
    return t.reshape(t.shape[0], width).contiguous()          # (explicit width: -1 is ambiguous for an empty tensor)


class PackedDataset:
    """All graphs of a dataset concatenated ("packed") and resident on ``device``; ``edge_index`` keeps graph-local ids."""

    PER_NODE = ('x', 'node_label')
    PER_EDGE = ('edge_attr', 'edge_label')

    def __init__(self, tensors: dict, node_ptr: np.ndarray, edge_ptr: np.ndarray, device):
        self.device = torch.device(device)
        self.node_ptr_host = np.ascontiguousarray(node_ptr, dtype=np.int64)
        self.edge_ptr_host = np.ascontiguousarray(edge_ptr, dtype=np.int64)
        self.num_graphs = len(node_ptr) - 1
        self.node_ptr = torch.from_numpy(self.node_ptr_host).to(self.device)
        self.edge_ptr = torch.from_numpy(self.edge_ptr_host).to(self.device)
        self.t = {k: (None if v is None else v.to(self.device)) for k, v in tensors.items()}
        for k in self.PER_NODE + self.PER_EDGE + ('y',):
            v = self.t.get(k)
            if v is not None and (v.element_size() * (v.numel() // max(v.shape[0], 1))) % 4 != 0:
                raise ValueError(f'{k}: rows must be a multiple of 4 bytes (got dtype {v.dtype}, shape {tuple(v.shape)})')

    @staticmethod
    def from_data_list(graphs: Sequence[Graph], device='cuda') -> 'PackedDataset':
        n = np.array([g.x.shape[0] for g in graphs], dtype=np.int64)
        e = np.array([g.edge_index.shape[1] for g in graphs], dtype=np.int64)
        node_ptr = np.concatenate([[0], np.cumsum(n)])
        edge_ptr = np.concatenate([[0], np.cumsum(e)])

        def cat(name, dim=0):
            vals = [getattr(g, name) for g in graphs]
            if any(v is None for v in vals):
                if not all(v is None for v in vals):
                    raise ValueError(f'{name} is set on some graphs only')
                return None
            return torch.cat(vals, dim=dim)
        tensors = {'x': cat('x'), 'edge_index': cat('edge_index', 1).to(torch.int64).contiguous(),
                   'edge_attr': cat('edge_attr'), 'edge_label': cat('edge_label'), 'node_label': cat('node_label'),
                   'y': cat('y')}
        if tensors['y'] is None or tensors['y'].shape[0] != len(graphs):
            raise ValueError('every graph needs a label y with a leading dimension of 1 (PyG collate concatenates them)')
        ei = tensors['edge_index']
        if ei.numel() and (int(ei.min()) < 0 or bool((ei >= torch.from_numpy(np.repeat(n, e)).unsqueeze(0)).any())):
            raise ValueError('edge_index must hold graph-local node ids in [0, num_nodes of its graph)')
        return PackedDataset(tensors, node_ptr, edge_ptr, device)

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.t.values() if v is not None)

    # -----------------------------------------------------------------------------------------------------
    def collate(self, ids: Iterable[int]) -> Batch:
        """``Batch.from_data_list([graph[i] for i in ids])`` gathered on the device."""
        ids_host = np.ascontiguousarray(np.fromiter(ids, dtype=np.int64) if not isinstance(ids, np.ndarray) else ids,
                                        dtype=np.int64)
        B = int(ids_host.shape[0])
        if B == 0:
            raise ValueError('empty batch')
        if int(ids_host.min()) < 0 or int(ids_host.max()) >= self.num_graphs:
            raise IndexError('graph id out of range')
        n = self.node_ptr_host[ids_host + 1] - self.node_ptr_host[ids_host]
        e = self.edge_ptr_host[ids_host + 1] - self.edge_ptr_host[ids_host]
        meta = np.zeros(3 * B + 2, dtype=np.int64)                 # [ids | out_node_ptr | out_edge_ptr]: one H2D copy
        meta[:B] = ids_host
        np.cumsum(n, out=meta[B + 1:2 * B + 1])
        np.cumsum(e, out=meta[2 * B + 2:3 * B + 2])
        N_out, E_out = int(meta[2 * B]), int(meta[3 * B + 1])
        meta_dev = torch.from_numpy(meta).to(self.device, non_blocking=True)
        ids_dev, out_node_ptr, out_edge_ptr = meta_dev[:B], meta_dev[B:2 * B + 1], meta_dev[2 * B + 1:]
        L, st = lib(), stream()
        dev = self.device

        def gather(name, ds_ptr, out_ptr, rows_out, want_batch=False):
            src = self.t.get(name)
            batch_vec = torch.empty(rows_out, dtype=torch.int64, device=dev) if want_batch else None
            if src is None:
                if want_batch:
                    L.call('gsatb_collate_rows', None, 0, ptr(ds_ptr), ptr(ids_dev), ptr(out_ptr), B, rows_out, None,
                           ptr(batch_vec), st)
                return None, batch_vec
            flat = _rows(src)
            out = torch.empty((rows_out,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
            L.call('gsatb_collate_rows', ptr(flat), flat.shape[1] * flat.element_size(), ptr(ds_ptr), ptr(ids_dev),
                   ptr(out_ptr), B, rows_out, ptr(out), ptr(batch_vec), st)
            return out, batch_vec

        x, batch_vec = gather('x', self.node_ptr, out_node_ptr, N_out, want_batch=True)
        node_label, _ = gather('node_label', self.node_ptr, out_node_ptr, N_out)
        edge_attr, _ = gather('edge_attr', self.edge_ptr, out_edge_ptr, E_out)
        edge_label, _ = gather('edge_label', self.edge_ptr, out_edge_ptr, E_out)
        slot_ptr = torch.arange(B + 1, dtype=torch.int64, device=dev)
        y, _ = gather('y', None, slot_ptr, B)
        ei = torch.empty((2, E_out), dtype=torch.int64, device=dev)
        ds_ei = self.t['edge_index']
        L.call('gsatb_collate_edge_index', ptr(ds_ei), ds_ei.shape[1], ptr(self.edge_ptr), ptr(ids_dev),
               ptr(out_edge_ptr), ptr(out_node_ptr), B, E_out, ptr(ei), st)
        return Batch(x, ei, batch_vec, y, edge_attr, edge_label, B, node_label=node_label)


class DeviceLoader:
    """``DataLoader(dataset[ids], batch_size, shuffle)`` of the reference loaders (get_data_loaders.py:130-145; the fork
    runs them with shuffle=False) over a PackedDataset: yields device-resident Batch objects."""

    def __init__(self, dataset: PackedDataset, ids: Optional[Sequence[int]] = None, batch_size: int = 128,
                 shuffle: bool = False, seed: int = 0, drop_last: bool = False, cache: bool = False):
        """``cache=True`` (shuffle=False only -- the fork's loaders, get_data_loaders.py:133-135): every collated batch
        stays resident and is handed out again in later epochs, at the same addresses.  The K0 index of a batch is then
        built once for the whole run (the index cache is sized to the loader) and a step can be replayed as a CUDA graph
        (parallel.TrainStep.enable_cuda_graph needs its inputs to stay put); costs one copy of the split in HBM."""
        self.dataset = dataset
        self.ids = np.arange(dataset.num_graphs, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self._rng = np.random.default_rng(seed)
        if cache and shuffle:
            raise ValueError('cache=True keeps the batches of an epoch resident: it needs shuffle=False')
        self.cache = cache
        self._batches = {}

    def __len__(self) -> int:
        n = len(self.ids)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = self._rng.permutation(self.ids) if self.shuffle else self.ids
        if self.cache:
            from .index import _CACHE_CAP, set_index_cache_capacity
            if _CACHE_CAP < len(self) + 8:
                set_index_cache_capacity(len(self) + 8)
        for i in range(len(self)):
            if not self.cache:
                yield self.dataset.collate(order[i * self.batch_size:(i + 1) * self.batch_size])
                continue
            if i not in self._batches:
                self._batches[i] = self.dataset.collate(order[i * self.batch_size:(i + 1) * self.batch_size])
            yield self._batches[i]


def split_batch(b: Batch) -> List[Graph]:
    """Inverse of collate on a HOST batch (test / tooling helper): the per-graph samples with graph-local node ids."""
    G = b.num_graphs
    node_ptr = np.concatenate([[0], np.cumsum(np.bincount(b.batch.numpy(), minlength=G))])
    eg = b.batch[b.edge_index[0]].numpy()
    edge_ptr = np.concatenate([[0], np.cumsum(np.bincount(eg, minlength=G))])
    out = []
    for g in range(G):
        n0, n1, e0, e1 = int(node_ptr[g]), int(node_ptr[g + 1]), int(edge_ptr[g]), int(edge_ptr[g + 1])
        sl = lambda t, a, c: None if t is None else t[a:c].clone()
        out.append(Graph(b.x[n0:n1].clone(), (b.edge_index[:, e0:e1] - n0).clone(), b.y[g:g + 1].clone(),
                         sl(b.edge_attr, e0, e1), sl(b.edge_label, e0, e1), sl(getattr(b, 'node_label', None), n0, n1)))
    return out
