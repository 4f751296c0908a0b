"""Graph-sharded data parallelism: one process per GPU, replicas of the ~70-200 k parameters, each
rank owns a contiguous range of graphs chosen on the prefix sum of edges per graph, and ONE all-reduce per step
over a single flat fp32 gradient bucket (NCCL over NVLink 5 / NVSwitch; payload < 1 MB, latency bound).

The reference has no distributed code at all (single device, src/run_gsat.py:1069); this is new capability.
Loss terms are means over GLOBAL counts: the local info loss is weighted by E_local/E_global and the local
prediction loss by G_local/G_global, so the SUM of the ranks' gradients equals the single-device gradient (up to
BatchNorm, whose batch statistics are shard-local by default, as in torch DDP; ``enable_sync_batchnorm`` /
``TrainStep(sync_bn=True)`` all-reduce them for exact single-device math).

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


def enable_sync_batchnorm(module: torch.nn.Module, group=None, enabled: bool = True) -> int:
    """Make every BatchNorm1d of ``module`` (the GIN node MLPs, gin.py:59; PNA's per-layer norms, pna.py:45) take its
    training-mode batch statistics over ALL ranks of ``group`` instead of over the local shard, so that the
    graph-sharded step equals the single-device step exactly.  Costs two small all-reduces per
    BatchNorm forward and one per backward (fp32 path), one each way on the tensor-core path.  Returns the number of
    layers switched.  ``enabled=False`` restores shard-local statistics (DDP semantics, the default)."""
    from .nn import BatchNorm1d
    n = 0
    for m in module.modules():
        if isinstance(m, BatchNorm1d):
            m.sync_group = (group if group is not None else dist.group.WORLD) if enabled else None
            n += 1
    return n


class TrainStep:
    """forward_pass -> backward -> (all-reduce) -> Adam.step for one (sharded) batch.  ``gsat`` is any object with
    the reference's ``forward_pass(data, epoch, training)`` and the ``pred_scale`` / ``info_scale`` attributes."""

    def __init__(self, gsat, lr: float = 1e-3, weight_decay: float = 0.0, group=None, fused_adam: Optional[bool] = None,
                 sync_bn: bool = False):
        self.gsat = gsat
        self.group = group
        if sync_bn and dist.is_available() and dist.is_initialized():
            enable_sync_batchnorm(gsat.clf, group)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # every rank draws its own noise / dropout streams: the sampler's Philox stream and the dropout hashes are
            # indexed by LOCAL edge / row ids, so identical seeds would give every shard the same draws
            r = dist.get_rank(group)
            for m in (gsat, gsat.clf, gsat.extractor):
                if hasattr(m, 'seed'):
                    m.seed = int(m.seed) + 1000003 * (r + 1)
        params = list(gsat.extractor.parameters()) + list(gsat.clf.parameters())   # order of src/run_gsat.py:1007
        self.bucket = FlatGradBucket(params)
        if fused_adam is None:
            fused_adam = self.bucket.flat.is_cuda
        # capturable: the step count lives on the device, so the optimizer step can sit inside a CUDA graph
        self.optimizer = torch.optim.Adam(self.bucket.params, lr=lr, weight_decay=weight_decay, fused=fused_adam,
                                          capturable=bool(fused_adam))
        self.graph = None            # torch.cuda.CUDAGraph of one whole step (enable_cuda_graph)
        self._graph_data = self._graph_index = None
        self._graph_key = None
        self._graph_out = None
        self.launches_per_step = None

    def set_shard_weights(self, data):
        """Weights of the local loss terms (E_local/E_global, G_local/G_global).  The global counts cost one tiny
        all-reduce; the result is cached ON THE BATCH OBJECT, so the hit / miss decision is a property of the batch
        object the (SPMD) training loop hands to every rank at the same step -- never of local counts that two
        different batches can share on one rank and not on another (which would leave ranks issuing different
        collectives)."""
        cache = getattr(data, '_cache', None)
        key = ('shard_weights', id(self))
        if cache is not None and key in cache:
            self.gsat.info_scale, self.gsat.pred_scale = cache[key]
            return
        learn = getattr(self.gsat, 'learn_edge_att', True)
        n_att = data.edge_index.shape[1] if learn or getattr(self.gsat, 'info_on', 'att') == 'edge_att' \
            else data.batch.numel()
        n_g = int(data.y.shape[0])
        eg, gg = global_counts(n_att, n_g, data.x.device, self.group)
        w = (n_att / max(eg, 1), n_g / max(gg, 1))
        if cache is not None:
            cache[key] = w
        self.gsat.info_scale, self.gsat.pred_scale = w

    @staticmethod
    def _key(data, epoch):
        return (data.x.data_ptr(), data.edge_index.data_ptr(), data.batch.data_ptr(), data.y.data_ptr(),
                tuple(data.edge_index.shape), int(epoch))

    def enable_cuda_graph(self, data, epoch: int, warmup: int = 3):
        """Capture forward_pass + backward + gradient all-reduce + Adam.step on ``data`` (which must stay resident at
        the same addresses: the reference's loaders replay identical batches every epoch) into ONE CUDA graph, so a
        step costs one launch on the host instead of several hundred.  Fresh noise / dropout per replay comes from the
        device step counter (gsat.step_counter).  Falls back to eager steps for any other batch / epoch / injected
        noise.  Returns True when the graph is in place."""
        from ._lib import lib
        self.set_shard_weights(data)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on a side stream, as graph capture requires
            for _ in range(warmup):
                self._eager(data, epoch, None)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = lib().launches
        try:
            with torch.cuda.graph(graph):
                out = self._eager(data, epoch, None)
        except Exception as exc:      # e.g. an op that cannot be captured on this build: stay on the eager path
            import warnings
            warnings.warn(f'CUDA graph capture of the training step failed ({exc!r}); using eager steps')
            torch.cuda.synchronize()
            return False
        self.launches_per_step = lib().launches - n0
        self.graph, self._graph_key, self._graph_out = graph, self._key(data, epoch), out
        # the batch and the index bundle the graph reads (load_batch refreshes them in place)
        from .index import get_graph_index
        self._graph_data = data
        self._graph_index = get_graph_index(data.edge_index, data.batch, getattr(data, 'num_graphs', None))
        return True

    def disable_cuda_graph(self):
        self.graph = self._graph_key = self._graph_out = None
        self._graph_data = self._graph_index = None

    _BATCH_FIELDS = ('x', 'edge_index', 'batch', 'y', 'edge_attr', 'edge_label', 'node_label')
    _INDEX_FIELDS = ('src', 'dst', 'rev', 'rowptr_dst', 'eid_by_dst', 'src_by_dst', 'rowptr_src', 'eid_by_src',
                     'dst_by_src', 'node_ptr', 'edge_ptr', 'node_graph', 'edge_graph', 'flags_dev')

    def load_batch(self, batch, index=None) -> bool:
        """Refresh the captured graph's RESIDENT batch with a fresh one of the same shape: a loader that streams new
        batches (``data.to(device)`` every step, src/run_gsat.py:296) keeps the one-launch step by copying each batch --
        and the index bundle K0 built for it, e.g. one step ahead with ``prefetch_graph_index`` -- into the static
        buffers the graph reads (device-to-device, on the current stream), then calling the step on the resident batch.
        Everything the capture baked in on the host is checked first: node / edge / graph counts, tensor shapes, the
        index flags (symmetry, duplicates, contiguity) and the fused extractor's tile counts; returns False (nothing
        copied) when the new batch does not fit, and the caller takes an eager step on it instead."""
        if self.graph is None or getattr(self, '_graph_data', None) is None:
            return False
        from .index import get_graph_index
        d, gi = self._graph_data, self._graph_index
        if index is None:
            index = get_graph_index(batch.edge_index, batch.batch, getattr(batch, 'num_graphs', None))
        if (index.N, index.E, index.G) != (gi.N, gi.E, gi.G) or list(index.flags) != list(gi.flags):
            return False
        pairs = []
        for name in self._BATCH_FIELDS:
            a, b = getattr(d, name, None), getattr(batch, name, None)
            if (a is None) != (b is None):
                return False
            if a is not None:
                if a.shape != b.shape or a.dtype != b.dtype:
                    return False
                pairs.append((a, b))
        for key, plan in (gi._plans or {}).items():
            if not (isinstance(plan, dict) and key[0] == 'ext'):
                return False                       # host-built plans are not refreshed
            other = index.ext_plan(key[1], key[2])
            if other['T'] != plan['T'] or other['oversize'] != plan['oversize']:
                return False
            pairs += [(plan['tile_seg'], other['tile_seg']), (plan['out2'], other['out2'])]
        pairs += [(getattr(gi, n), getattr(index, n)) for n in self._INDEX_FIELDS]
        with torch.no_grad():
            for dst, src in pairs:
                dst.copy_(src, non_blocking=True)
        return True

    def __call__(self, data, epoch: int, noise_u=None):
        if self.graph is not None and noise_u is None and self._key(data, epoch) == self._graph_key:
            self.graph.replay()
            return self._graph_out
        return self._eager(data, epoch, noise_u)

    def _eager(self, data, epoch: int, noise_u=None):
        self.set_shard_weights(data)
        edge_att, loss, loss_dict, clf_logits = self.gsat.forward_pass(data, epoch, True, noise_u=noise_u)
        self.bucket.zero()
        loss.backward()
        self.bucket.all_reduce(self.group)
        self.optimizer.step()
        return edge_att, loss, loss_dict, clf_logits
