"""torch.autograd.Function wrappers around the C-ABI kernels.  Every forward and backward here is a kernel from
libgsat_b200.so; nothing falls back to eager PyTorch or the CPU."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ._lib import lib, ptr, stream, require_cuda
from .index import GraphIndex

MODE_TRAINING, MODE_AVERAGE, MODE_INFO_ON_EDGE_ATT, MODE_NO_INFO = 1, 2, 4, 8


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise ValueError(f'fp32 tensor expected, got {t.dtype}')
    require_cuda(t)
    return t.contiguous()


# ------------------------------------------------------------------------------------------------------------
# K3  GIN aggregation   (reference src/models/conv_layers.py:14-34)
# ------------------------------------------------------------------------------------------------------------
class _GinAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, att, gi: GraphIndex, eps: float):
        x = _f32c(x)
        att_flat = None if att is None else _f32c(att).view(-1)
        N, H = x.shape
        if N != gi.N or (att_flat is not None and att_flat.numel() != gi.E):
            raise ValueError('x / edge_atten do not match the graph index')
        out = torch.empty_like(x)
        lib().call('gsatb_gin_aggregate_fwd', ptr(x), ptr(att_flat), ptr(gi.rowptr_dst), ptr(gi.eid_by_dst),
                   ptr(gi.src_by_dst), ctypes.c_float(eps), ptr(out), N, gi.E, H, stream())
        ctx.gi, ctx.eps = gi, eps
        ctx.att_shape = None if att is None else att.shape
        ctx.save_for_backward(x, att_flat)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, att_flat = ctx.saved_tensors
        gi = ctx.gi
        gout = _f32c(gout)
        N, H = x.shape
        need_x, need_att = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and att_flat is not None
        dx = torch.empty_like(x)
        datt = torch.empty(gi.E, dtype=torch.float32, device=x.device) if need_att else None
        lib().call('gsatb_gin_aggregate_bwd', ptr(gout), ptr(x), ptr(att_flat), ptr(gi.rowptr_src),
                   ptr(gi.eid_by_src), ptr(gi.dst_by_src), ctypes.c_float(ctx.eps), ptr(dx), ptr(datt), N, gi.E, H,
                   stream())
        return (dx if need_x else None), (datt.view(ctx.att_shape) if need_att else None), None, None


def gin_aggregate(x, edge_atten, gi: GraphIndex, eps: float = 0.0):
    """out[i] = sum_{e: dst(e)=i} edge_atten[e] * x[src(e)] + (1+eps) * x[i]."""
    return _GinAggregate.apply(x, edge_atten, gi, float(eps))


class _GineAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, edge_feat, att, gi: GraphIndex, eps: float):
        x, ef = _f32c(x), _f32c(edge_feat)
        att_flat = None if att is None else _f32c(att).view(-1)
        N, H = x.shape
        if N != gi.N or ef.shape != (gi.E, H) or (att_flat is not None and att_flat.numel() != gi.E):
            raise ValueError('x / edge features / edge_atten do not match the graph index')
        out = torch.empty_like(x)
        lib().call('gsatb_gine_aggregate_fwd', ptr(x), ptr(ef), ptr(att_flat), ptr(gi.rowptr_dst), ptr(gi.eid_by_dst),
                   ptr(gi.src_by_dst), ctypes.c_float(eps), ptr(out), N, gi.E, H, stream())
        ctx.gi, ctx.eps = gi, eps
        ctx.att_shape = None if att is None else att.shape
        ctx.save_for_backward(x, ef, att_flat)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, ef, att_flat = ctx.saved_tensors
        gi = ctx.gi
        gout = _f32c(gout)
        N, H = x.shape
        need_ef = ctx.needs_input_grad[1]
        need_att = ctx.needs_input_grad[2] and att_flat is not None
        dx = torch.empty_like(x)
        def_ = torch.empty_like(ef) if need_ef else None
        datt = torch.empty(gi.E, dtype=torch.float32, device=x.device) if need_att else None
        lib().call('gsatb_gine_aggregate_bwd', ptr(gout), ptr(x), ptr(ef), ptr(att_flat), ptr(gi.rowptr_src),
                   ptr(gi.eid_by_src), ptr(gi.dst_by_src), ctypes.c_float(ctx.eps), ptr(dx), ptr(def_), ptr(datt), N,
                   gi.E, H, stream())
        return dx, def_, (datt.view(ctx.att_shape) if need_att else None), None, None


def gine_aggregate(x, edge_feat, edge_atten, gi: GraphIndex, eps: float = 0.0):
    """out[i] = sum_{e: dst(e)=i} relu(x[src(e)] + edge_feat[e]) * edge_atten[e] + (1+eps) * x[i]  (GINEConv message,
    reference src/models/conv_layers.py:37-66)."""
    return _GineAggregate.apply(x, edge_feat, edge_atten, gi, float(eps))


class _LeAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, edge_weight, att, add, gi: GraphIndex):
        a, b, add = _f32c(a), _f32c(b), _f32c(add)
        w_flat = None if edge_weight is None else _f32c(edge_weight).reshape(-1)
        att_flat = None if att is None else _f32c(att).reshape(-1)
        N, H = a.shape
        if N != gi.N or b.shape != a.shape or (add is not None and add.shape != a.shape) \
                or (w_flat is not None and w_flat.numel() != gi.E) or (att_flat is not None and att_flat.numel() != gi.E):
            raise ValueError('a / b / edge_weight / edge_atten do not match the graph index')
        out = torch.empty_like(a)
        lib().call('gsatb_le_aggregate_fwd', ptr(a), ptr(b), ptr(w_flat), ptr(att_flat), ptr(gi.rowptr_dst),
                   ptr(gi.eid_by_dst), ptr(gi.src_by_dst), ptr(add), ptr(out), N, gi.E, H, stream())
        ctx.gi = gi
        ctx.w_shape = None if edge_weight is None else edge_weight.shape
        ctx.att_shape = None if att is None else att.shape
        ctx.has_add = add is not None
        ctx.save_for_backward(a, b, w_flat, att_flat)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b, w_flat, att_flat = ctx.saved_tensors
        gi = ctx.gi
        gout = _f32c(gout)
        N, H = a.shape
        need_w = ctx.needs_input_grad[2] and w_flat is not None
        need_att = ctx.needs_input_grad[3] and att_flat is not None
        da, db = torch.empty_like(a), torch.empty_like(b)
        dw = torch.empty(gi.E, dtype=torch.float32, device=a.device) if need_w else None
        datt = torch.empty(gi.E, dtype=torch.float32, device=a.device) if need_att else None
        lib().call('gsatb_le_aggregate_bwd', ptr(gout), ptr(a), ptr(b), ptr(w_flat), ptr(att_flat), ptr(gi.rowptr_src),
                   ptr(gi.eid_by_src), ptr(gi.dst_by_src), ptr(gi.rowptr_dst), ptr(gi.eid_by_dst), ptr(da), ptr(db),
                   ptr(dw), ptr(datt), N, gi.E, H, stream())
        return (da, db, (dw.view(ctx.w_shape) if need_w else None), (datt.view(ctx.att_shape) if need_att else None),
                (gout if ctx.has_add else None), None)


def le_aggregate(a, b, edge_weight, edge_atten, gi: GraphIndex, add=None):
    """out[i] = sum_{e: dst(e)=i} ((a[src(e)] - b[i]) * edge_weight[e]) * edge_atten[e] + add[i]  (LEConv message and
    root term, reference src/models/conv_layers.py:69-92); edge_weight, edge_atten and add are optional."""
    return _LeAggregate.apply(a, b, edge_weight, edge_atten, add, gi)


# ------------------------------------------------------------------------------------------------------------
# K5  readout   (global_add_pool / global_mean_pool, reference src/models/gin.py:34,53, pna.py:47,62)
# ------------------------------------------------------------------------------------------------------------
class _Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gi: GraphIndex, mean: bool):
        x = _f32c(x)
        gi.require_graph_contiguous()
        N, H = x.shape
        out = torch.empty((gi.G, H), dtype=torch.float32, device=x.device)
        lib().call('gsatb_pool_fwd', ptr(x), ptr(gi.node_ptr), ptr(out), N, gi.G, H, int(mean), stream())
        ctx.gi, ctx.mean, ctx.shape = gi, mean, (N, H)
        return out

    @staticmethod
    def backward(ctx, gout):
        gi = ctx.gi
        N, H = ctx.shape
        gout = _f32c(gout)
        dx = torch.empty((N, H), dtype=torch.float32, device=gout.device)
        lib().call('gsatb_pool_bwd', ptr(gout), ptr(gi.node_ptr), ptr(gi.node_graph), ptr(dx), N, gi.G, H,
                   int(ctx.mean), stream())
        return dx, None, None


def global_add_pool(x, gi: GraphIndex):
    return _Pool.apply(x, gi, False)


def global_mean_pool(x, gi: GraphIndex):
    return _Pool.apply(x, gi, True)


# ------------------------------------------------------------------------------------------------------------
# per-graph InstanceNorm   (PyG InstanceNorm inside reference src/utils/get_model.py:47-68)
# ------------------------------------------------------------------------------------------------------------
class _SegNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, seg_ptr, num_segments: int, eps: float):
        x = _f32c(x)
        M, C = x.shape
        y = torch.empty_like(x)
        rstd = torch.empty((num_segments, C), dtype=torch.float32, device=x.device)
        lib().call('gsatb_segnorm_fwd', ptr(x), ptr(seg_ptr), ptr(y), ptr(rstd), M, num_segments, C,
                   ctypes.c_float(eps), stream())
        ctx.save_for_backward(y, rstd, seg_ptr)
        ctx.G = num_segments
        return y

    @staticmethod
    def backward(ctx, gy):
        y, rstd, seg_ptr = ctx.saved_tensors
        gy = _f32c(gy)
        M, C = y.shape
        gx = torch.empty_like(y)
        lib().call('gsatb_segnorm_bwd', ptr(gy), ptr(y), ptr(rstd), ptr(seg_ptr), ptr(gx), M, ctx.G, C, stream())
        return gx, None, None, None


def segment_instance_norm(x, seg_ptr, num_segments: int, eps: float = 1e-5):
    return _SegNorm.apply(x, seg_ptr, num_segments, eps)


# ------------------------------------------------------------------------------------------------------------
# K2  sampler + reverse average + info loss
# ------------------------------------------------------------------------------------------------------------
class _SampleAvgInfo(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit, noise_u, rev, r_tensor, r_scalar, temp, mode, seed, offset):
        logit_c = _f32c(logit)
        flat = logit_c.view(-1)
        E = flat.numel()
        dev = flat.device
        att = torch.empty(E, dtype=torch.float32, device=dev)
        edge_att = torch.empty(E, dtype=torch.float32, device=dev)
        info = torch.zeros(1, dtype=torch.float32, device=dev)
        L = lib()
        ws_bytes = int(L.cdll.gsatb_sample_workspace(E))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        nu = None if noise_u is None else _f32c(noise_u).view(-1)
        rt = None if r_tensor is None else _f32c(r_tensor).view(-1)
        L.call('gsatb_sample_avg_info_fwd', ptr(flat), ptr(nu), ptr(rev), ptr(rt), ctypes.c_float(r_scalar),
               ctypes.c_float(temp), mode, ctypes.c_uint64(seed), ctypes.c_uint64(offset), ptr(att), ptr(edge_att),
               ptr(info), E, ptr(ws), ctypes.c_size_t(ws_bytes), stream())
        ctx.save_for_backward(att, edge_att, rev, rt)
        ctx.cfg = (float(r_scalar), float(temp), int(mode), logit.shape)
        return att.view(logit.shape), edge_att.view(logit.shape), info.view(())

    @staticmethod
    def backward(ctx, g_att, g_edge_att, g_info):
        att, edge_att, rev, rt = ctx.saved_tensors
        r_scalar, temp, mode, shape = ctx.cfg
        E = att.numel()
        dlogit = torch.empty(E, dtype=torch.float32, device=att.device)
        ga = None if g_att is None else _f32c(g_att).view(-1)
        ge = None if g_edge_att is None else _f32c(g_edge_att).view(-1)
        gi_ = None if g_info is None else _f32c(g_info).view(-1)
        lib().call('gsatb_sample_avg_info_bwd', ptr(ga), ptr(ge), ptr(gi_), ptr(att), ptr(edge_att), ptr(rev),
                   ptr(rt), ctypes.c_float(r_scalar), ctypes.c_float(temp), mode, ptr(dlogit), E, stream())
        return dlogit.view(shape), None, None, None, None, None, None, None, None


def sample_avg_info(logit, *, training: bool, rev: Optional[torch.Tensor], average: bool, r=0.5,
                    noise_u: Optional[torch.Tensor] = None, temp: float = 1.0, info_on_edge_att: bool = False,
                    want_info: bool = True, seed: int = 0, offset: int = 0):
    """Fused concrete_sample -> (att + att[rev])/2 -> info loss.  Returns (att, edge_att, info_mean)."""
    mode = (MODE_TRAINING if training else 0) | (MODE_AVERAGE if average else 0) \
        | (MODE_INFO_ON_EDGE_ATT if info_on_edge_att else 0) | (0 if want_info else MODE_NO_INFO)
    r_tensor = r if isinstance(r, torch.Tensor) else None
    r_scalar = 0.5 if r_tensor is not None else float(r)
    if average and rev is None:
        raise ValueError('average=True needs the reverse-edge map')
    return _SampleAvgInfo.apply(logit, noise_u, rev if average else None, r_tensor, r_scalar, float(temp), mode,
                                int(seed), int(offset))


class _GatherRev(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, rev, involution):
        vc = _f32c(v)
        E = vc.shape[0]
        C = vc.numel() // max(E, 1)
        out = torch.empty_like(vc)
        lib().call('gsatb_gather_rev', ptr(vc), ptr(rev), ptr(out), E, max(C, 1), stream())
        ctx.save_for_backward(rev)
        ctx.involution = bool(involution)
        return out

    @staticmethod
    def backward(ctx, g):
        (rev,) = ctx.saved_tensors
        gc = _f32c(g)
        E = gc.shape[0]
        C = gc.numel() // max(E, 1)
        if ctx.involution:
            # the reverse-edge map of a symmetric edge set is its own inverse: the adjoint is the same gather
            out = torch.empty_like(gc)
            lib().call('gsatb_gather_rev', ptr(gc), ptr(rev), ptr(out), E, max(C, 1), stream())
            return out, None, None
        # general matching permutation (reorder_like between two arbitrary orders; -1 = no partner, forward wrote 0):
        # out[i] = v[rev[i]]  =>  dv[rev[i]] += g[i]
        ok = rev >= 0
        idx = torch.where(ok, rev, torch.zeros_like(rev)).long()
        g2 = gc.reshape(E, max(C, 1)) * ok.view(-1, 1).to(gc.dtype)
        dv = torch.zeros_like(g2).index_add_(0, idx, g2)
        return dv.view_as(gc), None, None


def gather_reverse(values, rev, involution: bool = True):
    """values[rev] == reorder_like(transpose(edge_index, values), edge_index, values) on a symmetric edge set.
    ``involution=False``: ``rev`` is an arbitrary matching permutation (general reorder_like); the backward then
    scatters through it instead of re-using the gather."""
    return _GatherRev.apply(values, rev, involution)


class _Lift(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node_att, gi: GraphIndex):
        a = _f32c(node_att)
        out = torch.empty((gi.E,) + tuple(a.shape[1:]), dtype=torch.float32, device=a.device)
        lib().call('gsatb_lift_fwd', ptr(a), ptr(gi.src), ptr(gi.dst), ptr(out), gi.E, stream())
        ctx.gi = gi
        ctx.save_for_backward(a)
        return out

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        gi = ctx.gi
        g = _f32c(g)
        da = torch.empty_like(a)
        lib().call('gsatb_lift_bwd', ptr(g), ptr(a), ptr(gi.rowptr_dst), ptr(gi.eid_by_dst), ptr(gi.src_by_dst),
                   ptr(gi.rowptr_src), ptr(gi.eid_by_src), ptr(gi.dst_by_src), ptr(da), gi.N, stream())
        return da, None


def lift_node_att(node_att, gi: GraphIndex):
    """edge_att[e] = node_att[src(e)] * node_att[dst(e)]  (reference src/run_gsat.py:870-875)."""
    if node_att.numel() != gi.N:
        raise ValueError('node_att must hold one value per node')
    return _Lift.apply(node_att, gi)


class _GatherConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, gi: GraphIndex):
        emb = _f32c(emb)
        N, H = emb.shape
        out = torch.empty((gi.E, 2 * H), dtype=torch.float32, device=emb.device)
        lib().call('gsatb_gather_concat_fwd', ptr(emb), ptr(gi.src), ptr(gi.dst), ptr(out), gi.E, H, stream())
        ctx.gi, ctx.shape = gi, (N, H)
        return out

    @staticmethod
    def backward(ctx, g):
        gi = ctx.gi
        N, H = ctx.shape
        g = _f32c(g)
        demb = torch.empty((N, H), dtype=torch.float32, device=g.device)
        lib().call('gsatb_gather_concat_bwd', ptr(g), ptr(gi.rowptr_src), ptr(gi.eid_by_src), ptr(gi.rowptr_dst),
                   ptr(gi.eid_by_dst), ptr(demb), N, H, stream())
        return demb, None


def gather_concat(emb, gi: GraphIndex):
    """f12 = cat(emb[src], emb[dst])  (reference src/run_gsat.py:912-914)."""
    return _GatherConcat.apply(emb, gi)


# ------------------------------------------------------------------------------------------------------------
# K4  PNA multi-aggregator message passing   (reference src/models/conv_layers.py:160-226)
# ------------------------------------------------------------------------------------------------------------
AGG_CODES = {'sum': 0, 'mean': 1, 'min': 2, 'max': 3, 'var': 4, 'std': 5}


class _PnaAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, edge_feat, att, gi: GraphIndex, codes):
        x = _f32c(x)
        ef = _f32c(edge_feat)
        att_flat = None if att is None else _f32c(att).view(-1)
        N, H = x.shape
        He = 0 if ef is None else ef.shape[1]
        F_ = 2 * H + He
        dev = x.device
        out = torch.empty((N, len(codes) * F_), dtype=torch.float32, device=dev)
        mean = torch.empty((N, F_), dtype=torch.float32, device=dev)
        msq = torch.empty((N, F_), dtype=torch.float32, device=dev)
        amin = torch.empty((N, F_), dtype=torch.int32, device=dev)
        amax = torch.empty((N, F_), dtype=torch.int32, device=dev)
        carr = (ctypes.c_int * len(codes))(*codes)
        lib().call('gsatb_pna_aggregate_fwd', ptr(x), ptr(ef), ptr(att_flat), ptr(gi.rowptr_dst), ptr(gi.eid_by_dst),
                   ptr(gi.src_by_dst), ctypes.cast(carr, ctypes.c_void_p), len(codes), ptr(out), ptr(mean), ptr(msq),
                   ptr(amin), ptr(amax), N, gi.E, H, He, stream())
        ctx.gi, ctx.codes = gi, list(codes)
        ctx.att_shape = None if att is None else att.shape
        ctx.save_for_backward(x, ef, att_flat, mean, msq, amin, amax)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, ef, att_flat, mean, msq, amin, amax = ctx.saved_tensors
        gi, codes = ctx.gi, ctx.codes
        gout = _f32c(gout)
        N, H = x.shape
        He = 0 if ef is None else ef.shape[1]
        dev = x.device
        dx = torch.empty_like(x)
        need_ef = ef is not None and ctx.needs_input_grad[1]
        need_att = att_flat is not None and ctx.needs_input_grad[2]
        def_ = torch.empty_like(ef) if need_ef else None
        datt = torch.empty(gi.E, dtype=torch.float32, device=dev) if need_att else None
        carr = (ctypes.c_int * len(codes))(*codes)
        lib().call('gsatb_pna_aggregate_bwd', ptr(gout), ptr(x), ptr(ef), ptr(att_flat), ptr(gi.rowptr_dst),
                   ptr(gi.eid_by_dst), ptr(gi.src_by_dst), ptr(gi.rowptr_src), ptr(gi.eid_by_src), ptr(gi.dst_by_src),
                   ctypes.cast(carr, ctypes.c_void_p), len(codes), ptr(mean), ptr(msq), ptr(amin), ptr(amax), ptr(dx),
                   ptr(def_), ptr(datt), N, gi.E, H, He, stream())
        return dx, def_, (datt.view(ctx.att_shape) if need_att else None), None, None


def pna_aggregate(x, edge_feat, edge_atten, gi: GraphIndex, aggregators):
    """[N, A*(2H+He)] multi-aggregation of m_e = cat(x_i, x_j, edge_feat) * edge_atten over incoming edges."""
    return _PnaAggregate.apply(x, edge_feat, edge_atten, gi, [AGG_CODES[a] for a in aggregators])


# ------------------------------------------------------------------------------------------------------------
# node encoder Linear(x_dim, H) with a small x_dim  (reference src/models/gin.py:22-25, pna.py:20-25)
# ------------------------------------------------------------------------------------------------------------
class _SmallLinear(torch.autograd.Function):
    """y = x W^T + b for a narrow x [N, F] (F < 16): the node encoder.  Forward is the strict (split-bf16 x3) tcgen05
    GEMM of dense.py in BOTH precision modes (the raw input features are never rounded to bf16; K' = 6 F <= 96 costs
    nothing); the weight / bias gradient -- a K = N reduction -- is gsatb_linear_small_dw (exact fp32 FMA)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from .dense import linear_forward
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return linear_forward(_f32c(x), _f32c(weight.detach()), None if bias is None else _f32c(bias.detach()), True)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = _f32c(g)
        xc = _f32c(x)
        N, F_ = xc.shape
        H = weight.shape[0]
        L = lib()
        dW = torch.empty_like(weight)
        db = torch.empty(H, dtype=torch.float32, device=g.device) if ctx.has_bias else None
        ws_bytes = int(L.cdll.gsatb_linear_small_dw_workspace(N, H, F_))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
        L.call('gsatb_linear_small_dw', ptr(g), ptr(xc), ptr(dW), ptr(db), N, H, F_, ptr(ws), ctypes.c_size_t(ws_bytes),
               stream())
        dx = None
        if ctx.needs_input_grad[0]:
            from .dense import linear_forward
            dx = linear_forward(g, _f32c(weight.detach().t()), None, True)
        return dx, dW, db


def small_linear(x, weight, bias):
    return _SmallLinear.apply(x, weight, bias)


# ------------------------------------------------------------------------------------------------------------
# fused categorical encoders  (ogb AtomEncoder / BondEncoder, reference src/models/gin.py:22-25, pna.py:20-23)
# ------------------------------------------------------------------------------------------------------------
class _EmbeddingSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, table_cat, offsets_host: torch.Tensor, oob_flag):
        if idx.dtype != torch.int64 or idx.dim() != 2:
            raise ValueError(f'int64 [M, K] feature indices expected, got {idx.dtype} {tuple(idx.shape)}')
        require_cuda(idx)
        idx, table_cat = idx.contiguous(), _f32c(table_cat)
        (M, K), (R, H) = idx.shape, table_cat.shape
        if offsets_host.numel() != K + 1 or int(offsets_host[K]) != R:
            raise ValueError('feature offsets do not match the index width / the concatenated tables')
        out = torch.empty((M, H), dtype=torch.float32, device=idx.device)
        lib().call('gsatb_embedding_sum_fwd', ptr(idx), ptr(table_cat), ctypes.c_void_p(offsets_host.data_ptr()), ptr(out),
                   ptr(oob_flag), M, K, H, stream())
        ctx.save_for_backward(idx)
        ctx.offsets_host, ctx.R = offsets_host, R
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        gout = _f32c(gout)
        (M, K), H, R = idx.shape, gout.shape[1], ctx.R
        L = lib()
        ws_bytes = int(L.cdll.gsatb_embedding_sum_bwd_workspace(M, R, H))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gout.device)
        dtable = torch.empty((R, H), dtype=torch.float32, device=gout.device)
        L.call('gsatb_embedding_sum_bwd', ptr(gout), ptr(idx), ctypes.c_void_p(ctx.offsets_host.data_ptr()), ptr(dtable),
               M, K, H, ptr(ws), ctypes.c_size_t(ws_bytes), stream())
        return None, dtable, None, None


def embedding_sum(idx, tables, oob_flag=None):
    """out[m] = sum_k tables[k][idx[m, k]] in ONE kernel (and one deterministic backward into the tables), additions in
    feature order.  ``tables``: the K embedding weights [dim_k, H]; their gradients arrive through the row-wise
    concatenation.  ``oob_flag``: optional int32 [1] device word, set to 1 when an index was outside its table
    (the index is clamped: memory safe, the caller decides when to look at the flag)."""
    sizes = [int(t.shape[0]) for t in tables]
    offs = torch.zeros(len(sizes) + 1, dtype=torch.int32)
    offs[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0).to(torch.int32)
    return _EmbeddingSum.apply(idx, torch.cat(list(tables), dim=0), offs, oob_flag)
