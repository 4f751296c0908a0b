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
           want_stats: bool = False, pdrop: float = 0.0, drop_seed: int = 0, drop_mask: Optional[torch.Tensor] = None):
    """out = drop(act(pro(x) W^T + bias)) on the tensor cores; optional per-channel (sum z, sum z^2) in fp64."""
    x = x.contiguous()
    rows, K = x.shape
    out = torch.empty((rows, out_features), dtype=torch.float32, device=x.device)
    part = stats = None
    if want_stats:
        part = torch.empty(int(lib().cdll.gsatb_tc_stat_partials_elems(out_features)), dtype=torch.float32,
                           device=x.device)
        stats = torch.empty(2 * out_features, dtype=torch.float64, device=x.device)
    lib().call('gsatb_tc_linear_fwd', ptr(x), int(x.dtype == torch.bfloat16), K, ptr(in_scale), ptr(in_shift), ptr(wp),
               ptr(bias), ptr(out),
               out_features, int(relu_out), ptr(part), ptr(stats), ptr(drop_mask), ctypes.c_uint64(drop_seed),
               ctypes.c_float(pdrop), rows, K, out_features, stream())
    return (out, stats) if want_stats else out


def linear_bf16(x16: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], out_features: int,
                out_bf16: bool = True, relu_out: bool = False, want_stats: bool = False, pdrop: float = 0.0,
                drop_seed: int = 0, drop_mask: Optional[torch.Tensor] = None, posmask: Optional[torch.Tensor] = None):
    """out = drop(act(x16 W^T + bias)) with the B operand fed by TMA straight from the bf16 activations (four epilogue
    groups); bf16 or fp32 output; optional per-channel (sum z, sum z^2) in fp64, taken from the fp32 accumulators."""
    x16 = x16.contiguous()
    rows, K = x16.shape
    out = torch.empty((rows, out_features), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x16.device)
    part = stats = None
    if want_stats:
        part = torch.empty(int(lib().cdll.gsatb_tc_stat_partials_elems(out_features)), dtype=torch.float32,
                           device=x16.device)
        stats = torch.empty(2 * out_features, dtype=torch.float64, device=x16.device)
    lib().call('gsatb_tc_linear_bf16_fwd', ptr(x16), K, ptr(wp), ptr(bias), ptr(out), int(out_bf16), out_features,
               int(relu_out), ptr(part), ptr(stats), ptr(drop_mask), ctypes.c_uint64(drop_seed), ctypes.c_float(pdrop),
               ptr(posmask), rows, K, out_features, stream())
    return (out, stats) if want_stats else out


def extractor_forward(emb: torch.Tensor, gi: GraphIndex, w1: torch.Tensor, w2: torch.Tensor, w3: torch.Tensor,
                      b3: Optional[torch.Tensor], *, edge_mode: bool, pdrop: float, training: bool, seed: int,
                      mask1: Optional[torch.Tensor] = None, mask2: Optional[torch.Tensor] = None, eps: float = 1e-5,
                      keep_h1: bool = False):
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
    # f12 = [emb[src] | emb[dst]] as bf16: the B operand of GEMM1 (TMA-fed) and, in backward, of dW1 = dz1^T f12
    Kin = 2 * H if edge_mode else H
    f12 = torch.empty((rows, Kin), dtype=torch.bfloat16, device=dev)
    L.call('gsatb_tc_ext_make_f12', ptr(emb), ptr(gi.src) if edge_mode else None, ptr(gi.dst) if edge_mode else None,
           ptr(f12), rows, H, stream())
    L.call('gsatb_tc_ext_fwd1', ptr(f12), ptr(w1p), ptr(tile_row), ptr(tile_seg), ptr(seg_ptr), T, ptr(xhat1),
           ptr(rstd1), rows, Kin, C1, ctypes.c_float(eps), stream())
    if not keep_h1:
        f12 = None
    w3f = w3.detach().reshape(-1).contiguous()
    # h1 = Dropout(ReLU(xhat1)): the B operand of GEMM2 (TMA-fed) and, in backward, of dW2 = dz2^T h1
    h1 = torch.empty((rows, C1), dtype=torch.bfloat16, device=dev)
    L.call('gsatb_tc_ext_make_h1', ptr(xhat1), ptr(mask1), ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training),
           ptr(h1), rows, C1, stream())
    L.call('gsatb_tc_ext_fwd2', ptr(h1), ptr(w2p), ptr(w3f), ptr(b3), ptr(mask2), ctypes.c_uint64(seed),
           ctypes.c_float(pdrop), int(training), ptr(tile_row), ptr(tile_seg), ptr(seg_ptr), T, ptr(xhat2), ptr(rstd2),
           ptr(logit), rows, C1, H, ctypes.c_float(eps), stream())
    if not keep_h1:
        h1 = None
    return logit, dict(xhat1=xhat1, rstd1=rstd1, xhat2=xhat2, rstd2=rstd2, plan=plan, h1=h1, f12=f12)


class _FusedExtractor(torch.autograd.Function):
    """Extractor MLP (Linear -> InstanceNorm -> ReLU -> Dropout, twice, then Linear(H, 1)) as five tensor-core /
    segment kernels forward and backward.  Parameter gradients: dW2 = dz2^T h1 and dW1 = dz1^T f12 are plain
    library GEMMs on the re-materialised bf16 operands; b1 / b2 sit in front of an InstanceNorm and get exact zeros."""

    @staticmethod
    def forward(ctx, emb, w1, b1, w2, b2, w3, b3, gi, edge_mode, pdrop, training, seed, mask1, mask2, eps):
        res = extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=edge_mode, pdrop=pdrop, training=training,
                                seed=seed, mask1=mask1, mask2=mask2, eps=eps,
                                keep_h1=bool(ctx.needs_input_grad[3]))
        if res is None:
            raise ValueError('a graph exceeds one 128-row tile; use the fp32 extractor path for this batch')
        logit, saved = res
        ctx.saved = saved
        ctx.cfg = (gi, edge_mode, float(pdrop), bool(training), int(seed), mask1, mask2)
        ctx.save_for_backward(emb, w1, w2, w3)
        ctx.has_b = (b1 is not None, b2 is not None, b3 is not None)
        return logit

    @staticmethod
    def backward(ctx, dlogit):
        emb, w1, w2, w3 = ctx.saved_tensors
        gi, edge_mode, pdrop, training, seed, mask1, mask2 = ctx.cfg
        sv = ctx.saved
        tile_row, tile_seg, T = sv['plan']
        L = lib()
        dev = emb.device
        N, H = emb.shape
        C1 = w1.shape[0]
        Kin = w1.shape[1]
        rows = gi.E if edge_mode else gi.N
        seg_ptr = gi.edge_ptr if edge_mode else gi.node_ptr
        dlogit = dlogit.contiguous().view(-1).float()
        w3f = w3.detach().reshape(-1).contiguous()
        dz2 = torch.empty((rows, H), dtype=torch.bfloat16, device=dev)
        dw3_part = torch.empty((gi.G, H), dtype=torch.float32, device=dev)
        L.call('gsatb_tc_ext_bwd_head', ptr(dlogit), ptr(sv['xhat2']), ptr(sv['rstd2']), ptr(w3f), ptr(seg_ptr),
               ptr(mask2), ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training), ptr(dz2), ptr(dw3_part), rows,
               gi.G, H, stream())
        dz1 = torch.empty((rows, C1), dtype=torch.bfloat16, device=dev)
        w2t = prep_weight(w2, transpose=True)         # [C1, H]: A operand of dh1 = dz2 W2
        L.call('gsatb_tc_ext_bwd1', ptr(dz2), ptr(w2t), ptr(sv['xhat1']), ptr(sv['rstd1']), ptr(mask1),
               ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training), ptr(tile_row), ptr(tile_seg), ptr(seg_ptr),
               T, ptr(dz1), rows, H, C1, stream())
        # weight gradients (library GEMMs, fp32 accumulate) on re-materialised bf16 operands
        h1 = sv.get('h1')
        if h1 is None:
            h1 = torch.empty((rows, C1), dtype=torch.bfloat16, device=dev)
            L.call('gsatb_tc_ext_make_h1', ptr(sv['xhat1']), ptr(mask1), ctypes.c_uint64(seed), ctypes.c_float(pdrop),
                   int(training), ptr(h1), rows, C1, stream())
        dW2, _ = weight_grad(dz2, False, h1, False, rows, H, C1)
        del h1
        sv['h1'] = None
        f12 = sv.get('f12')
        if f12 is None:
            f12 = torch.empty((rows, Kin), dtype=torch.bfloat16, device=dev)
            L.call('gsatb_tc_ext_make_f12', ptr(emb), ptr(gi.src) if edge_mode else None,
                   ptr(gi.dst) if edge_mode else None, ptr(f12), rows, H, stream())
        dW1, _ = weight_grad(dz1, False, f12, False, rows, C1, Kin)
        del f12
        sv['f12'] = None
        # input gradient: d f12 = dz1 W1, then the deterministic scatter back to the nodes
        w1t = prep_weight(w1, transpose=True)         # [Kin, C1]
        df = torch.empty((rows, Kin), dtype=torch.float32, device=dev)
        L.call('gsatb_tc_linear_bf16in_fwd', ptr(dz1), C1, ptr(w1t), ptr(df), Kin, rows, C1, Kin, stream())
        if edge_mode:
            demb = torch.empty((N, H), dtype=torch.float32, device=dev)
            L.call('gsatb_gather_concat_bwd', ptr(df), ptr(gi.rowptr_src), ptr(gi.eid_by_src), ptr(gi.rowptr_dst),
                   ptr(gi.eid_by_dst), ptr(demb), N, H, stream())
        else:
            demb = df
        dw3 = dw3_part.sum(0).view_as(w3)
        db3 = dlogit.sum().view(1) if ctx.has_b[2] else None
        db1 = torch.zeros(C1, device=dev) if ctx.has_b[0] else None
        db2 = torch.zeros(H, device=dev) if ctx.has_b[1] else None
        return (demb, dW1, db1, dW2, db2, dw3, db3, None, None, None, None, None, None, None, None)


def weight_grad(a16: torch.Tensor, a_layout, b16: torch.Tensor, b_layout, rows: int, M: int, N: int,
                want_bias: bool = False):
    """dW [M, N] = sum_r A[r, m] B[r, n] (and db [M] = sum_r A[r, m]) on the tensor cores (gsatb_tc_dw: split-K tcgen05
    GEMM with a fixed-order reduction).  A / B are bf16 in layout 0 = row-major [rows, C], 1 = channel-major [C, rows]
    (leading dimension = stride(0)) or 2 = tile-major [rows / 128, Cpad, 128] (leading dimension = Cpad)."""
    dev = a16.device
    dW = torch.empty((M, N), dtype=torch.float32, device=dev)
    db = torch.empty(M, dtype=torch.float32, device=dev) if want_bias else None
    nb = int(lib().cdll.gsatb_tc_dw_workspace(rows, M, N))
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
    ld = lambda t, lay: int(t.shape[1]) if int(lay) == 2 else int(t.stride(0))
    lib().call('gsatb_tc_dw', ptr(a16), int(a_layout), ld(a16, a_layout), ptr(b16), int(b_layout), ld(b16, b_layout), rows, M, N,
               ptr(dW), N, ptr(db), 0, ptr(ws), ctypes.c_size_t(nb), stream())
    return dW, db


def ext_tile_slots(H: int, edge_mode: bool) -> int:
    return int(lib().cdll.gsatb_ext_tile_slots(int(H), int(edge_mode)))


def fused_extractor_supported(emb: torch.Tensor, gi: GraphIndex, edge_mode: bool) -> bool:
    """The fused tcgen05 extractor takes hidden widths H % 8 == 0, H <= 128 and batches whose graphs fit one tile
    (<= 128 rows, <= 112 when 2H > 128)."""
    H = emb.shape[1]
    if H % 8 != 0 or H > 128:
        return False
    return gi.ext_plan('edge' if edge_mode else 'node', ext_tile_slots(H, edge_mode))['oversize'] == 0


class _FusedExtractorV2(torch.autograd.Function):
    """The whole extractor MLP as ONE persistent tcgen05 kernel per direction (csrc/ext_fused_fwd.cu, ext_fused_bwd.cu):
    forward keeps nothing of width 4H in HBM (saved for backward: the logits' inputs xhat2 (bf16, tile-major slot space), rstd2, the
    centred input tiles xs and the effective dropout seeds); backward recomputes GEMM1, leaves the bf16 operands of the
    weight-gradient products in tile-major slot space ([tiles, channels, 128 slots]) and gsatb_tc_dw turns them into dW1 / dW2; the node gradient is
    the deterministic CSR reduction of d f12.  b1 / b2 sit in front of an InstanceNorm and get exact zeros."""

    @staticmethod
    def forward(ctx, emb, w1, b1, w2, b2, w3, b3, gi, edge_mode, pdrop, training, seed, mask1, mask2, eps):
        emb = emb.contiguous()
        N, H = emb.shape
        C1 = w1.shape[0]
        ms = ext_tile_slots(H, edge_mode)
        plan = gi.ext_plan('edge' if edge_mode else 'node', ms)
        if plan['oversize']:
            raise ValueError(f"{plan['oversize']} graph(s) exceed one {ms}-row tile of the fused extractor")
        rows, T = plan['rows'], plan['T']
        dev = emb.device
        need_grad = any(ctx.needs_input_grad[:7])
        ld = T * 128
        Kin = 2 * H if edge_mode else H
        logit = torch.empty((rows, 1), dtype=torch.float32, device=dev)
        # tile-major slot space; the kernel writes channels < H only, and the backward's TMA boxes cover all pad128(H)
        # channel rows of a block: the padding rows must read as zero (they meet zero weight columns in dh1 = W2^T dz2)
        xh2t = (torch.empty if H % 128 == 0 else torch.zeros)((T, _pad(H, 128), 128), dtype=torch.bfloat16, device=dev) \
            if need_grad else None
        rstd2 = torch.empty((max(gi.G, 1), H), dtype=torch.float32, device=dev) if need_grad else None
        xs = _xs_buffer(plan, ld, _pad(Kin, 64), dev) if need_grad else None
        token = object()
        if need_grad:
            plan['xs_token'] = token      # the dump is per batch: a second forward on it invalidates this one's backward
        seeds = torch.zeros(2, dtype=torch.int32, device=dev)
        w1p, w2p = prep_weight(w1), prep_weight(w2)
        w3f = w3.detach().reshape(-1).contiguous()
        e = bool(edge_mode)
        lib().call('gsatb_ext_fused_fwd', ptr(emb), ptr(gi.src) if e else None, ptr(gi.dst) if e else None,
                   ptr(gi.node_ptr) if e else None, ptr(gi.rowptr_src) if e else None, ptr(gi.rowptr_dst) if e else None,
                   ptr(plan['seg_ptr']), ptr(plan['tile_seg']), ptr(plan['out2']), max(gi.G, 1), ms, ptr(w1p), ptr(w2p),
                   ptr(w3f), ptr(b3), ptr(mask1), ptr(mask2), ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training),
                   ptr(logit), ptr(xh2t), ld, ptr(rstd2), ptr(xs), ptr(seeds), rows, H, C1, ctypes.c_float(eps), stream())
        ctx.cfg = (gi, e, float(pdrop), bool(training), mask1, mask2, float(eps), plan, ms, token)
        ctx.save_for_backward(w1, w2, w3f, xh2t, rstd2, xs, seeds)
        ctx.has_b = (b1 is not None, b2 is not None, b3 is not None)
        ctx.emb_shape = (N, H)
        return logit

    @staticmethod
    def backward(ctx, dlogit):
        w1, w2, w3f, xh2t, rstd2, xs, seeds = ctx.saved_tensors
        gi, e, pdrop, training, mask1, mask2, eps, plan, ms, token = ctx.cfg
        if plan.get('xs_token') is not token:
            raise RuntimeError('the fused extractor ran forward again on this batch before this backward: its saved input '
                               'tiles (one buffer per batch) were overwritten')
        N, H = ctx.emb_shape
        C1, Kin = w1.shape
        rows, T = plan['rows'], plan['T']
        ld = T * 128
        dev = dlogit.device
        L = lib()
        dl = dlogit.contiguous().view(-1).float()
        bf = dict(dtype=torch.bfloat16, device=dev)
        HP, C1P = _pad(H, 128), _pad(C1, 128)
        dz2t, dz1t, h1t = torch.empty((T, HP, 128), **bf), torch.empty((T, C1P, 128), **bf), torch.empty((T, C1P, 128), **bf)
        # edge mode: d f12 [E, 2H] is the step's largest intermediate (10 GB at cfg4 in fp32); it leaves the kernel in bf16
        # and the CSR reduction into d emb accumulates it in fp32.  Node mode: d f12 IS d emb (fp32).
        df12 = torch.empty((rows, Kin), dtype=torch.bfloat16 if e else torch.float32, device=dev)
        nslab = 2 * min(max(gi.G, 1), 148)
        dw3p = torch.zeros((nslab, H), dtype=torch.float32, device=dev)
        w1p, w2t, w1t = prep_weight(w1), prep_weight(w2, transpose=True), prep_weight(w1, transpose=True)
        L.call('gsatb_ext_fused_bwd', ptr(plan['seg_ptr']), ptr(plan['tile_seg']), ptr(plan['out2']), max(gi.G, 1), ms, int(e),
               ptr(w1p), ptr(w2t), ptr(w1t), ptr(w3f), ptr(dl), ptr(xh2t), ptr(rstd2), ptr(xs), ptr(mask1), ptr(mask2),
               ptr(seeds), ctypes.c_float(pdrop), int(training), ptr(dz2t), ptr(dz1t), ptr(h1t), ptr(df12), int(bool(e)), ptr(dw3p),
               ld, rows, H, C1, ctypes.c_float(eps), stream())
        dW2, _ = weight_grad(dz2t, 2, h1t, 2, ld, H, C1)
        dW1, _ = weight_grad(dz1t, 2, xs, 0, ld, C1, Kin)
        del dz1t, h1t, dz2t
        if e:
            demb = torch.empty((N, H), dtype=torch.float32, device=dev)
            L.call('gsatb_gather_concat_bwd_bf16', ptr(df12), ptr(gi.rowptr_src), ptr(gi.eid_by_src), ptr(gi.rowptr_dst),
                   ptr(gi.eid_by_dst), ptr(demb), N, H, stream())
        else:
            demb = df12
        dw3 = dw3p.sum(0).view(1, H)
        db3 = dl.sum().view(1) if ctx.has_b[2] else None
        db1 = torch.zeros(C1, device=dev) if ctx.has_b[0] else None
        db2 = torch.zeros(H, device=dev) if ctx.has_b[1] else None
        return (demb, dW1, db1, dW2, db2, dw3, db3, None, None, None, None, None, None, None, None)


def _xs_buffer(plan: dict, rows: int, ldx: int, dev) -> torch.Tensor:
    """The centred-input dump of the fused forward, [tiles * 128, pad64(Kin)] bf16, cached with the tile plan and re-used by
    every step on this batch.  The kernel writes rows [128 t, 128 t + max_slots) of every tile; the remaining rows of a
    tile are never written and meet zero columns of dz1 in dW1, so they only have to be finite: they are zeroed ONCE here
    (a strided fill of 128 - max_slots rows per tile, not a memset of the whole buffer)."""
    key = ('xs', ldx)
    buf = plan.get(key)
    if buf is None or buf.shape[0] != rows or buf.device != dev:
        buf = torch.empty((rows, ldx), dtype=torch.bfloat16, device=dev)
        ms = int(plan['max_slots'])
        if ms < 128:
            buf.view(rows // 128, 128, ldx)[:, ms:, :].zero_()
        plan[key] = buf
    return buf


def fused_extractor_v2(emb, w1, b1, w2, b2, w3, b3, gi, *, edge_mode: bool, pdrop: float, training: bool, seed: int,
                       mask1=None, mask2=None, eps: float = 1e-5):
    return _FusedExtractorV2.apply(emb, w1, b1, w2, b2, w3, b3, gi, edge_mode, pdrop, training, seed, mask1, mask2, eps)


def fused_extractor(emb, w1, b1, w2, b2, w3, b3, gi, *, edge_mode: bool, pdrop: float, training: bool, seed: int,
                    mask1=None, mask2=None, eps: float = 1e-5):
    return _FusedExtractor.apply(emb, w1, b1, w2, b2, w3, b3, gi, edge_mode, pdrop, training, seed, mask1, mask2, eps)


def _allreduce_stats(stats: torch.Tensor, n_local: int, group):
    """Sum per-channel statistics and the row count over the data-parallel group (sync BatchNorm).
    Returns the global row count as a 0-dim fp64 device tensor (no host sync: CUDA-graph capturable); ``stats`` is
    updated in place.  group None -> (stats untouched, python int)."""
    if group is None:
        return n_local
    import torch.distributed as dist
    buf = torch.empty(stats.numel() + 1, dtype=torch.float64, device=stats.device)
    buf[:-1] = stats
    buf[-1:].fill_(float(n_local))
    dist.all_reduce(buf, group=group)
    stats.copy_(buf[:-1])
    return buf[-1].clone()


def _rows_path(K: int, H1: int, H: int) -> bool:
    """The row-owner kernels of csrc/gin_rows.cu serve the node MLP when all three widths are equal and 64 or 128 (every
    GIN layer of the reference, src/models/gin.py:28-35); other shapes take the channel-owner skeleton kernels."""
    return bool(lib().cdll.gsatb_gin_rows_supported(int(K), int(H1), int(H)))


def _rows_lin1(agg16, w1p, b1, H1: int, want_stats: bool):
    N = agg16.shape[0]
    dev = agg16.device
    z1 = torch.empty((N, H1), dtype=torch.bfloat16, device=dev)
    L = lib()
    if not want_stats:
        L.call('gsatb_gin_rows_lin1', ptr(agg16), ptr(w1p), ptr(b1), ptr(z1), None, None, N, H1, stream())
        return z1
    part = torch.empty(int(L.cdll.gsatb_tc_stat_partials_elems(H1)), dtype=torch.float32, device=dev)
    stats = torch.empty(2 * H1, dtype=torch.float64, device=dev)
    L.call('gsatb_gin_rows_lin1', ptr(agg16), ptr(w1p), ptr(b1), ptr(z1), ptr(part), ptr(stats), N, H1, stream())
    return z1, stats


def _gin_mlp_forward(agg16, w1, b1, gamma, beta, w2, b2, running_mean, running_var, nbt, training, momentum, eps, pdrop,
                     drop_seed, drop_mask, keep_sign: bool = True, sync_group=None):
    """Dropout(ReLU(Linear2(ReLU(BatchNorm1d(Linear1(agg)))))) on bf16 ``agg16`` [N, K].  Linear1 is a TMA-fed tcgen05
    GEMM whose epilogue accumulates the BatchNorm batch statistics from the fp32 accumulators and stores z1 as bf16;
    BatchNorm + ReLU are folded into Linear2's operand load, the outer ReLU and the dropout into its epilogue."""
    N = agg16.shape[0]
    H1, H = w1.shape[0], w2.shape[0]
    dev = agg16.device
    w1p, w2p = prep_weight(w1), prep_weight(w2)
    mean, rstd, scale, shift = (torch.empty(H1, dtype=torch.float32, device=dev) for _ in range(4))
    rows_path = _rows_path(agg16.shape[1], H1, H)
    if training:
        z1, stats = _rows_lin1(agg16, w1p, b1, H1, True) if rows_path else linear_bf16(agg16, w1p, b1, H1, want_stats=True)
        n_rows = _allreduce_stats(stats, N, sync_group)          # python int, or the group-wide count on the device
        n_dev = n_rows if torch.is_tensor(n_rows) else None
        # batch statistics -> (mean, rstd, scale, shift) + running-statistics update, one launch (gsatb_bn_fold_fwd)
        lib().call('gsatb_bn_fold_fwd', ptr(stats), ctypes.c_double(float(N)), ptr(n_dev), ptr(gamma), ptr(beta),
                   ctypes.c_float(eps), ctypes.c_float(momentum), ptr(running_mean), ptr(running_var), ptr(nbt), 1, ptr(mean),
                   ptr(rstd), ptr(scale), ptr(shift), H1, stream())
    else:
        z1 = _rows_lin1(agg16, w1p, b1, H1, False) if rows_path else linear_bf16(agg16, w1p, b1, H1)
        lib().call('gsatb_bn_fold_fwd', None, ctypes.c_double(1.0), None, ptr(gamma), ptr(beta), ctypes.c_float(eps),
                   ctypes.c_float(momentum), ptr(running_mean), ptr(running_var), None, 0, ptr(mean), ptr(rstd), ptr(scale),
                   ptr(shift), H1, stream())
    p = float(pdrop) if training else 0.0
    a1 = torch.empty_like(z1)
    # sign bits of h: all the backward pass needs of the layer output (4 bytes per 32 channels instead of 128)
    posmask = torch.empty((N, (H + 31) // 32), dtype=torch.int32, device=agg16.device) if keep_sign else None
    if rows_path:
        # BatchNorm + ReLU are applied to the z1 tile in shared memory on its way into the second GEMM; a1 leaves as the
        # bf16 operand of dW2 from the same tile
        h = torch.empty((N, H), dtype=torch.float32, device=dev)
        mk = None if drop_mask is None else drop_mask.to(torch.uint8).contiguous()
        lib().call('gsatb_gin_rows_lin2', ptr(z1), ptr(scale), ptr(shift), ptr(w2p), ptr(b2), ptr(a1), ptr(h), ptr(posmask),
                   ptr(mk), ctypes.c_uint64(int(drop_seed) & (2 ** 64 - 1)), ctypes.c_float(p), N, H, stream())
        return h, (z1, a1, mean, rstd, scale, shift, p, posmask)
    lib().call('gsatb_bn_relu_bf16', ptr(z1), ptr(scale), ptr(shift), ptr(a1), N, H1, stream())
    h = linear_bf16(a1, w2p, b2, H, out_bf16=False, relu_out=True, pdrop=p, drop_seed=drop_seed, drop_mask=drop_mask,
                    posmask=posmask)
    return h, (z1, a1, mean, rstd, scale, shift, p, posmask)


def _dropout_scale(p: float, injected_mask: bool) -> float:
    """The scale the forward kernels apply to kept values (tc_ops_common.cuh make_dropout): with an injected mask the
    reference's 1/(1-p); with the in-kernel word scheme p is quantised to thr8 = round(256 p)/256 and the kept values
    are scaled by the exact keep rate 256/(256 - thr8).  Backward must use the SAME number."""
    if p <= 0:
        return 1.0
    if injected_mask:
        return 1.0 / (1.0 - p)
    thr8 = min(256, int(p * 256.0 + 0.5))
    return 256.0 / (256 - thr8) if thr8 < 256 else 1.0 / (1.0 - p)


def _gin_mlp_backward(dh, agg16, h, w1, w2, gamma, z1, a1, mean, rstd, scale, shift, training, p, posmask=None,
                      sync_group=None, injected_mask: bool = False):
    """-> (d agg fp32 [N, K], dW1, db1, dgamma, dbeta, dW2, db2).  With ``sync_group`` the BatchNorm backward sums span
    the group (the returned dgamma / dbeta stay the LOCAL sums: the step's gradient all-reduce adds them up)."""
    N, Kin = agg16.shape
    H1, H = w1.shape[0], w2.shape[0]
    dev = agg16.device
    L = lib()
    dh = dh.contiguous()
    d2 = torch.empty((N, H), dtype=torch.bfloat16, device=dev)
    g = torch.empty((N, H1), dtype=torch.bfloat16, device=dev)
    part = torch.empty(int(L.cdll.gsatb_tc_stat_partials_elems(H1)), dtype=torch.float32, device=dev)
    stats = torch.empty(2 * H1, dtype=torch.float32, device=dev)
    # (operands are held in locals until the launch has been enqueued: a temporary passed as ptr(f(x)) is released
    # before the call is made -- harmless with the stream-ordered caching allocator, but not something to lean on)
    w2t = prep_weight(w2, transpose=True)
    if posmask is not None and _rows_path(Kin, H1, H):
        part = torch.empty(int(L.cdll.gsatb_gin_rows_stat_partials_elems(H1)), dtype=torch.float32, device=dev)
        L.call('gsatb_gin_rows_bwd2', ptr(dh), ptr(posmask), ctypes.c_float(_dropout_scale(p, injected_mask)), ptr(w2t), ptr(z1),
               ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(d2), ptr(g), ptr(part), ptr(stats), N, H, stream())
    else:
        L.call('gsatb_tc_gin_bwd2', ptr(dh), None if posmask is not None else ptr(h), ptr(posmask),
               ctypes.c_float(_dropout_scale(p, injected_mask)),
               ptr(w2t), ptr(z1), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(d2),
               ptr(g), None, ptr(part), ptr(stats), N, H, H1, stream())      # a1 was kept by the forward
    dbeta, dgamma = stats[:H1], stats[H1:]
    n_glob = None
    if training and sync_group is not None:              # sums and row count over every rank's rows
        dbeta, dgamma = dbeta.clone(), dgamma.clone()                           # local sums -> parameter gradients
        n_glob = _allreduce_stats(stats, N, sync_group)
    # dz1 = cA * g + cB * z1 + cC (BatchNorm backward folded into three per-channel vectors), one launch
    cA, cB, cC = (torch.empty(H1, dtype=torch.float32, device=dev) for _ in range(3))
    L.call('gsatb_bn_fold_bwd', ptr(stats[:H1]), ptr(stats[H1:]), ctypes.c_double(float(max(N, 1))), ptr(n_glob), ptr(gamma),
           ptr(mean), ptr(rstd), int(bool(training)), ptr(cA), ptr(cB), ptr(cC), H1, stream())
    dz1 = torch.empty((N, H1), dtype=torch.bfloat16, device=dev)
    dagg = torch.empty((N, Kin), dtype=torch.float32, device=dev)
    w1t = prep_weight(w1, transpose=True)
    if _rows_path(Kin, H1, H):
        L.call('gsatb_gin_rows_bwd1', ptr(g), ptr(z1), ptr(cA), ptr(cB), ptr(cC), ptr(w1t), ptr(dz1), ptr(dagg), N, H1, stream())
    else:
        L.call('gsatb_tc_gin_bwd1', ptr(g), ptr(z1), ptr(cA), ptr(cB), ptr(cC), ptr(w1t), ptr(dz1), ptr(dagg), N, H1, Kin,
               stream())
    dW2, db2 = weight_grad(d2, False, a1, False, N, H, H1, want_bias=True)
    dW1, db1 = weight_grad(dz1, False, agg16, False, N, H1, Kin, want_bias=True)
    return dagg, dW1, db1, dgamma, dbeta, dW2, db2


class _GinMlpFused(torch.autograd.Function):
    """GIN node MLP  Dropout(ReLU(Linear2(ReLU(BatchNorm1d(Linear1(x))))))  (reference src/models/gin.py:55-62 + the
    ReLU / Dropout of :50-52) on tcgen05, for an fp32 input ``x`` (rounded to bf16 operands here).  Weight gradients
    are plain library GEMMs on the bf16 operands the kernels write out."""

    @staticmethod
    def forward(ctx, x, w1, b1, gamma, beta, w2, b2, running_mean, running_var, nbt, training, momentum, eps, pdrop,
                drop_seed, drop_mask, sync_group=None):
        agg16 = x.contiguous().bfloat16()
        h, (z1, a1, mean, rstd, scale, shift, p, posmask) = _gin_mlp_forward(
            agg16, w1, b1, gamma, beta, w2, b2, running_mean, running_var, nbt, training, momentum, eps, pdrop,
            drop_seed, drop_mask, sync_group=sync_group)
        ctx.save_for_backward(agg16, z1, a1, posmask, w1, w2, gamma, mean, rstd, scale, shift)
        ctx.cfg = (bool(training), p, sync_group, drop_mask is not None)
        return h

    @staticmethod
    def backward(ctx, dh):
        agg16, z1, a1, posmask, w1, w2, gamma, mean, rstd, scale, shift = ctx.saved_tensors
        training, p, sync_group, injected = ctx.cfg
        dagg, dW1, db1, dgamma, dbeta, dW2, db2 = _gin_mlp_backward(dh, agg16, None, w1, w2, gamma, z1, a1, mean, rstd,
                                                                    scale, shift, training, p, posmask, sync_group,
                                                                    injected)
        return (dagg, dW1, db1, dgamma, dbeta, dW2, db2) + (None,) * 10


def gin_mlp_relu(x, seq, training: bool, pdrop: float = 0.0, drop_seed: int = 0, drop_mask=None):
    """``dropout(relu(seq(x)))`` for seq = GIN.MLP(...) = Sequential(Linear, BatchNorm1d, ReLU, Linear)."""
    lin1, bn, _, lin2 = seq[0], seq[1], seq[2], seq[3]
    return _GinMlpFused.apply(x, lin1.weight, lin1.bias, bn.weight, bn.bias, lin2.weight, lin2.bias, bn.running_mean,
                              bn.running_var, bn.num_batches_tracked, training and bn.training,
                              bn.momentum if bn.momentum is not None else 0.1, bn.eps, pdrop if training else 0.0,
                              drop_seed, drop_mask, getattr(bn, 'sync_group', None))


class _GinLayerFused(torch.autograd.Function):
    """One whole attention-aware GIN layer of the reference (GINConv.forward, src/models/conv_layers.py:14-34, with
    its nn = GIN.MLP, then the ReLU / Dropout of src/models/gin.py:50-52):  K3 aggregation written as bf16 straight
    into the operand layout of the node MLP -> the tcgen05 MLP above.  Backward chains the MLP backward (d agg in
    fp32) into K3's backward (dx, d edge_atten)."""

    @staticmethod
    def forward(ctx, x, att, w1, b1, gamma, beta, w2, b2, running_mean, running_var, nbt, gi, conv_eps, training,
                momentum, eps, pdrop, drop_seed, drop_mask, sync_group=None):
        x = x.contiguous()
        N, K = x.shape
        att_flat = None if att is None else att.contiguous().view(-1)
        agg16 = torch.empty((N, K), dtype=torch.bfloat16, device=x.device)
        lib().call('gsatb_gin_aggregate_fwd_bf16', ptr(x), ptr(att_flat), ptr(gi.rowptr_dst), ptr(gi.eid_by_dst),
                   ptr(gi.src_by_dst), ctypes.c_float(conv_eps), ptr(agg16), N, gi.E, K, stream())
        h, (z1, a1, mean, rstd, scale, shift, p, posmask) = _gin_mlp_forward(
            agg16, w1, b1, gamma, beta, w2, b2, running_mean, running_var, nbt, training, momentum, eps, pdrop,
            drop_seed, drop_mask, sync_group=sync_group)
        ctx.save_for_backward(x, att_flat, agg16, z1, a1, posmask, w1, w2, gamma, mean, rstd, scale, shift)
        ctx.cfg = (bool(training), p, gi, float(conv_eps), None if att is None else att.shape, sync_group,
                   drop_mask is not None)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, att_flat, agg16, z1, a1, posmask, w1, w2, gamma, mean, rstd, scale, shift = ctx.saved_tensors
        training, p, gi, conv_eps, att_shape, sync_group, injected = ctx.cfg
        dagg, dW1, db1, dgamma, dbeta, dW2, db2 = _gin_mlp_backward(dh, agg16, None, w1, w2, gamma, z1, a1, mean, rstd,
                                                                    scale, shift, training, p, posmask, sync_group,
                                                                    injected)
        N, K = x.shape
        need_att = att_flat is not None and ctx.needs_input_grad[1]
        dx = torch.empty_like(x)
        datt = torch.empty(gi.E, dtype=torch.float32, device=x.device) if need_att else None
        lib().call('gsatb_gin_aggregate_bwd', ptr(dagg), ptr(x), ptr(att_flat), ptr(gi.rowptr_src), ptr(gi.eid_by_src),
                   ptr(gi.dst_by_src), ctypes.c_float(conv_eps), ptr(dx), ptr(datt), N, gi.E, K, stream())
        return (dx, datt.view(att_shape) if need_att else None, dW1, db1, dgamma, dbeta, dW2, db2) + (None,) * 12


def gin_layer(x, edge_atten, gi, conv, training: bool, pdrop: float = 0.0, drop_seed: int = 0, drop_mask=None):
    """``dropout(relu(conv(x, edge_index, edge_atten=edge_atten)))`` for conv = GINConv(GIN.MLP(...))."""
    lin1, bn, _, lin2 = conv.nn[0], conv.nn[1], conv.nn[2], conv.nn[3]
    return _GinLayerFused.apply(x, edge_atten, lin1.weight, lin1.bias, bn.weight, bn.bias, lin2.weight, lin2.bias,
                                bn.running_mean, bn.running_var, bn.num_batches_tracked, gi, conv.initial_eps,
                                training and bn.training, bn.momentum if bn.momentum is not None else 0.1, bn.eps,
                                pdrop if training else 0.0, drop_seed, drop_mask, getattr(bn, 'sync_group', None))
