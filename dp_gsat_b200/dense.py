"""Dense layers of the path on this library's tensor-core kernels, in both precision modes.

  reference layer                                                       -> here
  torch.nn.Linear  (gin.py:22-25,42,55-62; pna.py:20-50; get_model.py:57-68;
                    conv_layers.py:49,77-79,149)                         -> Linear (same parameters / state_dict keys)
  torch.nn.BatchNorm1d / torch_geometric BatchNorm (gin.py:59, pna.py:45) -> batch_norm (used by nn.BatchNorm1d)

Every product -- y = x W^T + b, dx = dy W, dW = dy^T x -- is ONE launch of the tcgen05 GEMM kernels
(gsatb_tc_linear_bf16_fwd, gsatb_tc_dw) on bf16 operands with fp32 accumulation in TMEM:

  precision 'bf16'   operands rounded to bf16 once (documented bf16 bound);
  precision 'fp32'   "split-bf16 x3" strict mode: every fp32 operand is the exact sum of three bf16 numbers and the six
                     significant partial products are summed by the same GEMM over six K-segments (csrc/dense.cu), so
                     the result carries fp32 accuracy (held to rtol 1e-5 against the fp32 oracle by the parity tests).

No library GEMM (cuBLAS) and no ATen batch-norm kernel is used in either mode."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as tnn

from ._lib import lib, ptr, stream

# 2 bits per K-segment (0 = h, 1 = m, 2 = l), segment 0 in the low bits
# smallest partial products first: l h + m m + h l + m h + h m + h h.  The fp32 accumulator then rounds the K large h h
# terms exactly as a plain fp32 dot product would, instead of rounding 5K small corrections at the ulp of the full sum.
_PARTS_A = (2, 1, 0, 1, 0, 0)      # the activation-side operand
_PARTS_B = (0, 1, 2, 0, 1, 0)      # the other operand
PATTERN_A = sum(p << (2 * i) for i, p in enumerate(_PARTS_A))
PATTERN_B = sum(p << (2 * i) for i, p in enumerate(_PARTS_B))
# 'bf16x2': two bf16 parts per operand (16 mantissa bits), three K-segments  m h + h m + h h  -- for layers whose
# gradients are small residuals of large cancelling terms (PNA post_nn in front of BatchNorm, see pna.py)
_MODES = {'bf16': (1, 0, 0), 'bf16x2': (3, 1 | (0 << 2) | (0 << 4), 0 | (1 << 2) | (0 << 4)), 'fp32': (6, PATTERN_A, PATTERN_B)}
_CHUNK_BYTES = 1 << 31             # operand bytes staged per row chunk (bounds the 6x expansion of the strict mode)
OUT_BLOCK = 512                    # output channels per GEMM launch (four 128-lane accumulators)


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def split_bf16(x: torch.Tensor, nseg: int, pattern: int, layout: int, out: Optional[torch.Tensor] = None,
               ld_out: Optional[int] = None) -> torch.Tensor:
    """fp32 [rows, C] -> bf16 GEMM operand (gsatb_split_bf16).  layout 0: [rows, pad8(nseg*C)], layout 1:
    [nseg*rows, pad8(C)]."""
    rows, C = x.shape
    if ld_out is None:
        ld_out = _pad(nseg * C if layout == 0 else C, 8)
    if out is None:
        out = torch.empty((rows if layout == 0 else nseg * rows, ld_out), dtype=torch.bfloat16, device=x.device)
    ldx = int(x.stride(0)) if rows > 1 else C          # (a one-row view may carry any stride in its size-1 dimension)
    lib().call('gsatb_split_bf16', ptr(x), rows, C, ldx, nseg, pattern, layout, ptr(out), ld_out, stream())
    return out


def _weight_operand(w: torch.Tensor, nseg: int, pattern: int):
    """fp32 [OUT, K] -> (zero-padded bf16 [pad128(OUT), pad64(Kc)] in the TMA box layout of the GEMM's A operand, Kc)."""
    OUT, K = w.shape
    Kc = _pad(nseg * K, 8)
    ldw = _pad(Kc, 64)
    wp = torch.zeros((_pad(OUT, 128), ldw), dtype=torch.bfloat16, device=w.device)
    split_bf16(w, nseg, pattern, 0, out=wp, ld_out=ldw)
    return wp, Kc


def _gemm(x16: torch.Tensor, Kc: int, wp: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, OUT: int):
    """out[:, :OUT] (fp32 view, row stride out.stride(0)) = x16[:, :Kc] W'^T + bias, 512 output channels per launch."""
    rows = x16.shape[0]
    for o0 in range(0, OUT, OUT_BLOCK):
        n = min(OUT_BLOCK, OUT - o0)
        lib().call('gsatb_tc_linear_bf16_fwd', ptr(x16), int(x16.stride(0)), ptr(wp[o0:]),
                   ptr(bias[o0:]) if bias is not None else None, ptr(out[:, o0:]), 0, int(out.stride(0)), 0, None, None,
                   None, ctypes.c_uint64(0), ctypes.c_float(0.0), None, rows, Kc, n, stream())


def _row_chunk(rows: int, width: int, nseg: int) -> int:
    per_row = max(1, nseg * width * 2)
    return max(128, min(rows, (_CHUNK_BYTES // per_row) // 128 * 128))


def colsum(x: torch.Tensor) -> torch.Tensor:
    rows, C = x.shape
    out = torch.empty(C, dtype=torch.float32, device=x.device)
    nb = int(lib().cdll.gsatb_col_workspace(rows, C))
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=x.device)
    lib().call('gsatb_colsum', ptr(x), rows, C, int(x.stride(0)), ptr(out), ptr(ws), ctypes.c_size_t(nb), stream())
    return out


def _mode(strict) -> tuple:
    """(segments, activation-side pattern, other-side pattern) of a precision mode ('bf16' | 'bf16x2' | 'fp32'; a bool
    means strict 'fp32' / plain 'bf16')."""
    if isinstance(strict, bool):
        strict = 'fp32' if strict else 'bf16'
    return _MODES[strict]


def linear_forward(x2: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], strict) -> torch.Tensor:
    """x2 [rows, K] fp32 (unit column stride), w [OUT, K] fp32 contiguous -> x2 w^T + b, no autograd bookkeeping."""
    rows, K = x2.shape
    OUT = w.shape[0]
    nseg, pat_a, pat_b = _mode(strict)
    out = torch.empty((rows, OUT), dtype=torch.float32, device=x2.device)
    if rows > 0:
        wp, Kc = _weight_operand(w, nseg, pat_b)
        step = _row_chunk(rows, K, nseg)
        for r0 in range(0, rows, step):
            xs = split_bf16(x2[r0:r0 + step], nseg, pat_a, 0)
            _gemm(xs, Kc, wp, b, out[r0:r0 + step], OUT)
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, strict):
        OUT, K = weight.shape
        lead = x.shape[:-1]
        x2 = x.reshape(-1, K)
        if x2.dtype != torch.float32 or x2.stride(-1) != 1:
            x2 = x2.float().contiguous()
        w = weight.detach().float().contiguous()
        b = None if bias is None else bias.detach().float().contiguous()
        out = linear_forward(x2, w, b, strict)
        ctx.save_for_backward(x2, w)
        ctx.cfg = (strict, lead, bias is not None)
        return out.view(*lead, OUT)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        strict, lead, has_bias = ctx.cfg
        OUT, K = w.shape
        rows = x2.shape[0]
        dy2 = dy.reshape(-1, OUT)
        if dy2.dtype != torch.float32 or not dy2.is_contiguous():
            dy2 = dy2.float().contiguous()
        nseg, pat_a, pat_b = _mode(strict)
        dev = dy2.device
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((rows, K), dtype=torch.float32, device=dev)
            if rows > 0:
                wtp, Oc = _weight_operand(w.t().contiguous(), nseg, pat_b)
                step = _row_chunk(rows, OUT, nseg)
                for r0 in range(0, rows, step):
                    ds = split_bf16(dy2[r0:r0 + step], nseg, pat_a, 0)
                    _gemm(ds, Oc, wtp, None, dx[r0:r0 + step], K)
            dx = dx.view(*lead, K)
        if ctx.needs_input_grad[1]:
            dW = torch.empty((OUT, K), dtype=torch.float32, device=dev)
            L = lib()
            step = _row_chunk(max(rows, 1), max(OUT, K), nseg)
            nb = int(L.cdll.gsatb_tc_dw_workspace(nseg * min(step, max(rows, 1)), OUT, K))
            ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
            if rows == 0:
                dW.zero_()
            for i, r0 in enumerate(range(0, rows, step)):
                a = split_bf16(dy2[r0:r0 + step], nseg, pat_a, 1)        # [nseg * n, pad8(OUT)]
                bq = split_bf16(x2[r0:r0 + step], nseg, pat_b, 1)        # [nseg * n, pad8(K)]
                L.call('gsatb_tc_dw', ptr(a), 0, int(a.stride(0)), ptr(bq), 0, int(bq.stride(0)), a.shape[0], OUT, K, ptr(dW),
                       K, None, int(i > 0), ptr(ws), ctypes.c_size_t(nb), stream())
        if has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        return dx, dW, db, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], precision: str = 'fp32') -> torch.Tensor:
    """F.linear(x, weight, bias) on the tcgen05 kernels; precision 'fp32' = split-bf16 x3 strict mode, 'bf16' = one pass,
    'bf16x2' = two bf16 parts per operand (three passes, 16 mantissa bits)."""
    if precision not in _MODES:
        raise ValueError(f'unknown precision {precision!r}')
    return _LinearFn.apply(x, weight, bias, precision)


class Linear(tnn.Linear):
    """torch.nn.Linear with the same parameters, initialisation and state_dict keys, computed by this library's kernels
    in the precision mode of its owner model (``precision``: 'fp32' strict / 'bf16')."""
    precision = 'fp32'
    bf16_mode = 'bf16'       # what precision 'bf16' means for THIS layer: 'bf16' (one pass) or 'bf16x2' (see pna.py)

    def forward(self, x):
        # raw input features (encoders, K < 16) and final logits (OUT < 8) are never rounded to bf16: those layers are a
        # negligible share of the FLOPs and always take the strict product
        narrow = self.in_features < 16 or self.out_features < 8
        mode = 'fp32' if (narrow or self.precision == 'fp32') else self.bf16_mode
        return linear(x, self.weight, self.bias, mode)


def set_precision(module: tnn.Module, precision: str) -> None:
    """Propagate a model's precision mode to the dense layers it owns."""
    if precision not in ('fp32', 'bf16'):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    for m in module.modules():
        if isinstance(m, Linear):
            m.precision = precision


class PrecisionMixin:
    """``model.precision = 'bf16' | 'fp32'`` on GIN / PNA / ExtractorMLP / SPMotifNet reaches every Linear underneath."""

    @property
    def precision(self) -> str:
        return self.__dict__.get('_precision', 'fp32')

    @precision.setter
    def precision(self, value: str) -> None:
        set_precision(self, value)
        self.__dict__['_precision'] = value


# ------------------------------------------------------------------------------------------------------------
# BatchNorm1d
# ------------------------------------------------------------------------------------------------------------
class _BatchNormFn(torch.autograd.Function):
    """BatchNorm1d over [rows, C] (+ optional fused ReLU) on gsatb_bn_* (csrc/dense.cu): training mode normalises with
    the biased batch variance and updates the running statistics with the unbiased one (momentum), eval mode uses the
    running statistics -- torch.nn.functional.batch_norm semantics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, relu):
        x = x.float().contiguous()
        rows, C = x.shape
        dev = x.device
        L = lib()
        if training:
            if rows == 0:
                raise ValueError('BatchNorm1d: empty batch in training mode')
            mean = torch.empty(C, dtype=torch.float32, device=dev)
            rstd = torch.empty(C, dtype=torch.float32, device=dev)
            nb = int(L.cdll.gsatb_col_workspace(rows, C))
            ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
            L.call('gsatb_bn_stats', ptr(x), rows, C, ctypes.c_float(eps), ctypes.c_float(momentum), ptr(running_mean),
                   ptr(running_var), ptr(mean), ptr(rstd), None, ptr(ws), ctypes.c_size_t(nb), stream())
        else:
            mean = running_mean.detach().float().contiguous()
            rstd = torch.rsqrt(running_var.detach().float() + eps)
        y = torch.empty_like(x)
        g = None if gamma is None else gamma.detach().float().contiguous()
        b = None if beta is None else beta.detach().float().contiguous()
        L.call('gsatb_bn_apply', ptr(x), ptr(mean), ptr(rstd), ptr(g), ptr(b), int(relu), ptr(y), rows, C, stream())
        ctx.save_for_backward(x, y if relu else None, mean, rstd, g)
        ctx.cfg = (bool(training), bool(relu), gamma is not None, beta is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, g = ctx.saved_tensors
        training, relu, has_g, has_b = ctx.cfg
        rows, C = x.shape
        dev = x.device
        L = lib()
        dy = dy.float().contiguous()
        dbeta = torch.empty(C, dtype=torch.float32, device=dev)
        dgamma = torch.empty(C, dtype=torch.float32, device=dev)
        nb = int(L.cdll.gsatb_col_workspace(rows, C))
        ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        if rows > 0:
            L.call('gsatb_bn_bwd_stats', ptr(dy), ptr(x), ptr(y), ptr(mean), ptr(rstd), rows, C, ptr(dbeta), ptr(dgamma), None,
                   ptr(ws), ctypes.c_size_t(nb), stream())
        else:
            dbeta.zero_()
            dgamma.zero_()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            L.call('gsatb_bn_bwd_apply', ptr(dy), ptr(x), ptr(y), ptr(mean), ptr(rstd), ptr(g), ptr(dbeta), ptr(dgamma),
                   ctypes.c_float(1.0 / max(rows, 1)), int(training), ptr(dx), rows, C, stream())
        return dx, (dgamma if has_g else None), (dbeta if has_b else None), None, None, None, None, None, None


def batch_norm(x, gamma, beta, running_mean, running_var, training: bool, momentum: float, eps: float,
               relu: bool = False) -> torch.Tensor:
    return _BatchNormFn.apply(x, gamma, beta, running_mean, running_var, training, momentum, eps, relu)
