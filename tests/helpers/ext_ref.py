"""Plain-PyTorch restatement of the extractor MLP (reference src/run_gsat.py:909-927 + src/utils/get_model.py:57-68,
PyG InstanceNorm per SURVEY App. A.3) used by the tests of the fused tensor-core kernels.  TEST INFRASTRUCTURE ONLY.

`rounding='bf16'` rounds to bf16 at exactly the points where the fused kernels do (the per-graph-centred input rows,
the weights, the hidden activation h1 that feeds GEMM2), keeping everything else in the working dtype, so that the
kernels can be held to a tight tolerance; `rounding=None` is the reference's own arithmetic (fp32 or fp64)."""
import torch


def _bf(t, rounding):
    return t.bfloat16().to(t.dtype) if rounding == 'bf16' else t


def seg_mean(x, seg_ptr):
    G = seg_ptr.numel() - 1
    cnt = (seg_ptr[1:] - seg_ptr[:-1]).clamp(min=1).to(x.dtype)
    ids = torch.repeat_interleave(torch.arange(G, device=x.device), (seg_ptr[1:] - seg_ptr[:-1]).long())
    s = torch.zeros(G, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, ids, x)
    return s / cnt[:, None], ids


def instance_norm(x, seg_ptr, eps=1e-5):
    m, ids = seg_mean(x, seg_ptr)
    xc = x - m[ids]
    v, _ = seg_mean(xc * xc, seg_ptr)
    return xc / torch.sqrt(v + eps)[ids]


def extractor_forward(emb, src, dst, seg_ptr, w1, w2, w3, b3, mask1=None, mask2=None, pdrop=0.0, rounding=None,
                      eps=1e-5, want_all=False):
    """emb [N,H]; src/dst int64 [E] or None (node mode); seg_ptr int64 [G+1]; masks are 0/1 tensors or None.
    Returns logit [rows, 1] (and the intermediates when want_all)."""
    x = torch.cat([emb[src], emb[dst]], dim=1) if src is not None else emb
    m, ids = seg_mean(x, seg_ptr)
    xc = _bf(x - m[ids], rounding)                      # Linear is linear: W (x - mean) = z - mean(z); the bias cancels
    z1 = xc @ _bf(w1, rounding).t()
    v1, _ = seg_mean(z1 * z1, seg_ptr)
    xh1 = z1 / torch.sqrt(v1 + eps)[ids]
    h1 = torch.relu(xh1)
    scale = 1.0 / (1.0 - pdrop) if pdrop > 0 else 1.0
    if mask1 is not None:
        h1 = h1 * mask1.to(h1.dtype) * scale
    h1 = _bf(h1, rounding)
    z2 = h1 @ _bf(w2, rounding).t()
    xh2 = instance_norm(z2, seg_ptr, eps)
    h2 = torch.relu(xh2)
    if mask2 is not None:
        h2 = h2 * mask2.to(h2.dtype) * scale
    logit = h2 @ w3.reshape(-1, 1) + (b3 if b3 is not None else 0.0)
    if want_all:
        return logit, dict(xc=xc, z1=z1, xh1=xh1, h1=h1, z2=z2, xh2=xh2, h2=h2)
    return logit


def extractor_backward_emulated(emb, src, dst, seg_ptr, w1, w2, w3, dlogit, mask1=None, mask2=None, pdrop=0.0, eps=1e-5):
    """The fused backward kernel's arithmetic restated in plain PyTorch with bf16 rounding at exactly its rounding
    points (centred input rows, weights, h1, the saved xhat2, dz2, dz1): what the kernel must reproduce to ~1e-3.
    Returns (d f12 [rows, Kin], dW1, dW2, dw3)."""
    bf = lambda t: t.bfloat16().to(t.dtype)
    x = torch.cat([emb[src], emb[dst]], dim=1) if src is not None else emb
    m, ids = seg_mean(x, seg_ptr)
    n = (seg_ptr[1:] - seg_ptr[:-1]).clamp(min=1).to(x.dtype)[ids][:, None]
    xc = bf(x - m[ids])
    W1, W2 = bf(w1), bf(w2)
    z1 = xc @ W1.t()
    v1, _ = seg_mean(z1 * z1, seg_ptr)
    r1 = 1.0 / torch.sqrt(v1 + eps)[ids]
    xh1 = z1 * r1
    s = 1.0 / (1.0 - pdrop) if pdrop > 0 else 1.0
    k1 = mask1.to(x.dtype) if mask1 is not None else torch.ones_like(z1)
    gate1 = (z1 > 0).to(x.dtype) * k1
    h1 = bf(xh1 * gate1 * s)
    z2 = h1 @ W2.t()
    mu2, _ = seg_mean(z2, seg_ptr)
    zc = z2 - mu2[ids]
    v2, _ = seg_mean(zc * zc, seg_ptr)
    r2 = 1.0 / torch.sqrt(v2 + eps)[ids]
    xh2 = bf(zc * r2)                                        # saved by the forward as bf16
    k2 = mask2.to(x.dtype) if mask2 is not None else torch.ones_like(z2)
    gate2 = (xh2 > 0).to(x.dtype) * k2
    dl = dlogit.reshape(-1, 1)
    dw3 = (dl * xh2 * gate2 * s).sum(0)
    g2 = dl * w3.reshape(1, -1) * s * gate2
    sum_over = lambda t: seg_mean(t, seg_ptr)[0][ids]       # per-graph mean, broadcast back to the rows
    dz2 = bf(r2 * (g2 - sum_over(g2) - xh2 * sum_over(g2 * xh2)))
    dh1 = dz2 @ W2
    dy = dh1 * s * gate1
    dz1 = bf(r1 * (dy - sum_over(dy) - xh1 * sum_over(dy * xh1)))
    df12 = dz1 @ W1
    return df12, dz1.t() @ xc, dz2.t() @ h1, dw3
