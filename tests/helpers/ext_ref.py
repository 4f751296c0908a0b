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
