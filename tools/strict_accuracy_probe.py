"""How accurate is the strict (split-bf16 x3) tensor-core product on the hardware?  Relative L2 error against fp64 of
y = x W^T, dx, dW for a few shapes, beside the error of torch's fp32 CPU product (the oracle's arithmetic) and of the
single-pass bf16 mode.  GPU tool; prints a table (committed as profiles/r2_strict_accuracy.txt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dp_gsat_b200 import dense


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


print(f'{"rows x K x OUT":>20s} | {"strict y":>9s} {"dx":>9s} {"dW":>9s} | {"fp32 CPU y":>10s} {"dx":>9s} {"dW":>9s} | {"bf16 y":>9s}')
for rows, K, OUT in [(4096, 64, 64), (4096, 128, 128), (4096, 256, 512), (4096, 512, 128), (4096, 640, 80), (100000, 128, 128)]:
    g = torch.Generator().manual_seed(K)
    x, w = torch.randn(rows, K, generator=g), torch.randn(OUT, K, generator=g) / K ** 0.5
    dy = torch.randn(rows, OUT, generator=g)
    r64 = [t.double().requires_grad_(True) for t in (x, w)]
    (r64[0] @ r64[1].t()).backward(dy.double())
    y64 = r64[0] @ r64[1].t()
    r32 = [t.clone().requires_grad_(True) for t in (x, w)]
    y32 = r32[0] @ r32[1].t()
    y32.backward(dy)
    got = [t.clone().cuda().requires_grad_(True) for t in (x, w)]
    y = dense.linear(got[0], got[1], None, 'fp32')
    y.backward(dy.cuda())
    yb = dense.linear(x.cuda(), w.cuda(), None, 'bf16')
    print(f'{f"{rows} x {K} x {OUT}":>20s} | {rel(y, y64):9.2e} {rel(got[0].grad, r64[0].grad):9.2e} {rel(got[1].grad, r64[1].grad):9.2e} | '
          f'{rel(y32, y64):10.2e} {rel(r32[0].grad, r64[0].grad):9.2e} {rel(r32[1].grad, r64[1].grad):9.2e} | {rel(yb, y64):9.2e}')
