"""Development probe (GPU): GIN node MLP (Linear -> BatchNorm1d -> ReLU -> Linear) backward on DEGENERATE rows (every row
a small integer multiple of one vector) -- this library's dense layers vs torch's library layers vs fp64, including the
scale-direction derivative sum_i <d agg_i, agg_i> (a cancellation residual after BatchNorm)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G

torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(0)
H, N = 64, 3200
h0 = torch.randn(H, generator=g)
c = torch.randint(2, 7, (N,), generator=g).float()
X = (c[:, None] * h0[None, :])
wt = torch.randn(N, H, generator=g)
own = G.GIN.MLP(H, H).cuda()
ref = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.BatchNorm1d(H), torch.nn.ReLU(inplace=True), torch.nn.Linear(H, H))
ref.load_state_dict(own.state_dict())
ref64 = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.BatchNorm1d(H), torch.nn.ReLU(inplace=True), torch.nn.Linear(H, H)).double()
ref64.load_state_dict(own.state_dict())
refg = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.BatchNorm1d(H), torch.nn.ReLU(inplace=True), torch.nn.Linear(H, H)).cuda()
refg.load_state_dict(own.state_dict())
res = {}
for name, m, x, w in [('fp64', ref64, X.double(), wt.double()), ('cpu fp32', ref, X.clone(), wt), ('torch cuda fp32', refg, X.cuda(), wt.cuda()),
                      ('own strict', own, X.cuda(), wt.cuda())]:
    m.train()
    x = x.clone().requires_grad_(True)
    out = torch.relu(m(x))
    (out * w).sum().backward()
    res[name] = dict(out=out.detach().double().cpu(), dx=x.grad.double().cpu(), S=float((x.grad.double().cpu() * x.detach().double().cpu()).sum()),
                     **{n: p.grad.double().cpu() for n, p in m.named_parameters()})
t = res['fp64']
print(f'{"":18s} ' + ' '.join(f'{k:>12s}' for k in t if k != 'S') + f' {"S=sum<dx,x>":>14s}')
for name, r in res.items():
    row = []
    for k in t:
        if k == 'S':
            continue
        row.append(float((r[k] - t[k]).abs().max() / t[k].abs().max().clamp_min(1e-30)))
    print(f'{name:18s} ' + ' '.join(f'{v:12.2e}' for v in row) + f' {r["S"]:14.6e}')
