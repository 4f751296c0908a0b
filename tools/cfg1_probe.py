"""Development probe (GPU): the ill-conditioned constant-feature step cfg1_L2 of tests/test_gpu_parity.py -- error of
every parameter gradient against the fp64 oracle, beside the fp32 oracle's own error, with the dense layers swapped
between this library's kernels and torch's library ops, to locate where the hardware path loses accuracy."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import dense
import dp_gsat_b200.nn as NN
import tests.test_gpu_parity as P
from dp_gsat_b200.data import ba2motifs_batch


def run(tag):
    b = ba2motifs_batch(128, seed=0)
    go, gg = P._build_pair(G, b, 64, 2, True, 0.3, 0.5, 'att')
    go64 = copy.deepcopy(go).double()
    for m in (go, gg, go64):
        m.train(True)
    u = torch.rand(b.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10)
    caps = {}

    def cap(name, store):
        def hook(mod, inp, out):
            store.setdefault(name, []).append(out.detach().double().cpu())
        return hook
    for tagm, m in (('o', go), ('t', go64), ('g', gg)):
        st = caps.setdefault(tagm, {})
        m.clf.convs[0].nn[0].register_forward_hook(cap('L0.lin1', st))
        m.clf.convs[0].nn[1].register_forward_hook(cap('L0.bn', st))
        m.clf.convs[0].nn[3].register_forward_hook(cap('L0.lin2', st))
        m.clf.convs[1].nn[1].register_forward_hook(cap('L1.bn', st))
        m.clf.convs[1].nn[3].register_forward_hook(cap('L1.lin2', st))
    ea_o, lo, _, _ = go.forward_pass(b, 12, True, noise_u=u)
    b64 = b.to('cpu')
    b64.x = b64.x.double()
    ea_t, lt, _, _ = go64.forward_pass(b64, 12, True, noise_u=u.double())
    ea_g, lg, _, _ = gg.forward_pass(b.to('cuda'), 12, True, noise_u=u.cuda())
    relf = lambda a, t: float((a.double().cpu() - t.double()).abs().max() / t.double().abs().max())
    print(f'  {tag}: forward max-abs-err / max vs fp64: edge_att gpu {relf(ea_g, ea_t):.2e} cpu32 {relf(ea_o, ea_t):.2e}')
    for name in caps['t']:
        for i in range(len(caps['t'][name])):
            print(f'      {name}[pass {i}]  gpu {relf(caps["g"][name][i], caps["t"][name][i]):.2e}  cpu32 {relf(caps["o"][name][i], caps["t"][name][i]):.2e}')
    lo.backward(), lt.backward(), lg.backward()
    named = lambda m: dict(list(m.clf.named_parameters()) + [('ext.' + k, v) for k, v in m.extractor.named_parameters()])
    po, pt, pg = named(go), named(go64), named(gg)
    worst = 0.0
    for k in po:
        if po[k].grad is None:
            continue
        t = pt[k].grad.double()
        if float(t.abs().max()) < 1e-10:
            continue
        eg = float((pg[k].grad.double().cpu() - t).abs().max())
        eo = float((po[k].grad.double() - t).abs().max())
        worst = max(worst, eg / max(eo, 1e-30))
        print(f'  {tag:24s} {k:34s} max {float(t.abs().max()):.2e} err_gpu {eg:.2e} err_cpu32 {eo:.2e} ratio {eg / max(eo, 1e-30):6.1f}')
    print(f'{tag}: worst ratio {worst:.1f}')


torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
run('own linear + own BN') if len(sys.argv) > 1 else None
orig_lin, orig_bn, orig_small = dense.Linear.forward, NN.BatchNorm1d.forward, NN.ops.small_linear
dense.Linear.forward = lambda self, x: torch.nn.functional.linear(x, self.weight, self.bias)
NN.ops.small_linear = lambda x, w, b: torch.nn.functional.linear(x, w, b)
run('torch linear + own BN') if len(sys.argv) > 1 else None
NN.BatchNorm1d.forward = lambda self, x: torch.nn.BatchNorm1d.forward(self, x)
run('torch linear + torch BN') if len(sys.argv) > 1 else None
dense.Linear.forward, NN.ops.small_linear = orig_lin, orig_small
run('own linear + torch BN') if len(sys.argv) > 1 else None

# ---- incoherent noise of a product on DEGENERATE rows (every row a small integer multiple of one vector, as the
# constant-feature batches produce): spread over rows of z_i / c_i, relative to |z / c|, per implementation
print('\nincoherent row-to-row noise on rows c_i * h0 (c_i in 2..6):')
g = torch.Generator().manual_seed(0)
h0 = torch.randn(64, generator=g)
c = torch.randint(2, 7, (4096,), generator=g).float()
X = c[:, None] * h0[None, :]
W = torch.randn(64, 64, generator=g) / 8
z64 = (X.double() @ W.double().t()) / c.double()[:, None]
for name, z in [('cpu fp32', X @ W.t()), ('cuBLAS fp32', (X.cuda() @ W.cuda().t()).cpu()),
                ('strict tcgen05', dense.linear_forward(X.cuda(), W.cuda(), None, True).cpu())]:
    zr = z.double() / c.double()[:, None]
    dev = (zr - z64)
    coherent = dev.mean(0)
    incoh = (dev - coherent).abs().max(0).values
    scale = z64.abs().mean(0)
    print(f'  {name:16s} coherent err / |z| (median over channels) {float((coherent.abs() / scale).median()):.2e}   '
          f'incoherent max err / |z| (median) {float((incoh / scale).median()):.2e}  (max over channels) {float((incoh / scale).max()):.2e}')
