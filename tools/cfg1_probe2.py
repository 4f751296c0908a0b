"""Development probe (GPU): cfg1_L2 gradients, this library's dense layers vs torch's, element-wise relation."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import dense
import dp_gsat_b200.nn as NN
import tests.test_gpu_parity as P
from dp_gsat_b200.data import ba2motifs_batch
torch.backends.cuda.matmul.allow_tf32 = False

def run(randx=False):
    b = ba2motifs_batch(128, seed=0)
    if randx:
        b.x = torch.rand(b.x.shape, generator=torch.Generator().manual_seed(5))
    go, gg = P._build_pair(G, b, 64, 2, True, 0.3, 0.5, 'att')
    go64 = copy.deepcopy(go).double()
    for m in (go64, gg):
        m.train(True)
    u = torch.rand(b.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10)
    b64 = b.to('cpu'); b64.x = b64.x.double()
    out = {}
    ea_t, lt, ld_t, _ = go64.forward_pass(b64, 12, True, noise_u=u.double())
    lt.backward()
    named = lambda m: dict(list(m.clf.named_parameters()) + [('ext.' + k, v) for k, v in m.extractor.named_parameters()])
    ea_g, lg, ld_g, _ = gg.forward_pass(b.to('cuda'), 12, True, noise_u=u.cuda())
    lg.backward()
    return {k: v.grad.double().cpu() for k, v in named(gg).items() if v.grad is not None}, {k: v.grad.double() for k, v in named(go64).items() if v.grad is not None}, (ld_t, ld_g)

for randx in (False, True):
    own, t64, lds = run(randx)
    ol, ob, osm = dense.Linear.forward, NN.BatchNorm1d.forward, NN.ops.small_linear
    dense.Linear.forward = lambda self, x: torch.nn.functional.linear(x, self.weight, self.bias)
    NN.ops.small_linear = lambda x, w, b: torch.nn.functional.linear(x, w, b)
    NN.BatchNorm1d.forward = lambda self, x: torch.nn.BatchNorm1d.forward(self, x)
    tor, _, _ = run(randx)
    dense.Linear.forward, NN.BatchNorm1d.forward, NN.ops.small_linear = ol, ob, osm
    print('randx', randx, 'loss dicts', lds)
    for k in ['ext.feature_extractor.8.bias', 'ext.feature_extractor.8.weight', 'node_encoder.bias', 'convs.1.nn.0.weight', 'convs.0.nn.0.weight', 'fc_out.0.weight']:
        o, t, r = own[k].flatten(), tor[k].flatten(), t64[k].flatten()
        i = int(r.abs().argmax())
        print(f'  {k:34s} fp64 {float(r[i]): .8e}  own {float(o[i]): .8e}  torch {float(t[i]): .8e}   own-fp64 {float(o[i]-r[i]): .2e}  torch-fp64 {float(t[i]-r[i]): .2e}  '
              f'proj(own-fp64 on fp64)/|fp64|^2 {float(torch.dot(o - r, r) / torch.dot(r, r)): .2e}  rel resid {float((o - r - torch.dot(o - r, r) / torch.dot(r, r) * r).norm() / r.norm()):.2e}')
