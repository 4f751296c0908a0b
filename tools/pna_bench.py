"""GSAT-PNA (BASELINE config 3 shape, scaled up) step time and K4 kernel bandwidth, for the record.
usage: python tools/pna_bench.py [graphs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import molhiv_like_batch, in_degree_histogram
from dp_gsat_b200.parallel import TrainStep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = 'cuda'
b = molhiv_like_batch(ng, seed=0, with_edge_attr=False)
cfg = {'model_name': 'PNA', 'hidden_size': 80, 'n_layers': 4, 'dropout_p': 0.3, 'atom_encoder': True,
       'use_edge_attr': False, 'aggregators': ['mean', 'min', 'max', 'std'], 'scalers': False,
       'deg': in_degree_histogram(b)}
shared = {'learn_edge_att': False, 'extractor_dropout_p': 0.5}
torch.manual_seed(0)
clf = G.get_model(9, 0, 2, False, cfg, dev)
ext = G.ExtractorMLP(80, shared).to(dev)
gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=False, final_r=0.7, lazy_metrics=True)
gsat.train()
data = b.to(dev)
step = TrainStep(gsat, lr=1e-3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for precision in ('bf16', 'fp32'):
    clf.precision = ext.precision = precision
    for _ in range(3):
        step(data, 0)
    torch.cuda.synchronize()
    e0.record()
    n = 5
    for _ in range(n):
        step(data, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    what = {'bf16': "bf16 mode: post_nn as 'bf16x2' tcgen05 GEMM, fc_out / extractor on tcgen05",
            'fp32': 'fp32 strict mode: split-bf16 x3 tcgen05 GEMMs'}[precision]
    print(f'GSAT-PNA (molhiv-shaped, {what}): graphs={ng} N={data.num_nodes} E={data.num_edges} H=80 L=4: '
          f'{ms:.2f} ms/step = {data.num_edges / ms / 1e3:.2f} M edges/s', flush=True)
gi = G.get_graph_index(data.edge_index, data.batch, data.num_graphs)
N, E, H = gi.N, gi.E, 80
x = torch.randn(N, H, device=dev)
att = torch.rand(E, 1, device=dev)
aggs = ['mean', 'min', 'max', 'std']
for _ in range(3):
    G.ops.pna_aggregate(x, None, att, gi, aggs)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    G.ops.pna_aggregate(x, None, att, gi, aggs)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
F_ = 2 * H
nbytes = 4.0 * N * H + 8.0 * E + 4.0 * N + 4.0 * N * len(aggs) * F_ + 4.0 * N * F_ * 4     # out + saved mean/msq/argmin/argmax
print(f'K4 pna_aggregate fwd: {t:.3f} ms, algorithmic {nbytes / 1e9:.2f} GB -> {nbytes / t / 1e6:.0f} GB/s = {nbytes / t / 1e6 / peak:.2f} of measured peak')
