"""Kernel-time table of one GSAT-PNA training step (molhiv-shaped, fp32 strict path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import dp_gsat_b200 as G
from dp_gsat_b200.data import molhiv_like_batch, in_degree_histogram
from dp_gsat_b200.parallel import TrainStep
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = 'cuda'
b = molhiv_like_batch(ng, seed=0, with_edge_attr=False)
cfg = {'model_name': 'PNA', 'hidden_size': 80, 'n_layers': 4, 'dropout_p': 0.3, 'atom_encoder': True, 'use_edge_attr': False,
       'aggregators': ['mean', 'min', 'max', 'std'], 'scalers': False, 'deg': in_degree_histogram(b)}
shared = {'learn_edge_att': False, 'extractor_dropout_p': 0.5}
torch.manual_seed(0)
clf = G.get_model(9, 0, 2, False, cfg, dev)
ext = G.ExtractorMLP(80, shared).to(dev)
gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=False, final_r=0.7, lazy_metrics=True)
gsat.train()
clf.precision = ext.precision = (sys.argv[2] if len(sys.argv) > 2 else 'bf16')
data = b.to(dev)
step = TrainStep(gsat, lr=1e-3)
for _ in range(3):
    step(data, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step(data, 0)
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev) / 2
print(f'# PNA graphs={ng} N={data.num_nodes} E={data.num_edges}: {tot / 1e3:.2f} ms of kernel time per step')
for e in sorted(ev, key=lambda e: -e.device_time_total)[:30]:
    print(f'{e.device_time_total / 2e3:9.3f} ms {100 * e.device_time_total / 2 / tot:5.1f}% x{e.count // 2:4d}  {e.key[:110]}')
