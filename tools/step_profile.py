"""Kernel-time table of one GSAT-GIN training step (torch.profiler / CUPTI), at any batch size.
usage: python tools/step_profile.py [graphs] [hidden] [precision]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch
from dp_gsat_b200.parallel import TrainStep

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 196000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
prec = sys.argv[3] if len(sys.argv) > 3 else 'bf16'
dev = 'cuda'
cfg = {'model_name': 'GIN', 'hidden_size': H, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
shared = {'learn_edge_att': True, 'extractor_dropout_p': 0.5}
torch.manual_seed(0)
data = ba2motifs_batch(ng, seed=0).to(dev)
clf = G.get_model(10, 0, 2, False, cfg, dev)
ext = G.ExtractorMLP(H, shared).to(dev)
clf.precision = ext.precision = prec
gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5, lazy_metrics=True)
gsat.train()
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
print(f'# graphs={ng} H={H} precision={prec}: {tot / 1e3:.2f} ms of kernel time per step (E={data.num_edges})')
for e in sorted(ev, key=lambda e: -e.device_time_total)[:40]:
    print(f'{e.device_time_total / 2e3:9.3f} ms {100 * e.device_time_total / 2 / tot:5.1f}% x{e.count // 2:4d}  {e.key[:120]}')
