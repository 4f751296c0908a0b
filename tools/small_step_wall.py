"""Wall-clock vs kernel time of one training step at a per-rank size of the 8-GPU run (host-overhead check)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch
from dp_gsat_b200.parallel import TrainStep

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 24500
H = 128
dev = 'cuda'
cfg = {'model_name': 'GIN', 'hidden_size': H, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
shared = {'learn_edge_att': True, 'extractor_dropout_p': 0.5}
torch.manual_seed(0)
data = ba2motifs_batch(ng, seed=0).to(dev)
clf = G.get_model(10, 0, 2, False, cfg, dev)
ext = G.ExtractorMLP(H, shared).to(dev)
clf.precision = ext.precision = 'bf16'
gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5, lazy_metrics=True)
gsat.train()
step = TrainStep(gsat, lr=1e-3)
for _ in range(5):
    step(data, 0)
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    step(data, 0)
e1.record()
t_host = time.perf_counter() - t0          # host time to ENQUEUE n steps
torch.cuda.synchronize()
t_wall = time.perf_counter() - t0
print(f'graphs={ng} E={data.num_edges}: device {e0.elapsed_time(e1) / n:.3f} ms/step, host enqueue {t_host / n * 1e3:.3f} ms/step, wall {t_wall / n * 1e3:.3f} ms/step')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        step(data, 0)
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev) / 4
cnt = sum(e.count for e in ev) / 4
print(f'kernel time {tot / 1e3:.3f} ms/step over {cnt:.0f} kernels/step')
