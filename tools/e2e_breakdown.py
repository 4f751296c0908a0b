"""Where the end-to-end step (bench.py `e2e`: pinned host batch -> H2D -> K0 index build -> step -> loss.item()) spends
its time beyond the device-resident step: the same loop as bench.py with a synchronise + wall clock after each phase
(which removes the overlap, so the phases add up to MORE than the pipelined e2e figure -- it attributes, not measures)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch
from dp_gsat_b200.parallel import TrainStep

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 196000
dev = torch.device('cuda', 0)
cfg = {'model_name': 'GIN', 'hidden_size': 128, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
host = ba2motifs_batch(graphs, seed=0).pin_memory()
torch.manual_seed(0)
clf = G.get_model(host.x.shape[1], 0, 2, False, cfg, dev)
ext = G.ExtractorMLP(128, {'learn_edge_att': True, 'extractor_dropout_p': 0.5}).to(dev)
clf.precision = ext.precision = 'bf16'
gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5, lazy_metrics=True)
gsat.train()
step = TrainStep(gsat, lr=1e-3)
sync = torch.cuda.synchronize


def clock(fn):
    sync()
    t0 = time.perf_counter()
    out = fn()
    sync()
    return out, (time.perf_counter() - t0) * 1e3


for it in range(4):
    d, t_h2d = clock(lambda: host.to(dev, non_blocking=True))
    gi, t_k0 = clock(lambda: G.get_graph_index(d.edge_index, d.batch, d.num_graphs))
    _, t_plan = clock(lambda: gi.ext_plan('edge', 112))
    out, t_fwd = clock(lambda: gsat.forward_pass(d, 0, True))
    loss = out[1]
    _, t_bwd = clock(lambda: (step.bucket.zero(), loss.backward()))
    _, t_opt = clock(lambda: step.optimizer.step())
    _, t_item = clock(lambda: float(loss.item()))
    _, t_clear = clock(lambda: G.clear_index_cache())
    del d, gi, out, loss
    print(f'iter {it}: H2D {t_h2d:.2f}  K0 {t_k0:.2f}  tile plan {t_plan:.2f}  forward {t_fwd:.2f}  backward {t_bwd:.2f}  adam {t_opt:.2f}  '
          f'item {t_item:.2f}  clear {t_clear:.2f}  sum {t_h2d + t_k0 + t_plan + t_fwd + t_bwd + t_opt + t_item + t_clear:.2f} ms', flush=True)
print('memory: allocated %.1f GB, reserved %.1f GB' % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
