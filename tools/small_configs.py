"""BASELINE configs 1-3 at their own (small) sizes: step time of this repo (eager launches and one-CUDA-graph replay)
next to the CPU oracle on the host cores.  At E ~ 6.5 k .. 14 k a step is launch-latency bound.
usage: python tools/small_configs.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import dp_gsat_b200 as G
from oracle import gsat_oracle as O
from dp_gsat_b200.data import (ba2motifs_batch, molhiv_like_batch, in_degree_histogram, load_mutag_fixture,
                               line_graph_dual, graph_contiguous_relabel, batch_from_edge_list)
from dp_gsat_b200.parallel import TrainStep

dev = 'cuda'


def mutag_dual(n_graphs=128):
    src, dst, ng = load_mutag_fixture(os.path.join(ROOT, 'tests', 'golden', 'mutag_slice.npz'))
    keep = ng[src] < n_graphs
    ds, dd, dng = line_graph_dual(src[keep], dst[keep], ng)
    ds, dd = graph_contiguous_relabel(ds, dd, dng)
    return batch_from_edge_list(ds, dd, dng, x_dim=31, seed=0)


def build(mod, batch, cfg, shared, hidden, device=None):
    torch.manual_seed(0)
    ea_dim = 0 if batch.edge_attr is None else batch.edge_attr.shape[1]
    x_dim = batch.x.shape[1]
    if mod is O:
        clf, ext = O.get_model(x_dim, ea_dim, 2, False, cfg), O.ExtractorMLP(hidden, shared)
    else:
        clf, ext = G.get_model(x_dim, ea_dim, 2, False, cfg, device), G.ExtractorMLP(hidden, shared).to(device)
    g = mod.GSAT(clf, ext, mod.Criterion(2, False), learn_edge_att=shared['learn_edge_att'], final_r=0.5,
                 **({'lazy_metrics': True} if mod is G else {}))
    g.train()
    return g


def time_gpu(gsat, data, graph):
    step = TrainStep(gsat, lr=1e-3)
    for _ in range(5):
        step(data, 0)
    if graph and not step.enable_cuda_graph(data, 0, warmup=2):
        return float('nan')
    for _ in range(5):
        step(data, 0)
    torch.cuda.synchronize()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(data, 0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def time_cpu(gsat, batch, threads):
    torch.set_num_threads(threads)
    opt = torch.optim.Adam(list(gsat.extractor.parameters()) + list(gsat.clf.parameters()), lr=1e-3)
    ts = []
    for i in range(7):
        t0 = time.perf_counter()
        _, loss, _, _ = gsat.forward_pass(batch, 0, True)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if i >= 2:
            ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


gin = lambda L: {'model_name': 'GIN', 'hidden_size': 64, 'n_layers': L, 'dropout_p': 0.3, 'use_edge_attr': False}
cases = [
    ('cfg1 BA-2Motifs B=128 GIN H=64 L=2', ba2motifs_batch(128, seed=0), gin(2), {'learn_edge_att': True, 'extractor_dropout_p': 0.5}, 64, 'bf16'),
    ('cfg1 BA-2Motifs B=128 GIN H=64 L=3', ba2motifs_batch(128, seed=0), gin(3), {'learn_edge_att': True, 'extractor_dropout_p': 0.5}, 64, 'bf16'),
    ('cfg2 mutag-dual B=128 GIN H=64 L=2 (reverse-edge average)', mutag_dual(128), gin(2), {'learn_edge_att': True, 'extractor_dropout_p': 0.5}, 64, 'bf16'),
]
b3 = molhiv_like_batch(256, seed=0, with_edge_attr=False)
pna = {'model_name': 'PNA', 'hidden_size': 80, 'n_layers': 4, 'dropout_p': 0.3, 'atom_encoder': True, 'use_edge_attr': False,
       'aggregators': ['mean', 'min', 'max', 'std'], 'scalers': False, 'deg': in_degree_histogram(b3)}
cases.append(('cfg3 molhiv-shaped B=256 PNA H=80 L=4 (lift path)', b3, pna, {'learn_edge_att': False, 'extractor_dropout_p': 0.5}, 80, 'fp32'))

ncpu = os.cpu_count() or 1
print(f'# host cores {ncpu}; CPU = oracle/gsat_oracle.py (pure PyTorch restatement of the reference path), median of 5 steps')
print('# config | N | E | GPU eager ms | GPU graph ms | M edges/s (graph) | CPU 5 threads ms | CPU all cores ms | speed-up vs CPU best')
for name, b, cfg, shared, hidden, prec in cases:
    data = b.to(dev)
    gg = build(G, b, cfg, shared, hidden, dev)
    gg.clf.precision = gg.extractor.precision = prec
    t_eager = time_gpu(gg, data, False)
    gg2 = build(G, b, cfg, shared, hidden, dev)
    gg2.clf.precision = gg2.extractor.precision = prec
    t_graph = time_gpu(gg2, data, True)
    go = build(O, b, cfg, shared, hidden)
    c5 = time_cpu(go, b, 5)              # the reference's own setting, src/run_gsat.py:1049
    call = time_cpu(go, b, ncpu)
    best = min(t_eager, t_graph) if t_graph == t_graph else t_eager
    print(f'{name} | {b.num_nodes} | {b.num_edges} | {t_eager:.3f} | {t_graph:.3f} | {b.num_edges / best / 1e3:.2f} | '
          f'{c5:.1f} | {call:.1f} | {min(c5, call) / best:.0f}x', flush=True)
