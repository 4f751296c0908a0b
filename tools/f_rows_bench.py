"""Throughput of the SURVEY section 8f rows on one B200: GINE / LE aggregation, the fused ogb
encoders, device collate (GB/s of algorithmic bytes against the measured HBM peak) and the fork's two-model training step
`DualGSAT.dual_train_one_batch` (src/run_gsat.py:620-637) on a BA-2Motifs primal batch + its line-graph dual.
usage: python tools/f_rows_bench.py [graphs]      (writes a table to stdout; committed as profiles/r2_f_rows.txt)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import ops
from dp_gsat_b200.data import ba2motifs_batch, molhiv_like_batch, line_graph_dual, graph_contiguous_relabel

try:
    PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    PEAK = 6650.0
dev = 'cuda'
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 100000


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def row(name, ms, nbytes):
    print(f'{name:46s} {ms:8.3f} ms  {nbytes / 1e9:7.2f} GB  {nbytes / ms / 1e6:7.0f} GB/s  {nbytes / ms / 1e6 / PEAK:5.2f} of measured HBM peak', flush=True)


b = ba2motifs_batch(ng, seed=0).to(dev)
gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
N, E = gi.N, gi.E
print(f'# BA-2Motifs-shaped batch: graphs={ng} N={N} E={E}; HBM peak {PEAK:.0f} GB/s (MEASURED_PEAKS.json)')
for H in (64, 128):
    x = torch.randn(N, H, device=dev, requires_grad=True)
    ef = torch.randn(E, H, device=dev, requires_grad=True)
    att = torch.rand(E, 1, device=dev, requires_grad=True)
    w = torch.rand(E, device=dev)
    g = torch.randn(N, H, device=dev)
    row(f'gine_aggregate fwd H={H}', timeit(lambda: ops.gine_aggregate(x.detach(), ef.detach(), att.detach(), gi, 0.0)),
        8.0 * N * H + 4.0 * E * H + 12.0 * E)
    out = ops.gine_aggregate(x, ef, att, gi, 0.0)
    row(f'gine_aggregate bwd H={H}', timeit(lambda: torch.autograd.grad(out, (x, ef, att), g, retain_graph=True)),
        12.0 * N * H + 8.0 * E * H + 16.0 * E)
    a_, b_ = torch.randn(N, H, device=dev, requires_grad=True), torch.randn(N, H, device=dev, requires_grad=True)
    row(f'le_aggregate fwd H={H}', timeit(lambda: ops.le_aggregate(a_.detach(), b_.detach(), w, att.detach(), gi)),
        12.0 * N * H + 12.0 * E)
    out = ops.le_aggregate(a_, b_, w, att, gi)
    row(f'le_aggregate bwd H={H}', timeit(lambda: torch.autograd.grad(out, (a_, b_, att), g, retain_graph=True)),
        20.0 * N * H + 24.0 * E)
    del x, ef, att, out, a_, b_

# fused ogb encoders (AtomEncoder: 9 tables) on a molhiv-shaped batch
mb = molhiv_like_batch(ng, seed=0, with_edge_attr=True).to(dev)
for H in (80, 128):
    enc = G.AtomEncoder(H).to(dev)
    M, K = mb.x.shape
    row(f'embedding_sum fwd (AtomEncoder) H={H} M={M}', timeit(lambda: enc(mb.x).detach() if False else ops.embedding_sum(mb.x, [t.detach() for t in enc._tables()])),
        8.0 * M * K + 4.0 * M * H)
    out = enc(mb.x)
    gm = torch.randn(M, H, device=dev)
    row(f'embedding_sum bwd (AtomEncoder) H={H}', timeit(lambda: torch.autograd.grad(out, list(enc.parameters()), gm, retain_graph=True)),
        4.0 * M * H + 8.0 * M * K)
    del out

# device collate: the packed dataset lives in HBM, a batch = a gather of whole graphs
from dp_gsat_b200.loader import PackedDataset
host = ba2motifs_batch(min(ng, 50000), seed=1)
nptr = np.concatenate([[0], np.cumsum(np.bincount(host.batch.numpy(), minlength=host.num_graphs))])
eb = host.batch[host.edge_index[0]].numpy()
eptr = np.concatenate([[0], np.cumsum(np.bincount(eb, minlength=host.num_graphs))])
local = host.edge_index - torch.from_numpy(nptr[:-1])[host.batch[host.edge_index[0]]].unsqueeze(0)
ds = PackedDataset({'x': host.x, 'edge_index': local.contiguous(), 'edge_attr': None, 'edge_label': host.edge_label,
                    'node_label': None, 'y': host.y}, nptr, eptr, dev)
ids = np.random.default_rng(0).permutation(host.num_graphs)
ms = timeit(lambda: ds.collate(ids), n=5, warm=2)
moved = 2.0 * (host.x.numel() * 4 + host.edge_index.numel() * 8 + host.edge_label.numel() * host.edge_label.element_size() + host.y.numel() * host.y.element_size()) + host.batch.numel() * 8
row(f'collate (PackedDataset, {host.num_graphs} graphs, shuffled ids)', ms, moved)

# the fork's two-model training step
p = ba2motifs_batch(min(ng, 40000), seed=2)
p.x = torch.rand(p.num_nodes, 10, generator=torch.Generator().manual_seed(2))
src, dst = p.edge_index[0].numpy(), p.edge_index[1].numpy()
dsrc, ddst, dng = line_graph_dual(src, dst, p.batch.numpy())
dsrc, ddst = graph_contiguous_relabel(dsrc, ddst, dng)
d = G.Batch(torch.rand(p.num_edges, 7), torch.from_numpy(np.stack([dsrc, ddst])), torch.from_numpy(dng), p.y.clone(), None,
            torch.zeros(dsrc.shape[0]), p.num_graphs)
for H, precision in ((64, 'bf16'), (64, 'fp32')):
    cfg = {'model_name': 'GIN', 'hidden_size': H, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    sc = {'learn_edge_att': False, 'extractor_dropout_p': 0.5, 'precision_k': 5, 'num_viz_samples': 0, 'viz_interval': 10,
          'viz_norm_att': True}
    mc = {'method_name': 'GSAT', 'pred_loss_coef': 1, 'info_loss_coef': 1, 'epochs': 100, 'decay_interval': 10,
          'decay_r': 0.1, 'final_r': 0.5, 'lr': 1e-3}
    torch.manual_seed(0)
    pc, pe = G.get_model(10, 0, 2, False, cfg, dev), G.ExtractorMLP(H, sc, 'primal').to(dev)
    dc, de = G.get_model(7, 0, 2, False, cfg, dev), G.ExtractorMLP(H, sc, 'dual').to(dev)
    for m in (pc, pe, dc, de):
        m.precision = precision
    po = torch.optim.Adam(list(pe.parameters()) + list(pc.parameters()), lr=1e-3)
    do = torch.optim.Adam(list(de.parameters()) + list(dc.parameters()), lr=1e-3)
    gsat = G.DualGSAT.from_reference_args(pc, pe, po, None, None, dev, None, 'mutag', 2, False, 0, mc, sc, cfg,
                                          dc, de, do, None, None, dev, None, 'mutag_dual', 2, False, 0, mc, sc, cfg)
    pd_, dd_ = p.to(dev), d.to(dev)
    ms = timeit(lambda: gsat.dual_train_one_batch(pd_, dd_, 3), n=5, warm=3)
    print(f'DualGSAT.dual_train_one_batch (GIN H={H} L=2, precision {precision}): primal E={p.num_edges}, dual N={d.num_nodes} E={d.num_edges}: '
          f'{ms:.2f} ms/step = {(p.num_edges + d.num_edges) / ms / 1e3:.1f} M (primal + dual) edges/s', flush=True)
