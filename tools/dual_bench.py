"""GPU line-graph (dual) builder throughput on the cfg4 primal batch, with the CPU (numpy, vectorised) builder beside it.
usage: python tools/dual_bench.py [graphs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch, line_graph_dual as np_dual

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 196000
b = ba2motifs_batch(ng, seed=0)
d = b.to('cuda')
G.get_graph_index(d.edge_index, d.batch, d.num_graphs)       # K0 of the primal batch (cached, as in training)
for _ in range(2):
    dei, db = G.line_graph_dual(d.edge_index, d.batch, d.num_graphs)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    dei, db = G.line_graph_dual(d.edge_index, d.batch, d.num_graphs)
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / n
E, Ed = d.num_edges, dei.shape[1]
nbytes = 16.0 * Ed + 8.0 * E + 12.0 * E
print(f'primal E={E}, dual nodes={E}, dual edges={Ed}: GPU {t_gpu * 1e3:.2f} ms per build ({Ed / t_gpu / 1e6:.0f} M dual edges/s, '
      f'{nbytes / t_gpu / 1e9:.0f} GB/s of output + index traffic, includes the size read-back sync)')
sub = ba2motifs_batch(min(ng, 20000), seed=0)
src, dst, ngraph = sub.edge_index[0].numpy(), sub.edge_index[1].numpy(), sub.batch.numpy()
t0 = time.perf_counter()
ds, dd, dng = np_dual(src, dst, ngraph)
t_cpu = time.perf_counter() - t0
print(f'CPU numpy builder on {sub.num_edges} primal edges: {t_cpu * 1e3:.0f} ms ({ds.shape[0] / t_cpu / 1e6:.2f} M dual edges/s)')
