"""One fused-extractor forward + backward on a mid-sized BA-2Motifs batch: the command line profiled under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch
from tools.ext_fused_probe import fwd_full, bwd, dev

n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
b = ba2motifs_batch(n_graphs, seed=0).to(dev)
gi = G.get_graph_index(b.edge_index, b.batch)
emb = torch.randn(b.num_nodes, H, device=dev)
w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
w2 = torch.randn(H, 4 * H, device=dev) / 22
w3 = torch.randn(H, device=dev) / 11
b3 = torch.zeros(1, device=dev)
dlogit = torch.randn(gi.E, device=dev)
logit, xh2t, rstd2, seeds = fwd_full(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, 1)
for _ in range(3):
    bwd(emb, gi, w1, w2, w3, dlogit, xh2t, rstd2, seeds, True, None, None, 0.5, 1)
torch.cuda.synchronize()
print('done', gi.E)
