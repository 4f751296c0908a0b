"""One fused-extractor forward (and backward when available) on a mid-sized BA-2Motifs batch: the command line profiled
under ncu (tools/ext_fused_probe.py has the correctness checks; this script only launches the kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch
from tools.ext_fused_probe import fwd, dev

n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
b = ba2motifs_batch(n_graphs, seed=0).to(dev)
gi = G.get_graph_index(b.edge_index, b.batch)
emb = torch.randn(b.num_nodes, H, device=dev)
w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
w2 = torch.randn(H, 4 * H, device=dev) / 22
w3 = torch.randn(H, device=dev) / 11
b3 = torch.zeros(1, device=dev)
for _ in range(3):
    fwd(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, 1)
torch.cuda.synchronize()
print('done', gi.E)
