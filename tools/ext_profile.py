"""Small driver for ncu: a few fused extractor forwards (and node linears) at a moderate size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import tc
from dp_gsat_b200.data import ba2motifs_batch

dev = 'cuda'
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
b = ba2motifs_batch(ng, seed=0).to(dev)
gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
H = 128
torch.manual_seed(0)
emb = torch.relu(torch.randn(gi.N, H, device=dev))
w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
w2 = torch.randn(H, 4 * H, device=dev) / 22
w3 = torch.randn(1, H, device=dev) / 11
b3 = torch.zeros(1, device=dev)
for _ in range(3):
    out = tc.extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=True, pdrop=0.5, training=True, seed=1)
torch.cuda.synchronize()
print('ok', float(out[0].abs().mean()))

