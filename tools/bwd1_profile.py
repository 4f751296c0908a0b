"""Small driver for ncu: extractor forward then ext_bwd1 / linear_bf16in at a moderate size."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import tc
from dp_gsat_b200._lib import lib, ptr, stream
from dp_gsat_b200.data import ba2motifs_batch

dev = 'cuda'
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
H = 128
b = ba2motifs_batch(ng, seed=0).to(dev)
gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
torch.manual_seed(0)
emb = torch.relu(torch.randn(gi.N, H, device=dev))
w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
w2 = torch.randn(H, 4 * H, device=dev) / 22
w3 = torch.randn(1, H, device=dev) / 11
b3 = torch.zeros(1, device=dev)
L = lib()
tile_row, tile_seg, T = gi.tile_plan('edge')
C1 = 4 * H
logit, sv = tc.extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=True, pdrop=0.5, training=True, seed=1)
dz2 = (torch.randn(gi.E, H, device=dev) / 10).bfloat16()
dz1 = torch.empty(gi.E, C1, dtype=torch.bfloat16, device=dev)
w2t = tc.prep_weight(w2, transpose=True)
w1t = tc.prep_weight(w1, transpose=True)
df = torch.empty(gi.E, 2 * H, device=dev)
for _ in range(3):
    L.call('gsatb_tc_ext_bwd1', ptr(dz2), ptr(w2t), ptr(sv['xhat1']), ptr(sv['rstd1']), None, ctypes.c_uint64(1), ctypes.c_float(0.5), 1,
           ptr(tile_row), ptr(tile_seg), ptr(gi.edge_ptr), T, ptr(dz1), gi.E, H, C1, stream())
    L.call('gsatb_tc_linear_bf16in_fwd', ptr(dz1), C1, ptr(w1t), ptr(df), 2 * H, gi.E, C1, 2 * H, stream())
torch.cuda.synchronize()
print('ok', float(df.abs().mean()))
