"""K3 (GIN aggregation) micro-benchmark: algorithmic GB/s of forward / backward against the measured HBM peak.
usage: python tools/k3_bench.py [graphs] [hidden]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 196000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = 'cuda'
b = ba2motifs_batch(ng, seed=0).to(dev)
gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
N, E = gi.N, gi.E
torch.manual_seed(0)
x = torch.randn(N, H, device=dev, requires_grad=True)
att = torch.rand(E, 1, device=dev, requires_grad=True)
gout = torch.randn(N, H, device=dev)
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for use_att in (True, False):
    a = att if use_att else None
    out = G.ops.gin_aggregate(x, a, gi, 0.0)
    ref = torch.zeros_like(x).index_add_(0, b.edge_index[1], x.detach()[b.edge_index[0]] * (att.detach() if use_att else 1.0)) + x.detach()
    err = float((out.detach() - ref).abs().max())
    t_f = timeit(lambda: G.ops.gin_aggregate(x.detach(), None if a is None else a.detach(), gi, 0.0))
    out = G.ops.gin_aggregate(x, a, gi, 0.0)
    t_b = timeit(lambda: torch.autograd.grad(out, [x] + ([att] if use_att else []), gout, retain_graph=True))
    bf, bb = 8.0 * N * H + 8.0 * E + 4.0 * N, 12.0 * N * H + 16.0 * E
    print(f'N={N} E={E} H={H} att={use_att}: max err vs torch {err:.2e}; fwd {t_f:.3f} ms = {bf / t_f / 1e6:.0f} GB/s '
          f'({bf / t_f / 1e6 / peak:.3f} of measured {peak:.0f}); bwd {t_b:.3f} ms = {bb / t_b / 1e6:.0f} GB/s ({bb / t_b / 1e6 / peak:.3f})')
