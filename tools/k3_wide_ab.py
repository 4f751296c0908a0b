import os, sys, torch
sys.path.insert(0, '/root/repo')
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch
peak = 6537.6
b = ba2motifs_batch(196000, seed=0).to('cuda')
gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
N, E = gi.N, gi.E
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for H in (64, 128):
    x = torch.randn(N, H, device='cuda', requires_grad=True); att = torch.rand(E, 1, device='cuda', requires_grad=True); g = torch.randn(N, H, device='cuda')
    tf = timeit(lambda: G.ops.gin_aggregate(x.detach(), att.detach(), gi, 0.0))
    out = G.ops.gin_aggregate(x, att, gi, 0.0)
    tb = timeit(lambda: torch.autograd.grad(out, [x, att], g, retain_graph=True))
    tb2 = timeit(lambda: torch.autograd.grad(out, [x], g, retain_graph=True))
    bf, bb = 8.0*N*H + 8.0*E + 4.0*N, 12.0*N*H + 16.0*E
    print(f"WIDE={os.environ.get('GSATB_K3_WIDE')} H={H}: fwd {tf:.3f} ms {bf/tf/1e6/peak:.2f} | bwd(datt) {tb:.3f} ms {bb/tb/1e6/peak:.2f} | bwd(no datt) {tb2:.3f} ms", flush=True)
