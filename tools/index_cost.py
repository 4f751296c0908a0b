"""Cost of the per-batch index work at cfg4 (K0 build, flag read-back, tile plan), i.e. what e2e pays per fresh batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 196000
b = ba2motifs_batch(ng, seed=0).pin_memory()
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = b.to('cuda', non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    gi = G.get_graph_index(d.edge_index, d.batch, d.num_graphs)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    sym = gi.symmetric
    t3 = time.perf_counter()
    plan = gi.tile_plan('edge')
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f'iter {it}: H2D {1e3 * (t1 - t0):.2f} ms ({b.nbytes() / (t1 - t0) / 1e9:.1f} GB/s), K0 build {1e3 * (t2 - t1):.2f} ms, '
          f'flags read {1e3 * (t3 - t2):.2f} ms, tile plan {1e3 * (t4 - t3):.2f} ms (T={plan[2]})')
    G.clear_index_cache()
