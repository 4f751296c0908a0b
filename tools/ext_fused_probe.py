"""GPU probe for the fused extractor kernels (gsatb_ext_fused_fwd / _bwd): correctness against the same-rounding torch
restatement (tests/helpers/ext_ref.py) on small ragged batches, then timing at the cfg4 shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200._lib import lib, ptr, stream
from dp_gsat_b200 import tc
from tests.helpers.ext_ref import extractor_forward

L = lib()
dev = 'cuda'


def fwd(emb, gi, w1, w2, w3, b3, edge, m1=None, m2=None, pdrop=0.0, training=0, seed=3, want_x=True):
    plan = gi.ext_plan('edge' if edge else 'node')
    rows, T = plan['rows'], plan['T']
    H, C1 = emb.shape[1], w1.shape[0]
    logit = torch.empty(rows, device=dev)
    xh2t = torch.empty(H, T * 128, dtype=torch.bfloat16, device=dev) if want_x else None
    seeds = torch.zeros(2, dtype=torch.int32, device=dev)
    w1p, w2p = tc.prep_weight(w1), tc.prep_weight(w2)
    L.call('gsatb_ext_fused_fwd', ptr(emb), ptr(gi.src) if edge else None, ptr(gi.dst) if edge else None,
           ptr(plan['seg_ptr']), ptr(plan['tile_seg']), ptr(plan['out2']), max(gi.G, 1), ptr(w1p), ptr(w2p), ptr(w3),
           ptr(b3), ptr(m1), ptr(m2), ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training), ptr(logit), ptr(xh2t),
           T * 128, ptr(seeds), rows, H, C1, ctypes.c_float(1e-5), stream())
    return logit, xh2t


def check(H, n_graphs, edge=True, masks=False, gen='molhiv'):
    from dp_gsat_b200.data import molhiv_like_batch, ba2motifs_batch
    b = (molhiv_like_batch(n_graphs, seed=H) if gen == 'molhiv' else ba2motifs_batch(n_graphs, seed=H)).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    g = torch.Generator().manual_seed(H)
    N = b.num_nodes
    emb = (torch.randn(N, H, generator=g) + 0.5).to(dev)
    Kin, C1 = (2 * H, 4 * H) if edge else (H, 2 * H)
    w1 = (torch.randn(C1, Kin, generator=g) / Kin ** 0.5).to(dev)
    w2 = (torch.randn(H, C1, generator=g) / C1 ** 0.5).to(dev)
    w3 = (torch.randn(H, generator=g) / H ** 0.5).to(dev)
    b3 = torch.randn(1, generator=g).to(dev)
    rows = gi.E if edge else gi.N
    pd = 0.5
    m1 = (torch.rand(rows, C1, generator=g) > pd).to(torch.uint8).to(dev) if masks else None
    m2 = (torch.rand(rows, H, generator=g) > pd).to(torch.uint8).to(dev) if masks else None
    logit, _ = fwd(emb, gi, w1, w2, w3, b3, edge, m1, m2, pd if masks else 0.0, 1 if masks else 0)
    torch.cuda.synchronize()
    seg = (gi.edge_ptr if edge else gi.node_ptr).long()
    ref = extractor_forward(emb, gi.src.long() if edge else None, gi.dst.long() if edge else None, seg, w1, w2, w3, b3,
                            m1, m2, pd if masks else 0.0, rounding='bf16').view(-1)
    ref64 = extractor_forward(emb.double(), gi.src.long() if edge else None, gi.dst.long() if edge else None, seg,
                              w1.double(), w2.double(), w3.double(), b3.double(), m1, m2, pd if masks else 0.0).view(-1)
    e = ((logit - ref).abs().max() / ref.abs().max()).item()
    e64 = ((logit.double() - ref64).norm() / ref64.norm()).item()
    ok = e < 5e-3
    print(f'H={H} graphs={n_graphs} rows={rows} edge={edge} masks={masks}: max rel err vs same-rounding ref {e:.2e}, '
          f'rel L2 vs fp64 reference {e64:.2e} {"OK" if ok else "FAIL"}', flush=True)
    return ok


def timing(n_graphs=196000, H=128):
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(n_graphs, seed=0).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    emb = torch.randn(b.num_nodes, H, device=dev)
    w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
    w2 = torch.randn(H, 4 * H, device=dev) / 22
    w3 = torch.randn(H, device=dev) / 11
    b3 = torch.zeros(1, device=dev)
    plan = gi.ext_plan('edge')
    print(f'E={gi.E} tiles={plan["T"]} slots/tile={gi.E / plan["T"]:.1f}', flush=True)
    for tr in (0, 1):
        for _ in range(3):
            fwd(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, tr)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fwd(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, tr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = gi.E * (24.0 * H * H + 2 * H)
        print(f'fused extractor fwd (training={tr}) E={gi.E} H={H}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    ok = True
    ok &= check(64, 40)
    ok &= check(64, 40, masks=True)
    ok &= check(128, 300, gen='ba')
    ok &= check(128, 300, masks=True, gen='ba')
    ok &= check(16, 60)
    ok &= check(80, 30, masks=True)
    ok &= check(64, 50, edge=False)
    ok &= check(128, 50, edge=False, masks=True)
    print('ALL OK' if ok else 'SOME FAILED', flush=True)
    if ok and len(sys.argv) > 1 and sys.argv[1] == 'time':
        timing()
    sys.exit(0 if ok else 1)
