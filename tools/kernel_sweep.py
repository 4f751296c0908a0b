"""BASELINE config 5: per-kernel sweep of the edge-attention kernels over E = 1e5 .. 1e8 directed edges and hidden 64 / 128 /
300: K3 aggregation fwd / bwd and K2 sampler + reverse average + info loss fwd / bwd (algorithmic GB/s against the measured
HBM peak), and the K1 extractor MLP in precision 'bf16' (fused tcgen05 kernel for H <= 128, layer-by-layer tcgen05 GEMMs for
H = 300): forward and forward + backward time, model TFLOP/s (24 H^2 E forward, 72 H^2 E with backward and dW) against the
sustained bf16 peak.   usage: python tools/kernel_sweep.py [max_edges]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200.data import ba2motifs_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    _pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    peak, tpeak = _pk['hbm_gbs'], _pk.get('bf16_tflops_sustained', 1400.0)
except Exception:
    peak, tpeak = 6650.0, 1400.0
max_e = float(sys.argv[1]) if len(sys.argv) > 1 else 1e8
dev = 'cuda'


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


print(f'# measured HBM peak {peak:.1f} GB/s; BA-2Motifs-shaped batches (25 nodes, 51 directed edges per graph)')
print('# E          H    K3 fwd ms  GB/s  frac | K3 bwd ms  GB/s  frac | K1 fwd ms  TF/s  frac | K1 f+b ms  TF/s  frac | K2 fwd ms  GB/s  frac | K2 bwd ms  GB/s  frac')
for E_target in (1e5, 1e6, 1e7, 1e8):
    if E_target > max_e:
        break
    ng = int(E_target / 51)
    b = ba2motifs_batch(ng, seed=0).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
    N, E = gi.N, gi.E
    logit = torch.randn(E, 1, device=dev, requires_grad=True)
    for H in (64, 128, 300):
        if 4.0 * N * H * 5 > 150e9:
            continue
        x = torch.randn(N, H, device=dev, requires_grad=True)
        att = torch.rand(E, 1, device=dev, requires_grad=True)
        gout = torch.randn(N, H, device=dev)
        n_it = 20 if E < 5e6 else 5
        tf = timeit(lambda: G.ops.gin_aggregate(x.detach(), att.detach(), gi, 0.0), n_it)
        out = G.ops.gin_aggregate(x, att, gi, 0.0)
        tb = timeit(lambda: torch.autograd.grad(out, [x, att], gout, retain_graph=True), n_it)
        bf, bb = 8.0 * N * H + 8.0 * E + 4.0 * N, 12.0 * N * H + 16.0 * E
        row = f'{E:<10d} {H:<4d} {tf:9.4f} {bf / tf / 1e6:6.0f} {bf / tf / 1e6 / peak:5.2f} | {tb:9.4f} {bb / tb / 1e6:6.0f} {bb / tb / 1e6 / peak:5.2f}'
        # K1: the extractor MLP on edge rows (learn_edge_att), training mode, hash dropout
        if H <= 128 or E <= 2e6:      # (the layer-by-layer path of H = 300 materialises [E, 4H] activations)
            ext = G.ExtractorMLP(H, {'learn_edge_att': True, 'extractor_dropout_p': 0.5}).to(dev)
            ext.precision = 'bf16'
            ext.train()
            emb = torch.relu(torch.randn(N, H, device=dev)).requires_grad_(True)
            n_e = max(2, n_it // 2)
            with torch.no_grad():
                te = timeit(lambda: ext(emb, b.edge_index, b.batch), n_e)

            def fb():
                for p_ in ext.parameters():
                    p_.grad = None
                o = ext(emb, b.edge_index, b.batch)
                torch.autograd.grad(o.sum(), [emb] + [p_ for p_ in ext.parameters()], allow_unused=True)
            tfb = timeit(fb, n_e)
            f1, f3 = 24.0 * H * H * E, 72.0 * H * H * E
            row += f' | {te:9.3f} {f1 / te / 1e9:6.0f} {f1 / te / 1e9 / tpeak:5.2f} | {tfb:9.3f} {f3 / tfb / 1e9:6.0f} {f3 / tfb / 1e9 / tpeak:5.2f}'
            del ext, emb
        else:
            row += ' |' + ' ' * 24 + '|' + ' ' * 24
        if H == 64:
            ts = timeit(lambda: G.ops.sample_avg_info(logit.detach(), training=True, rev=gi.rev, average=True, r=0.7, seed=1), n_it)
            a_, ea, info = G.ops.sample_avg_info(logit, training=True, rev=gi.rev, average=True, r=0.7, seed=1)
            gz = torch.randn_like(ea)
            tsb = timeit(lambda: torch.autograd.grad([ea, info], [logit], [gz, torch.ones_like(info)], retain_graph=True), n_it)
            k2f, k2b = 16.0 * E, 20.0 * E
            row += f' | {ts:9.4f} {k2f / ts / 1e6:6.0f} {k2f / ts / 1e6 / peak:5.2f} | {tsb:9.4f} {k2b / tsb / 1e6:6.0f} {k2b / tsb / 1e6 / peak:5.2f}'
        print(row, flush=True)
        del x, att, gout, out
    del b, gi, logit
    torch.cuda.empty_cache()
