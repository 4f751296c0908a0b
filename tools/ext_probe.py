"""GPU probe: fused tensor-core extractor forward vs a torch fp32 restatement on the device."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import tc
from dp_gsat_b200.data import ba2motifs_batch, molhiv_like_batch

dev = 'cuda'


def ref_forward(emb, gi, w1, w2, w3, b3, edge_mode, m1, m2, p, bf16=True):
    r = (lambda t: t.bfloat16().float()) if bf16 else (lambda t: t)
    if edge_mode:
        x = torch.cat([emb[gi.src.long()], emb[gi.dst.long()]], 1)
        seg = gi.edge_graph.long()
    else:
        x, seg = emb, gi.node_graph.long()
    Gn = gi.G

    def inorm(z):
        cnt = torch.bincount(seg, minlength=Gn).clamp(min=1).float().view(-1, 1)
        mean = torch.zeros(Gn, z.shape[1], device=dev).index_add_(0, seg, z) / cnt
        zc = z - mean[seg]
        var = torch.zeros(Gn, z.shape[1], device=dev).index_add_(0, seg, zc * zc) / cnt
        return zc / (var + 1e-5).sqrt()[seg]
    xh1 = inorm(r(x) @ r(w1).t())
    h1 = torch.relu(r(xh1))
    if m1 is not None:
        h1 = h1 * m1.float() / (1 - p)
    xh2 = inorm(r(h1) @ r(w2).t())
    h2 = torch.relu(r(xh2))
    if m2 is not None:
        h2 = h2 * m2.float() / (1 - p)
    return h2 @ w3.view(-1, 1) + b3, xh1, xh2


def run(b, H, edge_mode, p=0.5):
    torch.manual_seed(0)
    bd = b.to(dev)
    gi = G.get_graph_index(bd.edge_index, bd.batch, bd.num_graphs)
    emb = torch.relu(torch.randn(gi.N, H, device=dev))
    Kin, C1 = (2 * H, 4 * H) if edge_mode else (H, 2 * H)
    w1 = torch.randn(C1, Kin, device=dev) / Kin ** 0.5
    w2 = torch.randn(H, C1, device=dev) / C1 ** 0.5
    w3 = torch.randn(1, H, device=dev) / H ** 0.5
    b3 = torch.randn(1, device=dev)
    rows = gi.E if edge_mode else gi.N
    m1 = (torch.rand(rows, C1, device=dev) >= p).to(torch.uint8)
    m2 = (torch.rand(rows, H, device=dev) >= p).to(torch.uint8)
    out = tc.extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=edge_mode, pdrop=p, training=True, seed=1, mask1=m1, mask2=m2)
    assert out is not None, 'tile plan failed'
    logit, saved = out
    torch.cuda.synchronize()
    ref, xh1, xh2 = ref_forward(emb, gi, w1, w2, w3, b3, edge_mode, m1, m2, p)
    ref32, _, _ = ref_forward(emb, gi, w1, w2, w3, b3, edge_mode, m1, m2, p, bf16=False)
    e1 = (saved['xhat1'].float() - xh1).abs().max().item()
    e2 = (saved['xhat2'].float() - xh2).abs().max().item()
    el = (logit - ref).abs().max().item()
    el32 = (logit - ref32).abs().max().item()
    print(f'G={b.num_graphs} rows={rows} H={H} edge={edge_mode}: xhat1 err {e1:.3e} xhat2 err {e2:.3e} logit err {el:.3e} '
          f'(vs fp32 ref {el32:.3e}; |logit| max {ref.abs().max().item():.2f}) tiles={saved["plan"][2]}', flush=True)
    # hash dropout sanity: keep fraction
    out2 = tc.extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=edge_mode, pdrop=p, training=True, seed=7)
    out3 = tc.extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=edge_mode, pdrop=p, training=True, seed=7)
    assert torch.equal(out2[0], out3[0]), 'hash dropout not reproducible'


if __name__ == '__main__':
    run(ba2motifs_batch(16, seed=0), 64, True)
    run(ba2motifs_batch(300, seed=1), 128, True)
    run(molhiv_like_batch(64, seed=2), 64, False)
    run(molhiv_like_batch(64, seed=2), 80, True, p=0.3)
    # timing at cfg4 scale / 4
    b = ba2motifs_batch(49000, seed=0).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
    H = 128
    emb = torch.relu(torch.randn(gi.N, H, device=dev))
    w1 = torch.randn(4 * H, 2 * H, device=dev) / 16; w2 = torch.randn(H, 4 * H, device=dev) / 22; w3 = torch.randn(1, H, device=dev) / 11
    b3 = torch.zeros(1, device=dev)
    for _ in range(2):
        tc.extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=True, pdrop=0.5, training=True, seed=1)
    L = G._lib()
    L.timer = {'gsatb_tc_ext_fwd1': [], 'gsatb_tc_ext_fwd2': []}
    for _ in range(5):
        tc.extractor_forward(emb, gi, w1, w2, w3, b3, edge_mode=True, pdrop=0.5, training=True, seed=1)
    torch.cuda.synchronize()
    for k, v in L.timer.items():
        ms = sum(a.elapsed_time(c) for a, c in v) / len(v)
        print(f'{k}: {ms:.3f} ms for E={gi.E} (H=128) -> {gi.E / ms / 1e3:.1f} M edges/s; x4 for cfg4: {4 * ms:.2f} ms')
