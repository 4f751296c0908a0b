import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import tc
from dp_gsat_b200.data import ba2motifs_batch
from oracle import gsat_oracle as O

def run(H, ng, p, training=True, dt=torch.float32):
    b = ba2motifs_batch(ng, seed=3)
    torch.manual_seed(0)
    ext_o = O.ExtractorMLP(H, {'learn_edge_att': True, 'extractor_dropout_p': p}).to(dt)
    ext_o.train(training)
    ms = O.MaskSource(4); ext_o.masks = ms
    g = torch.Generator().manual_seed(1)
    emb = torch.relu(torch.randn(b.num_nodes, H, generator=g))
    rows = b.num_edges
    wt = torch.randn(rows, 1, generator=g)
    emb_o = emb.clone().to(dt).requires_grad_(True)
    out_o = ext_o(emb_o, b.edge_index, b.batch)
    (out_o * wt.to(dt)).sum().backward()
    mlp = ext_o.feature_extractor
    lin = [getattr(mlp, str(i)) for i in (0, 4, 8)]
    C1 = lin[0].weight.shape[0]
    gi = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda(), b.num_graphs)
    params = [t.detach().clone().float().cuda().requires_grad_(True) for t in (lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias)]
    m1 = ms.get('ext.0', (rows, C1), p).to(torch.uint8).cuda() if training else None
    m2 = ms.get('ext.1', (rows, H), p).to(torch.uint8).cuda() if training else None
    emb_g = emb.clone().cuda().requires_grad_(True)
    out_g = tc.fused_extractor(emb_g, *params, gi, edge_mode=True, pdrop=p, training=training, seed=3, mask1=m1, mask2=m2)
    (out_g * wt.cuda()).sum().backward()
    def rel(a, b_):
        a, b_ = a.detach().double().cpu(), b_.detach().double().cpu()
        return float((a - b_).norm() / b_.norm()), float((a - b_).abs().max() / b_.abs().max())
    print(f'H={H} ng={ng} p={p} train={training}: logit', rel(out_g, out_o), 'demb', rel(emb_g.grad, emb_o.grad), 'dW1', rel(params[0].grad, lin[0].weight.grad),
          'dW2', rel(params[2].grad, lin[1].weight.grad), 'dw3', rel(params[4].grad, lin[2].weight.grad))

run(64, 40, 0.5)
run(64, 40, 0.0)
run(64, 40, 0.5, training=False)
run(128, 40, 0.5)
run(64, 400, 0.5)
