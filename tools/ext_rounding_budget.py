"""Where does the bf16 extractor's distance to the fp32/fp64 reference come from?  CPU-only study (fp64 arithmetic, bf16
rounding switched on point by point) of the extractor MLP of src/utils/get_model.py:57-68 on a BA-2Motifs batch:

  x    the per-graph-centred input rows (GEMM1 B operand)      w    W1 / W2 as forward operands
  h1   Dropout(ReLU(InstanceNorm(z1))) as GEMM2's operand       wb   W1 / W2 as backward operands
  xh2  the saved xhat2        dz2, dz1  the backward's gradient tensors        xb, h1b  dW operands only

Columns: relative L2 error of  logits | d f12 | dW1 | dW2 | dw3  against the un-rounded run.
Finding: forward operand rounding alone (x, w) gives 4.8e-2 on the gradients (ReLU gate flips of near-zero
activations), every backward rounding point together 3e-3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers.ext_ref import seg_mean
from dp_gsat_b200.data import ba2motifs_batch
torch.set_num_threads(16)
def run(emb, src, dst, seg_ptr, w1, w2, w3, dlogit, mask1, mask2, pdrop, R, eps=1e-5):
    # R: set of rounding points enabled
    bf = lambda t, k: t.bfloat16().to(t.dtype) if k in R else t
    x = torch.cat([emb[src], emb[dst]], dim=1)
    m, ids = seg_mean(x, seg_ptr)
    xc = bf(x - m[ids], 'x')
    W1, W2 = bf(w1,'w'), bf(w2,'w')
    z1 = xc @ W1.t()
    v1,_ = seg_mean(z1*z1, seg_ptr)
    r1 = 1.0/torch.sqrt(v1+eps)[ids]
    xh1 = z1*r1
    s = 1.0/(1.0-pdrop)
    gate1 = (z1>0).to(x.dtype)*mask1
    h1 = bf(xh1*gate1*s, 'h1')
    z2 = h1 @ W2.t()
    mu2,_ = seg_mean(z2, seg_ptr)
    zc = z2-mu2[ids]
    v2,_ = seg_mean(zc*zc, seg_ptr)
    r2 = 1.0/torch.sqrt(v2+eps)[ids]
    xh2f = zc*r2
    xh2 = bf(xh2f, 'xh2')
    gate2 = (xh2>0).to(x.dtype)*mask2
    logit = (torch.relu(xh2f)*mask2*s) @ w3.reshape(-1,1)
    dl = dlogit.reshape(-1,1)
    dw3 = (dl*xh2*gate2*s).sum(0)
    g2 = dl*w3.reshape(1,-1)*s*gate2
    so = lambda t: seg_mean(t, seg_ptr)[0][ids]
    dz2 = bf(r2*(g2 - so(g2) - xh2*so(g2*xh2)), 'dz2')
    W2b = bf(w2,'wb'); W1b = bf(w1,'wb')
    dh1 = dz2 @ W2b
    dy = dh1*s*gate1
    dz1 = bf(r1*(dy - so(dy) - xh1*so(dy*xh1)), 'dz1')
    df12 = dz1 @ W1b
    xcb = bf(x - m[ids], 'xb') if 'x' not in R else xc
    h1b = bf(xh1*gate1*s, 'h1b') if 'h1' not in R else h1
    return logit, df12, dz1.t() @ xcb, dz2.t() @ h1b, dw3
H=64
b = ba2motifs_batch(40, seed=3)
g = torch.Generator().manual_seed(1)
emb = torch.relu(torch.randn(b.num_nodes, H, generator=g)).double()
torch.manual_seed(0)
import math
w1 = (torch.rand(4*H,2*H,dtype=torch.float64)*2-1)/math.sqrt(2*H)
w2 = (torch.rand(H,4*H,dtype=torch.float64)*2-1)/math.sqrt(4*H)
w3 = (torch.rand(H,dtype=torch.float64)*2-1)/math.sqrt(H)
E = b.num_edges
dl = torch.randn(E, generator=g).double()
m1 = (torch.rand(E,4*H,generator=g)>0.5).double(); m2=(torch.rand(E,H,generator=g)>0.5).double()
src,dst = b.edge_index[0], b.edge_index[1]
cnt = torch.bincount(b.batch[src], minlength=b.num_graphs)
seg = torch.cat([torch.zeros(1,dtype=torch.long), cnt.cumsum(0)])
ref = run(emb,src,dst,seg,w1,w2,w3,dl,m1,m2,0.5,set())
rel = lambda a,c: float((a-c).norm()/c.norm())
print('rounding points enabled | logits  d_f12  dW1  dW2  dw3')
for R in [set(), {'x','w'}, {'x','w','h1'}, {'x','w','h1','wb'}, {'x','w','h1','wb','xh2'}, {'x','w','h1','wb','dz2'}, {'x','w','h1','wb','dz1'}, {'x','w','h1','wb','xh2','dz2','dz1'}, {'wb','dz2','dz1','xb','h1b'}, {'h1'}, {'x'}, {'w'}]:
    out = run(emb,src,dst,seg,w1,w2,w3,dl,m1,m2,0.5,R)
    print(sorted(R), ' '.join(f'{rel(a,c):.2e}' for a,c in zip(out,ref)))
