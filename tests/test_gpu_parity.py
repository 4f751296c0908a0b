"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bars: index / permutation / CSR outputs bit-exact; fp32 values and gradients within rtol 1e-5 (+ atol scaled to the
magnitude of the tensor, since sums are re-ordered); whole-step loss/gradients within the documented looser bound
below because the step chains ~30 fp32 ops through two BatchNorms.  /root/reference is never read here."""
import os

import numpy as np
import pytest
import torch

from oracle import gsat_oracle as O
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def close(a, b, rtol=RTOL, atol_scale=1e-6):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    atol = atol_scale * max(1.0, float(b.abs().max())) if b.numel() else 0.0
    return torch.allclose(a, b, rtol=rtol, atol=atol)


def assert_close(a, b, rtol=RTOL, atol_scale=1e-6, what=''):
    a_, b_ = a.detach().cpu().double(), b.detach().cpu().double()
    if not close(a, b, rtol, atol_scale):
        err = (a_ - b_).abs().max().item()
        raise AssertionError(f'{what}: max abs err {err:.3e}, ref max {b_.abs().max().item():.3e}')


@pytest.fixture(scope='module')
def G():
    import dp_gsat_b200 as g
    assert torch.cuda.is_available()
    return g


def _cases():
    from dp_gsat_b200.data import ba2motifs_batch, molhiv_like_batch, load_mutag_fixture
    out = {}
    b = ba2motifs_batch(16, seed=0)
    out['ba2motifs'] = (b.edge_index, b.batch)
    m = molhiv_like_batch(24, seed=1)
    out['molhiv'] = (m.edge_index, m.batch)
    src, dst, ng = load_mutag_fixture(os.path.join(GOLDEN, 'mutag_slice.npz'))
    out['mutag'] = (torch.from_numpy(np.stack([src, dst])), torch.from_numpy(ng))
    g = torch.Generator().manual_seed(3)
    perm = torch.randperm(b.num_edges, generator=g)
    out['shuffled'] = (b.edge_index[:, perm].contiguous(), b.batch)           # unsorted edge order, still symmetric
    out['directed'] = (b.edge_index[:, b.edge_index[0] < b.edge_index[1]].contiguous(), b.batch)
    dup = torch.cat([b.edge_index, b.edge_index[:, :7]], dim=1)
    dup = dup[:, torch.argsort(b.batch[dup[0]], stable=True)].contiguous()
    out['duplicates'] = (dup, b.batch)
    out['empty'] = (torch.zeros((2, 0), dtype=torch.int64), torch.zeros(5, dtype=torch.int64))
    out['one_big_graph'] = (torch.randint(0, 3000, (2, 70000), generator=g), torch.zeros(3000, dtype=torch.int64))
    return out


@pytest.mark.parametrize('name', ['ba2motifs', 'molhiv', 'mutag', 'shuffled', 'directed', 'duplicates', 'empty',
                                  'one_big_graph'])
def test_index_build_bit_exact(G, name):
    ei, batch = _cases()[name]
    ref = O.build_index_oracle(ei, batch)
    gi = G.GraphIndex(ei.cuda(), batch.cuda())
    for k in ('src', 'dst', 'rev', 'rowptr_dst', 'eid_by_dst', 'src_by_dst', 'rowptr_src', 'eid_by_src', 'dst_by_src',
              'node_ptr', 'edge_ptr', 'edge_graph'):
        if k == 'edge_ptr' and not ref['graph_contiguous']:
            continue            # segment pointers are only defined for graph-grouped edges (flag checked below)
        assert torch.equal(getattr(gi, k).cpu(), ref[k]), f'{name}: {k} differs'
    assert gi.symmetric == ref['symmetric'] == O.is_undirected(ei)
    assert gi.has_duplicates == ref['has_dup']
    assert gi.graph_contiguous == ref['graph_contiguous']


def test_prefetched_index_equals_the_index_built_in_the_step(G):
    """prefetch_graph_index (the loader's prefetch stage: K0 + flags + extractor tile plan on a side stream, one batch
    ahead) caches exactly the bundle a direct build produces, the consumer finds it under the same key, and
    evict_graph_index drops that one entry only."""
    from dp_gsat_b200 import tc
    ei, batch = _cases()['ba2motifs']
    ei_a, b_a, ei_b, b_b = ei.cuda(), batch.cuda(), ei.cuda(), batch.cuda()
    ref = G.GraphIndex(ei_a, b_a)
    G.clear_index_cache()
    side, main = torch.cuda.Stream(), torch.cuda.current_stream()
    slots = tc.ext_tile_slots(64, True)
    gi = G.prefetch_graph_index(ei_b, b_b, 16, on_stream=side, for_stream=main, ext_plans=[('edge', slots)])
    ev = torch.cuda.Event()
    ev.record(side)
    main.wait_event(ev)
    assert G.get_graph_index(ei_b, b_b, 16) is gi                       # the step finds the prefetched entry
    assert gi._flags_host is not None and ('ext', 'edge', slots) in gi._plans       # nothing left to read back in the step
    for k in ('src', 'dst', 'rev', 'rowptr_dst', 'eid_by_dst', 'src_by_dst', 'rowptr_src', 'eid_by_src', 'dst_by_src',
              'node_ptr', 'edge_ptr', 'node_graph', 'edge_graph'):
        assert torch.equal(getattr(gi, k), getattr(ref, k)), k
    assert gi.ext_plan('edge', slots)['T'] == ref.ext_plan('edge', slots)['T']
    other = G.get_graph_index(ei_a, b_a)
    assert G.evict_graph_index(ei_b, b_b) and not G.evict_graph_index(ei_b, b_b)
    assert G.get_graph_index(ei_a, b_a) is other                        # the other batch's entry stayed
    G.clear_index_cache()


def test_mutag_reverse_is_xor1_on_gpu(G):
    ei, batch = _cases()['mutag']
    gi = G.GraphIndex(ei.cuda(), batch.cuda())
    assert torch.equal(gi.rev.cpu().long(), torch.arange(ei.shape[1]) ^ 1)


def test_reorder_like_drop_in(G):
    gold = torch.load(os.path.join(GOLDEN, 'ref_functions.pt'))
    for name in ('kat4', 'rand_a', 'rand_b'):
        ei, vals, out = (gold[f'reorder_like/{name}/{k}'] for k in ('edge_index', 'values', 'out'))
        ei_d, v_d = ei.cuda(), vals.cuda()
        assert G.is_undirected(ei_d)
        t_idx, t_val = G.transpose(ei_d, v_d, None, None, coalesced=False)
        assert torch.equal(G.reorder_like(t_idx, ei_d, t_val).cpu(), out)            # fast path (cached rev)
        t2 = torch.stack([ei_d[1], ei_d[0]])
        assert torch.equal(G.reorder_like(t2, ei_d, v_d).cpu(), out)                 # general path
    d = torch.tensor([[0, 1, 2], [1, 2, 0]]).cuda()
    assert not G.is_undirected(d)
    with pytest.raises(ValueError):
        G.reorder_like(torch.stack([d[1], d[0]]), d, torch.arange(3.).cuda())


@pytest.mark.parametrize('H', [16, 64, 80, 128, 300])
@pytest.mark.parametrize('with_att', [True, False])
def test_gin_aggregate_fwd_bwd(G, H, with_att):
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(12, seed=4)
    g = torch.Generator().manual_seed(H)
    x = torch.randn(b.num_nodes, H, generator=g)
    att = torch.rand(b.num_edges, 1, generator=g) if with_att else None
    w = torch.randn(b.num_nodes, H, generator=g)
    xr = x.clone().requires_grad_(True)
    ar = att.clone().requires_grad_(True) if with_att else None
    conv = O.GINConv(torch.nn.Identity())
    ref = conv(xr, b.edge_index, edge_atten=ar)
    (ref * w).sum().backward()
    gi = G.GraphIndex(b.edge_index.cuda(), b.batch.cuda())
    xd = x.cuda().requires_grad_(True)
    ad = att.cuda().requires_grad_(True) if with_att else None
    out = G.ops.gin_aggregate(xd, ad, gi, 0.0)
    (out * w.cuda()).sum().backward()
    assert_close(out, ref, what='fwd')
    assert_close(xd.grad, xr.grad, what='dx')
    if with_att:
        assert_close(ad.grad, ar.grad, what='datt')


@pytest.mark.parametrize('deg,H', [(500, 64), (3000, 64), (3000, 300), (1500, 128)])
def test_gin_aggregate_high_degree_and_isolated(G, deg, H):
    """A star (degree 500: many batches of one row; degree >= 1500: the row's tile holds more edges than the staged
    entry list, so the tile takes the unstaged path), isolated nodes and self loops."""
    g = torch.Generator().manual_seed(0)
    N = deg + 100
    hub = torch.zeros(deg, dtype=torch.int64)
    leaves = torch.arange(1, deg + 1)
    ei = torch.cat([torch.stack([leaves, hub]), torch.stack([hub, leaves]), torch.tensor([[7, deg + 50], [7, deg + 50]])], 1)
    x = torch.randn(N, H, generator=g)
    att = torch.rand(ei.shape[1], 1, generator=g)
    xr, ar = x.clone().requires_grad_(True), att.clone().requires_grad_(True)
    ref = O.GINConv(torch.nn.Identity())(xr, ei, edge_atten=ar)
    ref.square().sum().backward()
    gi = G.get_graph_index(ei.cuda(), None, num_nodes=N)
    xd, ad = x.cuda().requires_grad_(True), att.cuda().requires_grad_(True)
    out = G.ops.gin_aggregate(xd, ad, gi, 0.0)
    out.square().sum().backward()
    assert_close(out, ref, atol_scale=2e-6, what='fwd')
    assert_close(xd.grad, xr.grad, atol_scale=2e-6, what='dx')
    assert_close(ad.grad, ar.grad, atol_scale=2e-6, what='datt')


@pytest.mark.parametrize('mean', [False, True])
def test_pool(G, mean):
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(20, seed=2)
    x = torch.randn(b.num_nodes, 80, generator=torch.Generator().manual_seed(0))
    xr = x.clone().requires_grad_(True)
    ref = (O.global_mean_pool if mean else O.global_add_pool)(xr, b.batch)
    ref.square().sum().backward()
    gi = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda())
    xd = x.cuda().requires_grad_(True)
    out = (G.ops.global_mean_pool if mean else G.ops.global_add_pool)(xd, gi)
    out.square().sum().backward()
    assert_close(out, ref, what='pool')
    assert_close(xd.grad, xr.grad, what='dpool')


@pytest.mark.parametrize('C', [8, 64, 256, 300])
def test_instance_norm(G, C):
    g = torch.Generator().manual_seed(C)
    sizes = [1, 7, 52, 0, 130, 3]
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    batch = batch[batch != 3]
    x = torch.randn(batch.numel(), C, generator=g) * 3 + 1
    w = torch.randn(batch.numel(), C, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = O.InstanceNorm(C)(xr, batch, num_graphs=len(sizes))
    (ref * w).sum().backward()
    xd = x.cuda().requires_grad_(True)
    out = G.InstanceNorm(C)(xd, batch.cuda())
    (out * w.cuda()).sum().backward()
    assert_close(out, ref, atol_scale=2e-6, what='instnorm')
    assert_close(xd.grad, xr.grad, rtol=1e-4, atol_scale=1e-5, what='d instnorm')


@pytest.mark.parametrize('training', [True, False])
@pytest.mark.parametrize('info_on', ['att', 'edge_att'])
@pytest.mark.parametrize('tensor_r', [False, True])
def test_sample_avg_info(G, training, info_on, tensor_r):
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(10, seed=6)
    E = b.num_edges
    g = torch.Generator().manual_seed(1)
    logit = torch.randn(E, 1, generator=g) * 2
    u = torch.rand(E, 1, generator=g).clamp(1e-10, 1 - 1e-10)
    r = (torch.rand(E, 1, generator=g) * 0.8 + 0.1) if tensor_r else 0.7
    w = torch.randn(E, 1, generator=g)
    lr = logit.clone().requires_grad_(True)
    att = O.concrete_sample(lr, 1, training, u)
    ea = O.undirected_average(att, b.edge_index)
    il = O.info_loss(att if info_on == 'att' else ea, r)
    ((ea * w).sum() + 3.0 * il + (att * w.flip(0)).sum()).backward()
    gi = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda())
    ld = logit.cuda().requires_grad_(True)
    att_d, ea_d, il_d = G.ops.sample_avg_info(ld, training=training, rev=gi.rev, average=gi.symmetric,
                                               r=(r.cuda() if tensor_r else r), noise_u=u.cuda(),
                                               info_on_edge_att=(info_on == 'edge_att'))
    ((ea_d * w.cuda()).sum() + 3.0 * il_d + (att_d * w.flip(0).cuda()).sum()).backward()
    assert_close(att_d, att, what='att')
    assert_close(ea_d, ea, what='edge_att')
    assert_close(il_d, il, rtol=1e-5, what='info')
    assert_close(ld.grad, lr.grad, rtol=2e-5, atol_scale=2e-6, what='dlogit')


def test_sampler_philox_is_regenerable_and_uniform(G):
    lg = torch.zeros(200000, 1, device='cuda')
    a1 = G.concrete_sample(lg, 1, True, seed=5, offset=100)
    a2 = G.concrete_sample(lg, 1, True, seed=5, offset=100)
    a3 = G.concrete_sample(lg, 1, True, seed=6, offset=100)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)
    # att = sigmoid(logit(u)) = u  -> uniform on (0,1)
    assert abs(float(a1.mean()) - 0.5) < 5e-3 and abs(float(a1.var()) - 1 / 12) < 5e-3
    assert float(a1.min()) > 0 and float(a1.max()) < 1


def test_lift(G):
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(12, seed=5)
    g = torch.Generator().manual_seed(2)
    na = torch.rand(b.num_nodes, 1, generator=g)
    w = torch.randn(b.num_edges, 1, generator=g)
    nr = na.clone().requires_grad_(True)
    ref = O.lift_node_att_to_edge_att(nr, b.edge_index)
    (ref * w).sum().backward()
    nd = na.cuda().requires_grad_(True)
    out = G.lift_node_att_to_edge_att(nd, b.edge_index.cuda(), b.batch.cuda())
    (out * w.cuda()).sum().backward()
    assert_close(out, ref, what='lift')
    assert_close(nd.grad, nr.grad, what='dlift')


def _build_pair(G, batch, hidden, n_layers, learn_edge_att, p_drop, p_ext, info_on, seed=0):
    cfg = {'model_name': 'GIN', 'hidden_size': hidden, 'n_layers': n_layers, 'dropout_p': p_drop,
           'use_edge_attr': False}
    shared = {'learn_edge_att': learn_edge_att, 'extractor_dropout_p': p_ext}
    torch.manual_seed(seed)
    clf_o = O.get_model(batch.x.shape[1], 0, 2, False, cfg)
    ext_o = O.ExtractorMLP(hidden, shared)
    clf_g = G.get_model(batch.x.shape[1], 0, 2, False, cfg, 'cuda')
    ext_g = G.ExtractorMLP(hidden, shared).cuda()
    clf_g.load_state_dict(clf_o.state_dict())          # identical state_dict keys are part of the contract
    ext_g.load_state_dict(ext_o.state_dict())
    ms = O.MaskSource(2)
    for m in (clf_o, ext_o, clf_g, ext_g):
        m.masks = ms
    go = O.GSAT(clf_o, ext_o, O.Criterion(2, False), learn_edge_att=learn_edge_att, final_r=0.5, info_on=info_on)
    gg = G.GSAT(clf_g, ext_g, G.Criterion(2, False), learn_edge_att=learn_edge_att, final_r=0.5, info_on=info_on)
    return go, gg


@pytest.mark.parametrize('cfgname', ['cfg1_L2', 'cfg1_L3', 'mutag_dual_avg', 'lift_path', 'fork_info_on_edge_att',
                                     'eval_mode', 'cfg1_randx_strict'])
def test_gsat_step_parity(G, cfgname):
    """Whole step: forward_pass + backward on identical weights, noise and dropout masks.
    Documented bound: rtol 2e-4 / atol 2e-5*scale on loss, logits, attention and every parameter gradient (fp32,
    ~30 chained ops incl. two BatchNorms per layer; summation order differs from the CPU's)."""
    from dp_gsat_b200.data import (ba2motifs_batch, load_mutag_fixture, line_graph_dual, graph_contiguous_relabel,
                                   batch_from_edge_list)
    training, learn, info_on, L, H = True, True, 'att', 2, 64
    if cfgname.startswith('cfg1'):
        b = ba2motifs_batch(128, seed=0)
        L = 3 if cfgname.endswith('L3') else 2
        if cfgname == 'cfg1_randx_strict':      # non-constant node features: well-conditioned BatchNorm statistics, so
            b.x = torch.rand(b.x.shape, generator=torch.Generator().manual_seed(5))    # only the strict bound applies
    elif cfgname == 'mutag_dual_avg':
        src, dst, ng = load_mutag_fixture(os.path.join(GOLDEN, 'mutag_slice.npz'))
        keep = ng[src] < 128
        ds, dd, dng = line_graph_dual(src[keep], dst[keep], ng)
        ds, dd = graph_contiguous_relabel(ds, dd, dng)
        b = batch_from_edge_list(ds, dd, dng, x_dim=31, seed=0)
    else:
        b = ba2motifs_batch(32, seed=1)
        learn = cfgname != 'lift_path'
        info_on = 'edge_att' if cfgname == 'fork_info_on_edge_att' else 'att'
        training = cfgname != 'eval_mode'
    go, gg = _build_pair(G, b, H, L, learn, 0.3, 0.5, info_on)
    import copy
    go64 = copy.deepcopy(go).double()          # fp64 ground truth: same weights, noise and masks
    for m in (go, gg, go64):
        m.train(training)
    n_noise = b.num_edges if learn else b.num_nodes
    u = torch.rand(n_noise, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10)
    ea_o, loss_o, ld_o, logit_o = go.forward_pass(b, 12, training, noise_u=u)
    b64 = b.to('cpu')
    b64.x = b64.x.double()
    ea_t, loss_t, ld_t, logit_t = go64.forward_pass(b64, 12, training, noise_u=u.double())
    bd = b.to('cuda')
    ea_g, loss_g, ld_g, logit_g = gg.forward_pass(bd, 12, training, noise_u=u.cuda())

    def check(g_val, o_val, t_val, what, rtol=2e-4, atol_scale=2e-5):
        """Pass if within the documented fp32 bound of the fp32 oracle, or if the CUDA result is at least as close
        to the fp64 ground truth as 4x the fp32 oracle's own error (ill-conditioned cases: constant BA-2Motifs
        features make BatchNorm variances tiny, which amplifies fp32 rounding in BOTH fp32 implementations)."""
        if close(g_val, o_val, rtol, atol_scale):
            return
        assert cfgname != 'cfg1_randx_strict', f'{what}: outside rtol {rtol} of the fp32 oracle (strict case: no fp64 criterion)'
        t = t_val.detach().cpu().double()
        err_g = (g_val.detach().cpu().double() - t).abs().max().item()
        err_o = (o_val.detach().cpu().double() - t).abs().max().item()
        if err_g <= 4 * err_o + 1e-7 * max(1.0, t.abs().max().item()):
            return
        # Constant node features (BA-2Motifs: x = 0.1 everywhere) make every node of one degree carry EXACTLY the same
        # activations, and channels whose pre-activation barely varies put whole degree classes of BatchNorm outputs
        # within rounding distance of the ReLU kink (beta = 0).  Which side a class lands on is decided by the last bit of
        # the arithmetic, so two correct fp32 implementations differ by O(1e-3) in the gradients that pass those units
        # (tools/cfg1_probe.py, tools/cfg1_probe2.py: forward values agree to 1e-6, the gradient difference is the same
        # whichever dense layer is swapped, and it vanishes with non-constant features -- the 'cfg1_randx_strict' case,
        # which must hold the plain rtol).  For these cases the bar is a relative L2 distance to the fp64 oracle.
        assert cfgname.startswith('cfg1') and what.startswith('grad'), \
            f'{what}: cuda-vs-fp64 {err_g:.3e} > 4 x oracle32-vs-fp64 {err_o:.3e} (ref max {t.abs().max().item():.3e})'
        rl2 = float((g_val.detach().cpu().double() - t).norm() / t.norm().clamp_min(1e-30))
        assert rl2 <= 5e-3, f'{what}: relative L2 vs fp64 {rl2:.3e} (constant-feature case)'

    check(ea_g, ea_o, ea_t, 'edge_att')
    check(logit_g, logit_o, logit_t, 'clf_logits')
    check(loss_g, loss_o, loss_t, 'loss')
    assert ld_g['info'] == pytest.approx(ld_o['info'], rel=2e-4, abs=1e-6)
    if training:
        loss_o.backward()
        loss_t.backward()
        loss_g.backward()
        named = lambda m: dict(list(m.clf.named_parameters()) + [('ext.' + k, v) for k, v in m.extractor.named_parameters()])
        po, pt, pg = named(go), named(go64), named(gg)
        assert po.keys() == pg.keys()
        for k in po:
            if po[k].grad is None:
                assert pg[k].grad is None or float(pg[k].grad.abs().max()) == 0.0
                continue
            check(pg[k].grad, po[k].grad, pt[k].grad, f'grad {k}', rtol=1e-3, atol_scale=2e-4)
        # BatchNorm running statistics are updated twice per step, in call order (get_emb, then clf)
        for k, v in go.clf.state_dict().items():
            if 'running' in k or 'num_batches' in k:
                check(gg.clf.state_dict()[k].float(), v.float(), go64.clf.state_dict()[k].double(), k)


def test_fork_glue_composes_with_autograd(G):
    """Fork step pieces (src/run_gsat.py:222-253, :126-132): gumbel_sigmoid at tau 0.1, f1 loss, epoch>50 mixing,
    per-edge tensor r -- plain torch ops composed with the CUDA autograd ops."""
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(8, seed=9)
    g = torch.Generator().manual_seed(0)
    E = b.num_edges
    logit_p, logit_d = torch.randn(E, 1, generator=g), torch.randn(E, 1, generator=g)
    u, U = torch.rand(E, 1, generator=g).clamp(1e-6, 1 - 1e-6), torch.rand(E, 1, generator=g)
    x = torch.randn(b.num_nodes, 16, generator=g)

    def run(mod, ops_side, dev):
        lp = logit_p.clone().to(dev).requires_grad_(True)
        ld_ = logit_d.clone().to(dev).requires_grad_(True)
        dual_att = mod.gumbel_sigmoid(ld_, tau=0.1, noise_u=U.to(dev))
        f1 = mod.f1_sparsity_loss(dual_att, b.edge_label.to(dev))
        att, edge_att, il = ops_side(lp, ld_.sigmoid().detach(), dev)
        mixed = 0.3 * dual_att + 0.7 * edge_att
        out = ops_side.conv(x.to(dev), mixed, dev)
        (out.square().mean() + il + f1).backward()
        return lp.grad, ld_.grad, out

    class OracleSide:
        def __call__(self, lp, r, dev):
            att = O.concrete_sample(lp, 1, True, u)
            ea = O.undirected_average(att, b.edge_index)
            return att, ea, O.info_loss(ea, r)

        def conv(self, xx, a, dev):
            return O.GINConv(torch.nn.Identity())(xx, b.edge_index, edge_atten=a)

    class CudaSide:
        def __init__(self):
            self.gi = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda())

        def __call__(self, lp, r, dev):
            return G.ops.sample_avg_info(lp, training=True, rev=self.gi.rev, average=True, r=r, noise_u=u.cuda(),
                                         info_on_edge_att=True)

        def conv(self, xx, a, dev):
            return G.ops.gin_aggregate(xx, a, self.gi, 0.0)

    gp_o, gd_o, out_o = run(O, OracleSide(), 'cpu')
    gp_g, gd_g, out_g = run(G, CudaSide(), 'cuda')
    assert_close(out_g, out_o, rtol=1e-4, atol_scale=1e-5, what='conv out')
    assert_close(gp_g, gp_o, rtol=1e-4, atol_scale=1e-5, what='d primal logits')
    assert_close(gd_g, gd_o, rtol=1e-4, atol_scale=1e-5, what='d dual logits')


def test_large_batch_properties(G):
    """BASELINE-size properties that need no oracle pass: cfg4-like index (2 M edges here, bounded for test time)
    is an involution with consistent CSR/CSC; aggregation is linear in x and equals a dense check on a sample."""
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(40000, seed=11).to('cuda')
    gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
    assert gi.symmetric and gi.graph_contiguous and not gi.has_duplicates
    rev = gi.rev.long()
    E = gi.E
    ar = torch.arange(E, device='cuda')
    assert torch.equal(rev[rev], ar)                                                        # involution
    assert torch.equal(b.edge_index[0][rev], b.edge_index[1]) and torch.equal(b.edge_index[1][rev], b.edge_index[0])
    assert torch.equal(torch.sort(gi.eid_by_dst.long()).values, ar)                         # permutations
    assert torch.equal(torch.sort(gi.eid_by_src.long()).values, ar)
    d_sorted = b.edge_index[1][gi.eid_by_dst.long()]
    assert bool((d_sorted[1:] >= d_sorted[:-1]).all())                                      # sortedness
    assert torch.equal(torch.bincount(b.edge_index[1], minlength=gi.N),
                       (gi.rowptr_dst[1:] - gi.rowptr_dst[:-1]).long())
    H = 128
    x1, x2 = torch.randn(gi.N, H, device='cuda'), torch.randn(gi.N, H, device='cuda')
    att = torch.rand(E, 1, device='cuda')
    f = lambda t: G.ops.gin_aggregate(t, att, gi, 0.0)
    assert torch.allclose(f(x1 + 2 * x2), f(x1) + 2 * f(x2), rtol=1e-4, atol=1e-4)           # linearity
    ref = torch.zeros_like(x1).index_add_(0, b.edge_index[1], x1[b.edge_index[0]] * att) + x1
    assert torch.allclose(f(x1), ref, rtol=1e-4, atol=1e-4)                                  # checksum vs torch on device
    assert torch.equal(f(x1), f(x1))                                                         # run-to-run determinism


# ---------------------------------------------------------------------------------------------------------------
# K4  PNA (BASELINE config 3: molhiv-shaped batches, edge features, multi-aggregator)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('H,with_ea,with_att', [(16, False, True), (80, False, False), (80, True, True), (32, True, False)])
def test_pna_aggregate_fwd_bwd(G, H, with_ea, with_att):
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(24, seed=7)
    g = torch.Generator().manual_seed(H)
    N, E = b.num_nodes, b.num_edges
    x = torch.randn(N, H, generator=g)
    ea = torch.randn(E, H, generator=g) if with_ea else None
    att = torch.rand(E, 1, generator=g) if with_att else None
    aggs = ['mean', 'min', 'max', 'std', 'sum', 'var']
    F_ = (3 if with_ea else 2) * H
    w = torch.randn(N, len(aggs) * F_, generator=g)

    def oracle(xr, er, ar):
        src, dst = b.edge_index
        m = torch.cat([xr[dst], xr[src]] + ([er] if er is not None else []), dim=-1)
        if ar is not None:
            m = m * ar
        return torch.cat([O.AGGREGATORS[a](m, dst, N) for a in aggs], dim=-1)
    def run_oracle(dt):
        xr = x.clone().to(dt).requires_grad_(True)
        er = ea.clone().to(dt).requires_grad_(True) if with_ea else None
        ar = att.clone().to(dt).requires_grad_(True) if with_att else None
        ref = oracle(xr, er, ar)
        (ref * w.to(dt)).sum().backward()
        return ref, xr.grad, (er.grad if with_ea else None), (ar.grad if with_att else None)
    r32, r64 = run_oracle(torch.float32), run_oracle(torch.float64)
    gi = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda(), b.num_graphs)
    xd = x.cuda().requires_grad_(True)
    ed = ea.cuda().requires_grad_(True) if with_ea else None
    ad = att.cuda().requires_grad_(True) if with_att else None
    out = G.ops.pna_aggregate(xd, ed, ad, gi, aggs)
    (out * w.cuda()).sum().backward()

    def check(g_val, o_val, t_val, what):
        """rtol 1e-5 against the fp32 oracle, or -- var/std evaluate E[m^2] - E[m]^2 in fp32, which cancels -- at
        least as close to the fp64 ground truth as 4x the fp32 oracle itself."""
        if close(g_val, o_val, 1e-5, 2e-6):
            return
        t = t_val.detach().cpu().double()
        err_g = (g_val.detach().cpu().double() - t).abs().max().item()
        err_o = (o_val.detach().cpu().double() - t).abs().max().item()
        assert err_g <= 4 * err_o + 1e-6 * max(1.0, t.abs().max().item()), f'{what}: {err_g:.3e} vs oracle32 {err_o:.3e}'
    got = (out, xd.grad, ed.grad if with_ea else None, ad.grad if with_att else None)
    for name, gv, ov, tv in zip(('pna fwd', 'pna dx', 'pna d edge_feat', 'pna d att'), got, r32, r64):
        if gv is not None:
            check(gv, ov, tv, name)


def test_pna_empty_rows_and_constant_segments(G):
    """torch_scatter semantics: empty rows give 0 for mean/min/max/sum and sqrt(1e-5) for std."""
    ei = torch.tensor([[0, 1, 1], [2, 2, 0]])
    x = torch.ones(4, 4)
    gi = G.get_graph_index(ei.cuda(), None, num_nodes=4)
    out = G.ops.pna_aggregate(x.cuda(), None, None, gi, ['mean', 'min', 'max', 'std']).cpu()
    F_ = 8
    assert torch.equal(out[3, :3 * F_], torch.zeros(3 * F_)) and torch.equal(out[1, :3 * F_], torch.zeros(3 * F_))
    assert torch.allclose(out[:, 3 * F_:], torch.full((4, F_), 1e-5).sqrt())
    assert torch.equal(out[2, :F_], torch.ones(F_))


@pytest.mark.parametrize('use_edge_attr,learn_edge_att', [(False, False), (True, False), (True, True)])
def test_gsat_pna_step_parity(G, use_edge_attr, learn_edge_att):
    """BASELINE config 3: GSAT + PNA on molhiv-shaped batches (atom / bond encoders, mean-min-max-std, lift path of
    the reference yml and the edge-attention path), whole step against the oracle."""
    import copy
    from dp_gsat_b200.data import molhiv_like_batch, in_degree_histogram
    b = molhiv_like_batch(64, seed=3, with_edge_attr=use_edge_attr)
    cfg = {'model_name': 'PNA', 'hidden_size': 80, 'n_layers': 4, 'dropout_p': 0.3, 'atom_encoder': True,
           'use_edge_attr': use_edge_attr, 'aggregators': ['mean', 'min', 'max', 'std'], 'scalers': False,
           'deg': in_degree_histogram(b)}
    shared = {'learn_edge_att': learn_edge_att, 'extractor_dropout_p': 0.5}
    ea_dim = 3 if use_edge_attr else 0
    torch.manual_seed(0)
    clf_o, ext_o = O.get_model(9, ea_dim, 2, False, cfg), O.ExtractorMLP(80, shared)
    clf_g, ext_g = G.get_model(9, ea_dim, 2, False, cfg, 'cuda'), G.ExtractorMLP(80, shared).cuda()
    clf_g.load_state_dict(clf_o.state_dict())
    ext_g.load_state_dict(ext_o.state_dict())
    ms = O.MaskSource(2)
    for m in (clf_o, ext_o, clf_g, ext_g):
        m.masks = ms
    go = O.GSAT(clf_o, ext_o, O.Criterion(2, False), learn_edge_att=learn_edge_att, final_r=0.7)
    gg = G.GSAT(clf_g, ext_g, G.Criterion(2, False), learn_edge_att=learn_edge_att, final_r=0.7)
    go64 = copy.deepcopy(go).double()
    for m in (go, gg, go64):
        m.train()
    n_noise = b.num_edges if learn_edge_att else b.num_nodes
    u = torch.rand(n_noise, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10)
    ea_o, loss_o, _, logit_o = go.forward_pass(b, 5, True, noise_u=u)
    ea_t, loss_t, _, logit_t = go64.forward_pass(b, 5, True, noise_u=u.double())
    ea_g, loss_g, _, logit_g = gg.forward_pass(b.to('cuda'), 5, True, noise_u=u.cuda())
    loss_o.backward()
    loss_t.backward()
    loss_g.backward()

    def check(g_val, o_val, t_val, what, rtol=2e-4, atol_scale=2e-5):
        if close(g_val, o_val, rtol, atol_scale):
            return
        t = t_val.detach().cpu().double()
        err_g = (g_val.detach().cpu().double() - t).abs().max().item()
        err_o = (o_val.detach().cpu().double() - t).abs().max().item()
        assert err_g <= 4 * err_o + 1e-7 * max(1.0, t.abs().max().item()), \
            f'{what}: cuda-vs-fp64 {err_g:.3e} > 4 x oracle32-vs-fp64 {err_o:.3e}'
    check(ea_g, ea_o, ea_t, 'edge_att')
    check(logit_g, logit_o, logit_t, 'logits')
    check(loss_g, loss_o, loss_t, 'loss')
    # Gradients: PNA's std aggregator evaluates sqrt(relu(E[m^2] - E[m]^2) + 1e-5) in fp32 (as the reference does);
    # the subtraction cancels and the relu gate / 1/std factor amplify it, so single elements of a gradient can be off
    # by tens of percent in EITHER fp32 implementation (measured: fp32 oracle up to 2.5e-2 relative L2 vs fp64 on some
    # parameters, CUDA up to 1.7e-2 on others).  Bar: relative L2 <= 5e-2 against the fp64 oracle for every parameter,
    # and the median over parameters within 3x of the fp32 oracle's own median error.  The kernel itself is held to
    # rtol 1e-5 / the fp64 criterion in test_pna_aggregate_fwd_bwd.
    named = lambda m: dict(list(m.clf.named_parameters()) + [('ext.' + k, v) for k, v in m.extractor.named_parameters()])
    po, pt, pg = named(go), named(go64), named(gg)
    assert po.keys() == pg.keys()
    errs_g, errs_o = [], []
    for k in po:
        if po[k].grad is None:
            continue
        t = pt[k].grad.double()
        if float(t.abs().max()) < 1e-12:      # biases in front of a norm: analytically zero
            continue
        eg = float((pg[k].grad.cpu().double() - t).norm() / t.norm())
        eo = float((po[k].grad.double() - t).norm() / t.norm())
        assert eg <= 5e-2, f'grad {k}: relative L2 vs fp64 oracle {eg:.3e} (fp32 oracle: {eo:.3e})'
        errs_g.append(eg)
        errs_o.append(eo)
    med = lambda v: sorted(v)[len(v) // 2]
    assert med(errs_g) <= 3 * med(errs_o) + 1e-4, (med(errs_g), med(errs_o))


def test_cuda_graph_training_step(G):
    """parallel.TrainStep.enable_cuda_graph: the whole step (forward_pass + backward + Adam) replays as one CUDA graph;
    every replay advances the device step counter, draws fresh noise / dropout and updates the parameters."""
    from dp_gsat_b200.data import ba2motifs_batch
    from dp_gsat_b200.parallel import TrainStep
    b = ba2motifs_batch(64, seed=5).to('cuda')
    cfg = {'model_name': 'GIN', 'hidden_size': 64, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    shared = {'learn_edge_att': True, 'extractor_dropout_p': 0.5}
    torch.manual_seed(0)
    clf, ext = G.get_model(10, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(64, shared).cuda()
    clf.precision = ext.precision = 'bf16'
    gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5, lazy_metrics=True)
    gsat.train()
    step = TrainStep(gsat, lr=1e-3)
    assert step.enable_cuda_graph(b, 0, warmup=2)
    c0 = int(gsat.step_counter.item())
    w0 = clf.convs[0].nn[0].weight.detach().clone()
    losses, atts = [], []
    for _ in range(3):
        edge_att, loss, _, _ = step(b, 0)
        torch.cuda.synchronize()
        losses.append(float(loss))
        atts.append(edge_att.detach().clone())
    assert int(gsat.step_counter.item()) == c0 + 3
    assert all(np.isfinite(l) for l in losses)
    assert not torch.equal(atts[0], atts[1]) and not torch.equal(atts[1], atts[2])      # fresh concrete noise per replay
    assert not torch.equal(w0, clf.convs[0].nn[0].weight)                               # Adam stepped inside the graph
    # averaged attention stays symmetric under replay: edge_att[e] == edge_att[rev[e]]
    gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
    assert torch.equal(atts[2].view(-1), atts[2].view(-1)[gi.rev.long()])
    step.disable_cuda_graph()
    _, loss_e, _, _ = step(b, 0)                                                       # eager path still works afterwards
    assert np.isfinite(float(loss_e))


def test_graph_step_load_batch_refreshes_the_resident_batch(G):
    """TrainStep.load_batch: a fresh batch of the captured shape (here: other features, labels and edge order inside the
    graphs) and the index K0 built for it are copied into the static buffers the step graph reads, so a streaming loader
    keeps the one-launch step; a batch that does not fit is refused untouched."""
    from dp_gsat_b200.data import ba2motifs_batch, Batch
    from dp_gsat_b200.parallel import TrainStep
    from dp_gsat_b200 import tc
    b = ba2motifs_batch(64, seed=5).to('cuda')
    cfg = {'model_name': 'GIN', 'hidden_size': 64, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    shared = {'learn_edge_att': True, 'extractor_dropout_p': 0.5}
    torch.manual_seed(0)
    clf, ext = G.get_model(10, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(64, shared).cuda()
    clf.precision = ext.precision = 'bf16'
    gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5, lazy_metrics=True)
    gsat.train()
    step = TrainStep(gsat, lr=1e-3)
    assert step.load_batch(b) is False                                   # no graph yet
    assert step.enable_cuda_graph(b, 0, warmup=2)
    _, loss0, _, _ = step(b, 0)
    loss0 = float(loss0)
    # the same graphs with each graph's edge list reversed (still grouped by graph, still symmetric), new features / labels
    gi0 = G.GraphIndex(b.edge_index.clone(), b.batch.clone(), b.num_graphs)
    ep = gi0.edge_ptr.long()
    e = torch.arange(b.num_edges, device='cuda')
    g_of_e = gi0.edge_graph.long()
    perm = ep[g_of_e] + (ep[g_of_e + 1] - 1 - e)
    torch.manual_seed(1)
    b2 = Batch(torch.randn_like(b.x), b.edge_index[:, perm].contiguous(), b.batch.clone(), 1 - b.y, None,
               None if b.edge_label is None else b.edge_label[perm].contiguous(), b.num_graphs)
    side = torch.cuda.Stream()
    gi2 = G.prefetch_graph_index(b2.edge_index, b2.batch, b2.num_graphs, on_stream=side,
                                 for_stream=torch.cuda.current_stream(), ext_plans=[('edge', tc.ext_tile_slots(64, True))])
    torch.cuda.current_stream().wait_stream(side)
    assert step.load_batch(b2, gi2) is True
    assert torch.equal(b.x, b2.x) and torch.equal(b.edge_index, b2.edge_index) and torch.equal(b.y, b2.y)
    ref = G.GraphIndex(b2.edge_index.clone(), b2.batch.clone(), b2.num_graphs)
    for k in ('src', 'dst', 'rev', 'rowptr_dst', 'eid_by_dst', 'src_by_dst', 'rowptr_src', 'eid_by_src', 'dst_by_src',
              'node_ptr', 'edge_ptr', 'node_graph', 'edge_graph'):
        assert torch.equal(getattr(step._graph_index, k), getattr(ref, k)), k
    edge_att, loss1, _, _ = step(b, 0)                                    # ONE replay, now on b2's contents
    torch.cuda.synchronize()
    assert np.isfinite(float(loss1)) and float(loss1) != loss0
    assert torch.equal(edge_att.view(-1), edge_att.view(-1)[ref.rev.long()])          # symmetric under b2's reverse map
    small = ba2motifs_batch(32, seed=5).to('cuda')
    assert step.load_batch(small) is False                                # other counts: refused, nothing copied
    assert torch.equal(b.x, b2.x)
    step.disable_cuda_graph()


@pytest.mark.parametrize('N,F_,H', [(1000, 10, 64), (33333, 14, 128), (777, 9, 80), (5, 1, 4)])
def test_small_linear_weight_gradient(G, N, F_, H):
    """Node-encoder Linear (src/models/gin.py:22-25): own K = N weight / bias gradient kernel against autograd."""
    g = torch.Generator().manual_seed(N)
    x = torch.randn(N, F_, generator=g)
    lin = torch.nn.Linear(F_, H)
    w = torch.randn(N, H, generator=g)
    (lin(x) * w).sum().backward()
    lg = torch.nn.Linear(F_, H).cuda()
    lg.load_state_dict(lin.state_dict())
    xg = x.cuda().requires_grad_(True)
    out = G.ops.small_linear(xg, lg.weight, lg.bias)
    (out * w.cuda()).sum().backward()
    assert close(out, lin(x), 1e-5, 1e-6)
    assert close(lg.weight.grad, lin.weight.grad, 1e-4, 1e-5)
    assert close(lg.bias.grad, lin.bias.grad, 1e-4, 1e-5)
    assert close(xg.grad, w @ lin.weight.detach(), 1e-4, 1e-5)


@pytest.mark.parametrize('case', ['kat4', 'mutag', 'ba2motifs', 'shuffled', 'directed', 'empty'])
@pytest.mark.parametrize('halve', [False, True])
def test_line_graph_dual_bit_exact(G, case, halve):
    """GPU line-graph builder against the restated reference loops: bit-exact dual edge list
    (same order) and dual batch vector."""
    from dp_gsat_b200.data import ba2motifs_batch, load_mutag_fixture
    if case == 'kat4':
        ei = torch.tensor([[0, 1, 0, 2, 1, 3, 0, 3, 1, 2], [1, 0, 2, 0, 3, 1, 3, 0, 2, 1]])
        batch = torch.zeros(4, dtype=torch.int64)
    elif case == 'mutag':
        src, dst, ng = load_mutag_fixture(os.path.join(GOLDEN, 'mutag_slice.npz'))
        keep = ng[src] < 96
        ei, batch = torch.from_numpy(np.stack([src[keep], dst[keep]])), torch.from_numpy(ng)
    elif case == 'empty':
        ei, batch = torch.zeros((2, 0), dtype=torch.int64), torch.zeros(3, dtype=torch.int64)
    else:
        b = ba2motifs_batch(60, seed=4)
        ei, batch = b.edge_index, b.batch
        if case == 'shuffled':       # rows shuffled in pairs: groups no longer appear in node order
            perm = torch.randperm(ei.shape[1] // 2, generator=torch.Generator().manual_seed(0))
            ei = ei[:, torch.stack([2 * perm, 2 * perm + 1], 1).reshape(-1)]
        elif case == 'directed':
            ei = ei[:, ::2].contiguous()
    if halve and (case == 'directed' or ei.shape[1] % 2):
        pytest.skip('halve needs both directions of every edge as consecutive rows')
    exp_ei, exp_b = O.line_graph_dual(ei, batch, halve=halve)
    got_ei, got_b = G.line_graph_dual(ei.cuda(), batch.cuda(), halve=halve)
    assert got_ei.dtype == torch.int64 and got_b.dtype == torch.int64
    assert torch.equal(got_ei.cpu(), exp_ei)
    assert torch.equal(got_b.cpu(), exp_b)


# ---------------------------------------------------------------------------------------------------------------
# GINEConv: GIN with edge features
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('H,with_att', [(16, True), (64, True), (128, False), (300, True)])
def test_gine_aggregate_fwd_bwd(G, H, with_att):
    """relu(x_j + edge_feat) * edge_atten summed over incoming edges + (1+eps) x, forward and all three gradients,
    against the oracle's GINEConv (reference conv_layers.py:37-66) with an identity nn and no lin."""
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(40, seed=9)
    g = torch.Generator().manual_seed(H)
    N, E = b.num_nodes, b.num_edges
    x, ef = torch.randn(N, H, generator=g), torch.randn(E, H, generator=g)
    att = torch.rand(E, 1, generator=g) if with_att else None
    w = torch.randn(N, H, generator=g)
    xr, er = x.clone().requires_grad_(True), ef.clone().requires_grad_(True)
    ar = att.clone().requires_grad_(True) if with_att else None
    ref = O.GINEConv(torch.nn.Identity())(xr, b.edge_index, edge_attr=er, edge_atten=ar)
    (ref * w).sum().backward()
    gi = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda(), b.num_graphs)
    xd, ed = x.cuda().requires_grad_(True), ef.cuda().requires_grad_(True)
    ad = att.cuda().requires_grad_(True) if with_att else None
    out = G.ops.gine_aggregate(xd, ed, ad, gi, 0.0)
    (out * w.cuda()).sum().backward()
    assert_close(out, ref, atol_scale=2e-6, what='gine fwd')
    assert_close(xd.grad, xr.grad, atol_scale=2e-6, what='gine dx')
    assert_close(ed.grad, er.grad, atol_scale=2e-6, what='gine d edge_feat')
    if with_att:
        assert_close(ad.grad, ar.grad, atol_scale=2e-6, what='gine d att')
    assert torch.equal(out, G.ops.gine_aggregate(xd, ed, ad, gi, 0.0))        # deterministic


@pytest.mark.parametrize('atom_encoder', [True, False])
def test_gsat_gin_with_edge_features_step_parity(G, atom_encoder):
    """GSAT + GIN on a batch WITH edge features (reference gin.py:28-38: edge_encoder + GINEConv layers), whole step
    against the oracle: state_dict keys (incl. convs.{l}.lin.*), edge attention, logits, loss, every gradient."""
    import copy
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(48, seed=11, with_edge_attr=True)
    if not atom_encoder:      # plain float features through Linear encoders
        gen = torch.Generator().manual_seed(0)
        b.x = torch.rand(b.num_nodes, 7, generator=gen)
        b.edge_attr = torch.rand(b.num_edges, 5, generator=gen)
    cfg = {'model_name': 'GIN', 'hidden_size': 64, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': True,
           'atom_encoder': atom_encoder}
    shared = {'learn_edge_att': True, 'extractor_dropout_p': 0.5}
    x_dim, ea_dim = b.x.shape[1], b.edge_attr.shape[1]
    torch.manual_seed(0)
    clf_o, ext_o = O.get_model(x_dim, ea_dim, 2, False, cfg), O.ExtractorMLP(64, shared)
    clf_g, ext_g = G.get_model(x_dim, ea_dim, 2, False, cfg, 'cuda'), G.ExtractorMLP(64, shared).cuda()
    assert set(clf_g.state_dict().keys()) == set(clf_o.state_dict().keys())
    assert 'convs.0.lin.weight' in clf_g.state_dict() and 'convs.1.eps' in clf_g.state_dict()
    clf_g.load_state_dict(clf_o.state_dict())
    ext_g.load_state_dict(ext_o.state_dict())
    ms = O.MaskSource(2)
    for m in (clf_o, ext_o, clf_g, ext_g):
        m.masks = ms
    go = O.GSAT(clf_o, ext_o, O.Criterion(2, False), learn_edge_att=True, final_r=0.7)
    gg = G.GSAT(clf_g, ext_g, G.Criterion(2, False), learn_edge_att=True, final_r=0.7)
    go64 = copy.deepcopy(go).double()
    for m in (go, gg, go64):
        m.train()
    u = torch.rand(b.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10)
    ea_o, loss_o, _, logit_o = go.forward_pass(b, 3, True, noise_u=u)
    b64 = b.to('cpu')
    if not atom_encoder:
        b64.x, b64.edge_attr = b64.x.double(), b64.edge_attr.double()
    ea_t, loss_t, _, logit_t = go64.forward_pass(b64, 3, True, noise_u=u.double())
    ea_g, loss_g, _, logit_g = gg.forward_pass(b.to('cuda'), 3, True, noise_u=u.cuda())
    loss_o.backward()
    loss_t.backward()
    loss_g.backward()

    def check(g_val, o_val, t_val, what, rtol=2e-4, atol_scale=2e-5):
        if close(g_val, o_val, rtol, atol_scale):
            return
        t = t_val.detach().cpu().double()
        err_g = (g_val.detach().cpu().double() - t).abs().max().item()
        err_o = (o_val.detach().cpu().double() - t).abs().max().item()
        assert err_g <= 4 * err_o + 1e-7 * max(1.0, t.abs().max().item()), \
            f'{what}: cuda-vs-fp64 {err_g:.3e} > 4 x oracle32-vs-fp64 {err_o:.3e}'
    check(ea_g, ea_o, ea_t, 'edge_att')
    check(logit_g, logit_o, logit_t, 'logits')
    check(loss_g, loss_o, loss_t, 'loss')
    named = lambda m: dict(list(m.clf.named_parameters()) + [('ext.' + k, v) for k, v in m.extractor.named_parameters()])
    po, pt, pg = named(go), named(go64), named(gg)
    assert po.keys() == pg.keys()
    for k in po:
        if po[k].grad is None:
            assert pg[k].grad is None or float(pg[k].grad.abs().max()) == 0.0, k
            continue
        check(pg[k].grad, po[k].grad, pt[k].grad, f'grad {k}', rtol=1e-3, atol_scale=2e-4)


@pytest.mark.parametrize('k', [1, 5, 60])
def test_on_device_metrics(G, k):
    """precision@k per graph and delta-KL on the device against the reference's Python loops;
    includes graphs with fewer than k edges, exact ties (attention averaged over reverse edges) and a 1500-edge graph."""
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(40, seed=2)
    g = torch.Generator().manual_seed(k)
    gi_att = torch.rand(b.num_edges, generator=g)
    ref_idx = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda(), b.num_graphs)
    att = ((gi_att + gi_att[ref_idx.rev.cpu().long()]) / 2).view(-1, 1)          # tied in reverse-edge pairs
    labels = ((b.edge_index[0] % 25 >= 20) & (b.edge_index[1] % 25 >= 20)).float()     # motif edges
    exp = O.get_precision_at_k(att, labels, k, b.batch, b.edge_index)
    got = G.get_precision_at_k(att.cuda(), labels.cuda(), k, b.batch.cuda(), b.edge_index.cuda(), b.num_graphs)
    assert got.shape == (b.num_graphs,)
    assert torch.allclose(got.cpu().double(), torch.tensor(exp, dtype=torch.float64), rtol=0, atol=1e-6)
    dk = G.get_delta_kl(labels.cuda(), att.cuda().view(-1))
    assert abs(float(dk) - O.get_delta_kl(labels, att.view(-1))) < 1e-3 * max(1.0, abs(O.get_delta_kl(labels, att.view(-1))))
    # one big graph (several staging chunks)
    n = 400
    src = torch.randint(0, n, (1500,), generator=g)
    dst = torch.randint(0, n, (1500,), generator=g)
    ei, batch = torch.stack([src, dst]), torch.zeros(n, dtype=torch.int64)
    a2, l2 = torch.rand(1500, generator=g), (torch.rand(1500, generator=g) > 0.5).float()
    exp2 = O.get_precision_at_k(a2, l2, k, batch, ei)
    got2 = G.get_precision_at_k(a2.cuda(), l2.cuda(), k, batch.cuda(), ei.cuda(), 1)
    assert torch.allclose(got2.cpu().double(), torch.tensor(exp2, dtype=torch.float64), rtol=0, atol=1e-6)
