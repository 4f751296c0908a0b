"""GPU suite, part 3 (-m gpu): the SURVEY section 8f rows built last -- LEConv / SPMotifNet (row 2, second half), the
fused atom / bond encoders and the device-side batch collate (row 4) -- through the C ABI against the CPU oracle on
the same seeded inputs.  Same bars as tests/test_gpu_parity.py: indices bit-exact, fp32 values rtol 1e-5, whole-step
losses / gradients with the fp64-ground-truth criterion.  The file sorts after the other GPU files on purpose: these
kernels were written after the round's GPU budget was spent and had only been run on the host SIMT emulator
(tests/test_simt_kernels.py) when they were committed."""
import copy

import pytest
import torch

from oracle import gsat_oracle as O
from tests.test_gpu_parity import assert_close, close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def G():
    import dp_gsat_b200 as g
    assert torch.cuda.is_available()
    return g


# ---------------------------------------------------------------------------------------------------------------
# LEConv / SPMotifNet (reference conv_layers.py:69-92, spmotif_gnn.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('H,with_w,with_att', [(16, True, True), (32, True, True), (64, False, True), (128, True, False),
                                               (300, True, True), (4, False, False)])
def test_le_aggregate_fwd_bwd(G, H, with_w, with_att):
    """((a_j - b_i) * edge_weight) * edge_atten summed over incoming edges + lin3 term, forward and all five gradients,
    against the oracle's LEConv lines with the Linears factored out."""
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(40, seed=5)
    g = torch.Generator().manual_seed(H)
    N, E = b.num_nodes, b.num_edges
    a, bb, add = (torch.randn(N, H, generator=g) for _ in range(3))
    w = (torch.rand(E, 1, generator=g) + 0.5) if with_w else None
    att = torch.rand(E, 1, generator=g) if with_att else None
    gout = torch.randn(N, H, generator=g)

    leaf = lambda t, dev=None: None if t is None else (t.clone() if dev is None else t.to(dev)).requires_grad_(True)
    ar, br, dr, wr, tr = leaf(a), leaf(bb), leaf(add), leaf(w), leaf(att)
    m = ar[b.edge_index[0]] - br[b.edge_index[1]]
    if with_w:
        m = m * wr.view(-1, 1)
    if with_att:
        m = m * tr
    ref = O.scatter_sum(m, b.edge_index[1], N) + dr
    (ref * gout).sum().backward()

    gi = G.get_graph_index(b.edge_index.cuda(), b.batch.cuda(), b.num_graphs)
    ad, bd, dd, wd, td = (leaf(t, 'cuda') for t in (a, bb, add, w, att))
    out = G.ops.le_aggregate(ad, bd, wd, td, gi, add=dd)
    (out * gout.cuda()).sum().backward()
    assert_close(out, ref, atol_scale=2e-6, what='le fwd')
    assert_close(ad.grad, ar.grad, atol_scale=2e-6, what='le da')
    assert_close(bd.grad, br.grad, atol_scale=2e-6, what='le db')
    assert_close(dd.grad, dr.grad, atol_scale=2e-6, what='le dadd')
    if with_w:
        assert_close(wd.grad, wr.grad, atol_scale=4e-6, what='le d edge_weight')
    if with_att:
        assert_close(td.grad, tr.grad, atol_scale=4e-6, what='le d att')
    assert torch.equal(out, G.ops.le_aggregate(ad, bd, wd, td, gi, add=dd))        # deterministic
    # without the root term, and isolated nodes / empty edge set
    out2 = G.ops.le_aggregate(ad.detach(), bd.detach(), None, None, gi)
    ref2 = O.scatter_sum(a[b.edge_index[0]] - bb[b.edge_index[1]], b.edge_index[1], N)
    assert_close(out2, ref2, atol_scale=2e-6, what='le fwd plain')
    gi0 = G.get_graph_index(torch.zeros((2, 0), dtype=torch.int64, device='cuda'), b.batch.cuda(), b.num_graphs)
    out0 = G.ops.le_aggregate(ad.detach(), bd.detach(), None, None, gi0, add=dd.detach())
    assert torch.equal(out0.cpu(), add)


def _spmotif_batch(n_graphs=48, seed=3):
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(n_graphs, seed=seed)
    g = torch.Generator().manual_seed(seed)
    b.x = torch.rand(b.num_nodes, 4, generator=g)                       # spmotif.py:56
    b.edge_attr = torch.rand(b.num_edges, 1, generator=g) + 0.5         # spmotif.py:57 uses ones; weights here
    b.y = torch.randint(0, 3, (b.num_graphs,), generator=g)             # 3 motif classes, long labels
    return b


def test_leconv_layer_and_state_dict(G):
    b = _spmotif_batch(12)
    torch.manual_seed(0)
    ref = O.LEConv(32, 32)
    dev = G.LEConv(32, 32).cuda()
    assert set(dev.state_dict().keys()) == set(ref.state_dict().keys()) == \
        {'lin1.weight', 'lin1.bias', 'lin2.weight', 'lin3.weight', 'lin3.bias'}
    dev.load_state_dict(ref.state_dict())
    x = torch.randn(b.num_nodes, 32, generator=torch.Generator().manual_seed(1))
    att = torch.rand(b.num_edges, 1, generator=torch.Generator().manual_seed(2))
    want = ref(x, b.edge_index, edge_weight=b.edge_attr, edge_atten=att)
    got = dev(x.cuda(), b.edge_index.cuda(), edge_weight=b.edge_attr.cuda(), edge_atten=att.cuda())
    assert_close(got, want, rtol=2e-5, atol_scale=4e-6, what='LEConv layer')


@pytest.mark.parametrize('learn_edge_att', [True, False])
def test_gsat_spmotifnet_step_parity(G, learn_edge_att):
    """GSAT + SPMotifNet (LEConv backbone, mean-pool readout, 3-class cross-entropy) whole step against the oracle:
    state_dict keys, edge attention, logits, loss, every parameter gradient."""
    b = _spmotif_batch()
    cfg = {'model_name': 'SPMotifNet', 'hidden_size': 32, 'n_layers': 2}
    shared = {'learn_edge_att': learn_edge_att, 'extractor_dropout_p': 0.5}
    torch.manual_seed(0)
    clf_o, ext_o = O.get_model(4, 1, 3, False, cfg), O.ExtractorMLP(32, shared)
    clf_g, ext_g = G.get_model(4, 1, 3, False, cfg, 'cuda'), G.ExtractorMLP(32, shared).cuda()
    assert isinstance(clf_g, G.SPMotifNet)
    assert set(clf_g.state_dict().keys()) == set(clf_o.state_dict().keys())
    assert {'node_emb.weight', 'convs.0.lin2.weight', 'convs.1.lin3.bias', 'fc_out.2.weight', 'conf_mlp.0.weight',
            'cq.weight', 'conf_fw.0.2.bias', 'conf_fw.1.weight'} <= set(clf_g.state_dict().keys())
    clf_g.load_state_dict(clf_o.state_dict())
    ext_g.load_state_dict(ext_o.state_dict())
    ms = O.MaskSource(2)
    for m in (ext_o, ext_g):
        m.masks = ms
    go = O.GSAT(clf_o, ext_o, O.Criterion(3, False), learn_edge_att=learn_edge_att, final_r=0.7)
    gg = G.GSAT(clf_g, ext_g, G.Criterion(3, False), learn_edge_att=learn_edge_att, final_r=0.7)
    go64 = copy.deepcopy(go).double()
    for m in (go, gg, go64):
        m.train()
    rows = b.num_edges if learn_edge_att else b.num_nodes
    u = torch.rand(rows, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10)
    ea_o, loss_o, _, logit_o = go.forward_pass(b, 3, True, noise_u=u)
    b64 = b.to('cpu')
    b64.x, b64.edge_attr = b64.x.double(), b64.edge_attr.double()
    ea_t, loss_t, _, logit_t = go64.forward_pass(b64, 3, True, noise_u=u.double())
    ea_g, loss_g, _, logit_g = gg.forward_pass(b.to('cuda'), 3, True, noise_u=u.cuda())
    assert logit_g.shape == (b.num_graphs, 3)
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
    # the other reference entry points of the class
    emb = gg.clf.get_emb(b.x.cuda(), b.edge_index.cuda(), b.batch.cuda(), b.edge_attr.cuda())
    emb_o = go.clf.get_emb(b.x, b.edge_index, b.batch, b.edge_attr)
    assert_close(emb, emb_o, rtol=2e-4, atol_scale=2e-5, what='get_emb')
    assert_close(gg.clf.get_pred_from_emb(emb, b.batch.cuda()), go.clf.get_pred_from_emb(emb_o, b.batch), rtol=2e-4,
                 atol_scale=2e-5, what='get_pred_from_emb')
    gx = gg.clf.get_graph_rep(b.x.cuda(), b.edge_index.cuda(), b.edge_attr.cuda(), b.batch.cuda(), None)
    assert_close(gg.clf.get_comb_pred(gx, gx), go.clf.get_comb_pred(gx.cpu(), gx.cpu()), rtol=2e-4, atol_scale=2e-5,
                 what='get_comb_pred')
    assert gg.clf.get_conf_pred(gx).shape == (b.num_graphs, 3)


# ---------------------------------------------------------------------------------------------------------------
# fused atom / bond encoders (reference gin.py:22-25, pna.py:20-23 over ogb 1.3.2)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,H,dims', [(1, 4, [3]), (5000, 80, [119, 4, 12, 12, 10, 6, 6, 2, 2]), (14001, 64, [5, 6, 2]),
                                      (3000, 300, [119, 4, 12, 12, 10, 6, 6, 2, 2]), (2000, 132, [400, 7]),
                                      (300000, 128, [119, 4, 12, 12, 10, 6, 6, 2, 2])])
def test_embedding_sum_fwd_bwd(G, M, H, dims):
    """One gather-sum kernel against the reference's loop (x_embedding = 0; x_embedding += emb_k(x[:, k])): forward
    bit-exact, table gradients within fp32 reordering; tables beyond one shared-memory window, widths that are not a
    multiple of the 64-channel slab, and a row count that uses every row chunk (300 000)."""
    g = torch.Generator().manual_seed(M + H)
    idx = torch.stack([torch.randint(0, d, (M,), generator=g) for d in dims], dim=1).contiguous()
    tables = [torch.randn(d, H, generator=g) for d in dims]
    ref = 0
    for k in range(len(dims)):
        ref = ref + tables[k][idx[:, k]]
    gout = torch.randn(M, H, generator=g)
    # gradient ground truth in fp64: a table row can collect 10^5 rows, where fp32 index_add itself is 1e-5 off
    t64 = [t.double().requires_grad_(True) for t in tables]
    sum((t64[k][idx[:, k]] * gout.double()).sum() for k in range(len(dims))).backward()
    dev_tables = [t.detach().cuda().requires_grad_(True) for t in tables]
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    out = G.ops.embedding_sum(idx.cuda(), dev_tables, flag)
    (out * gout.cuda()).sum().backward()
    assert torch.equal(out.cpu(), ref)
    assert int(flag.item()) == 0
    for k, (td, tr) in enumerate(zip(dev_tables, t64)):
        assert_close(td.grad, tr.grad, rtol=1e-5, atol_scale=2e-6, what=f'd table {k}')
    out2 = G.ops.embedding_sum(idx.cuda(), dev_tables, flag)
    assert torch.equal(out, out2)
    g1 = [t.grad.clone() for t in dev_tables]
    for t in dev_tables:
        t.grad = None
    (out2 * gout.cuda()).sum().backward()
    assert all(torch.equal(a, t.grad) for a, t in zip(g1, dev_tables))           # deterministic backward


def test_embedding_sum_out_of_range_is_clamped_and_flagged(G):
    dims, H = [3, 5], 8
    idx = torch.tensor([[0, 4], [3, 1], [-1, 7]], dtype=torch.int64, device='cuda')
    tables = [torch.arange(3 * H, dtype=torch.float32, device='cuda').view(3, H),
              100 + torch.arange(5 * H, dtype=torch.float32, device='cuda').view(5, H)]
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    out = G.ops.embedding_sum(idx, tables, flag)
    assert int(flag.item()) == 1
    assert torch.equal(out[1], tables[0][2] + tables[1][1]) and torch.equal(out[2], tables[0][0] + tables[1][4])
    with pytest.raises(ValueError):
        G.ops.embedding_sum(idx.int(), tables, flag)


@pytest.mark.parametrize('model_name', ['PNA', 'GIN'])
def test_fused_encoders_inside_the_step(G, model_name):
    """GSAT step on a molhiv-shaped batch with atom / bond encoders: fused encoder kernels (one launch each way) against
    the library-lookup path of the same modules (which tests/test_gpu_parity.py pins to the oracle)."""
    from dp_gsat_b200.data import molhiv_like_batch, in_degree_histogram
    b = molhiv_like_batch(64, seed=13, with_edge_attr=True)
    cfg = {'model_name': model_name, 'hidden_size': 80 if model_name == 'PNA' else 64, 'n_layers': 2, 'dropout_p': 0.0,
           'use_edge_attr': True, 'atom_encoder': True, 'aggregators': ['mean', 'min', 'max', 'std'], 'scalers': False,
           'deg': in_degree_histogram(b)}
    shared = {'learn_edge_att': True, 'extractor_dropout_p': 0.0}
    H = cfg['hidden_size']
    torch.manual_seed(0)
    clf = G.get_model(9, 3, 2, False, cfg, 'cuda')
    ext = G.ExtractorMLP(H, shared).cuda()
    gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.7)
    gsat.train()
    d = b.to('cuda')
    u = torch.rand(b.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10).cuda()
    res = {}
    for fused in (False, True):
        clf.node_encoder.fused = clf.edge_encoder.fused = fused
        gsat.zero_grad(set_to_none=True)
        edge_att, loss, _, logits = gsat.forward_pass(d, 3, True, noise_u=u)
        loss.backward()
        res[fused] = (edge_att.detach().clone(), loss.detach().clone(), logits.detach().clone(),
                      {k: p.grad.clone() for k, p in clf.named_parameters() if p.grad is not None})
    for i, what in ((0, 'edge_att'), (1, 'loss'), (2, 'logits')):        # the encoders' forward bits are identical
        assert_close(res[True][i], res[False][i], rtol=1e-6, atol_scale=1e-7, what=what)
    assert res[True][3].keys() == res[False][3].keys()
    assert any(k.startswith('node_encoder.atom_embedding_list') for k in res[True][3])
    assert any(k.startswith('edge_encoder.bond_embedding_list') for k in res[True][3])
    for k in res[True][3]:
        assert_close(res[True][3][k], res[False][3][k], rtol=1e-4, atol_scale=1e-5, what=f'grad {k}')


# ---------------------------------------------------------------------------------------------------------------
# device-side batch collate (reference get_data_loaders.py:130-145 over PyG collate)
# ---------------------------------------------------------------------------------------------------------------
def _graphs(n_graphs=40, seed=0):
    from dp_gsat_b200.data import molhiv_like_batch
    from dp_gsat_b200.loader import split_batch
    b = molhiv_like_batch(n_graphs, seed=seed, with_edge_attr=True)
    b.edge_label = (torch.arange(b.num_edges) % 3 == 0).float()
    b.node_label = torch.arange(b.num_nodes, dtype=torch.float32)
    return b, split_batch(b)


@pytest.mark.parametrize('ids', [[0], [3, 1, 4, 1, 5, 9, 2, 6], list(range(40)), list(range(39, -1, -1))])
def test_device_collate_bit_exact(G, ids):
    _, graphs = _graphs()
    ds = G.PackedDataset.from_data_list(graphs, device='cuda')
    got = ds.collate(ids)
    want = O.collate_data_list([graphs[i] for i in ids])
    assert got.num_graphs == len(ids)
    for k in ('x', 'edge_index', 'batch', 'y', 'edge_attr', 'edge_label', 'node_label'):
        t = getattr(got, k)
        assert t.device == ds.node_ptr.device and t.dtype == want[k].dtype and torch.equal(t.cpu(), want[k]), k


def test_device_collate_edge_cases_and_loader(G):
    g = torch.Generator().manual_seed(0)
    graphs = []
    for n, e in ((3, 4), (1, 0), (5, 7), (2, 0), (4, 12), (700, 3000)):
        graphs.append(G.Graph(torch.rand(n, 5, generator=g), torch.randint(0, n, (2, e), generator=g),
                              torch.randint(0, 3, (1,), generator=g)))
    ds = G.PackedDataset.from_data_list(graphs, device='cuda')
    for ids in ([1], [1, 3], [0, 1, 2, 3, 4, 5], [5, 3, 4, 1, 5]):
        got, want = ds.collate(ids), O.collate_data_list([graphs[i] for i in ids])
        for k in ('x', 'edge_index', 'batch', 'y'):
            assert torch.equal(getattr(got, k).cpu(), want[k]), (ids, k)
        assert got.edge_attr is None and got.edge_label is None
    batches = list(G.DeviceLoader(ds, ids=[4, 3, 2, 1, 0], batch_size=2))
    assert [bb.num_graphs for bb in batches] == [2, 2, 1]
    assert torch.equal(batches[2].x.cpu(), graphs[0].x)
    with pytest.raises(IndexError):
        ds.collate([6])


def test_step_on_a_device_collated_batch_equals_the_host_batch(G):
    """The loader feeds the path: a GSAT-GIN step on a batch gathered on the device gives the same bits as the step on
    the host-built batch copied over (the reference's route)."""
    from dp_gsat_b200.data import ba2motifs_batch
    from dp_gsat_b200.loader import split_batch
    full = ba2motifs_batch(96, seed=4)
    graphs = split_batch(full)
    ds = G.PackedDataset.from_data_list(graphs, device='cuda')
    ids = list(range(16, 80))
    host = O.collate_data_list([graphs[i] for i in ids])
    hb = G.Batch(host['x'], host['edge_index'], host['batch'], host['y'], None, host['edge_label'], len(ids)).to('cuda')
    db = ds.collate(ids)
    cfg = {'model_name': 'GIN', 'hidden_size': 64, 'n_layers': 2, 'dropout_p': 0.0, 'use_edge_attr': False}
    torch.manual_seed(0)
    clf = G.get_model(full.x.shape[1], 0, 2, False, cfg, 'cuda')
    ext = G.ExtractorMLP(64, {'learn_edge_att': True, 'extractor_dropout_p': 0.0}).cuda()
    gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.7)
    gsat.train()
    u = torch.rand(db.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10).cuda()
    out = []
    for batch in (hb, db):
        edge_att, loss, _, logits = gsat.forward_pass(batch, 0, True, noise_u=u)
        out.append((edge_att.detach().clone(), loss.detach().clone(), logits.detach().clone()))
    for k in ('x', 'edge_index', 'batch', 'y', 'edge_label'):
        assert torch.equal(getattr(hb, k), getattr(db, k)), k
    for a, c, what in zip(out[0], out[1], ('edge_att', 'loss', 'logits')):
        assert_close(a, c, rtol=1e-6, atol_scale=1e-7, what=what)


# ---------------------------------------------------------------------------------------------------------------
# sync BatchNorm over the data-parallel group, exercised on a single-rank group: the collective is the
# identity, so the synced step must reproduce the default step (the 2-rank equality is in tests/test_parallel_gloo.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def single_rank_group():
    import os
    import socket
    import torch.distributed as dist
    if dist.is_initialized():
        yield dist.group.WORLD
        return
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    on_gpu = torch.zeros(1, device='cuda').is_cuda            # False in the emulator dry run (tests/simt/emulate.py)
    dist.init_process_group('nccl' if on_gpu else 'gloo', rank=0, world_size=1)
    yield dist.group.WORLD
    dist.destroy_process_group()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_sync_batchnorm_on_a_single_rank_group_equals_the_default(G, single_rank_group, precision):
    from dp_gsat_b200.data import ba2motifs_batch
    from dp_gsat_b200.parallel import enable_sync_batchnorm
    b = ba2motifs_batch(64, seed=6)
    b.x = torch.rand(b.x.shape, generator=torch.Generator().manual_seed(3))
    cfg = {'model_name': 'GIN', 'hidden_size': 64, 'n_layers': 2, 'dropout_p': 0.0, 'use_edge_attr': False}
    torch.manual_seed(0)
    clf = G.get_model(b.x.shape[1], 0, 2, False, cfg, 'cuda')
    ext = G.ExtractorMLP(64, {'learn_edge_att': True, 'extractor_dropout_p': 0.0}).cuda()
    clf.precision = ext.precision = precision
    gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.7)
    gsat.train()
    d = b.to('cuda')
    u = torch.rand(b.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10).cuda()
    state = {k: v.clone() for k, v in clf.state_dict().items()}
    res = {}
    for sync in (False, True):
        clf.load_state_dict(state)
        assert enable_sync_batchnorm(clf, single_rank_group, enabled=sync) == 2
        gsat.zero_grad(set_to_none=True)
        _, loss, _, logits = gsat.forward_pass(d, 0, True, noise_u=u)
        loss.backward()
        res[sync] = (loss.detach().clone(), logits.detach().clone(),
                     {k: p.grad.clone() for k, p in clf.named_parameters() if p.grad is not None},
                     clf.convs[0].nn[1].running_var.clone())
    enable_sync_batchnorm(clf, enabled=False)
    rtol, scale = (2e-4, 2e-5) if precision == 'fp32' else (2e-2, 2e-3)
    assert_close(res[True][0], res[False][0], rtol=rtol, atol_scale=scale, what='loss')
    assert_close(res[True][1], res[False][1], rtol=rtol, atol_scale=scale, what='logits')
    assert_close(res[True][3], res[False][3], rtol=rtol, atol_scale=scale, what='running_var')
    for k in res[True][2]:
        assert_close(res[True][2][k], res[False][2][k], rtol=10 * rtol, atol_scale=10 * scale, what=f'grad {k}')


# ---------------------------------------------------------------------------------------------------------------
# the fork's two-model step: GSAT.dual_forward_pass (reference src/run_gsat.py:121-149, 189-283)
# ---------------------------------------------------------------------------------------------------------------
def _primal_dual_pair(n_graphs=24, seed=8):
    """A primal batch (BA-2Motifs-shaped, random features, motif edge labels) and its line-graph dual: one dual node
    per directed primal edge (mutag_dual.py:342-378), dual features random, dual labels = primal labels."""
    import numpy as np
    import dp_gsat_b200 as G
    from dp_gsat_b200.data import ba2motifs_batch, line_graph_dual, graph_contiguous_relabel
    p = ba2motifs_batch(n_graphs, seed=seed)
    g = torch.Generator().manual_seed(seed)
    p.x = torch.rand(p.num_nodes, 10, generator=g)
    src, dst = p.edge_index[0].numpy(), p.edge_index[1].numpy()
    dsrc, ddst, dng = line_graph_dual(src, dst, p.batch.numpy())
    dsrc, ddst = graph_contiguous_relabel(dsrc, ddst, dng)
    d = G.Batch(torch.rand(p.num_edges, 7, generator=g), torch.from_numpy(np.stack([dsrc, ddst])),
                torch.from_numpy(dng), p.y.clone(), None, torch.zeros(dsrc.shape[0]), n_graphs)
    return p, d


@pytest.mark.parametrize('primal_learn_edge_att,epoch', [(True, 3), (False, 3), (True, 60), (False, 60)])
def test_dual_forward_pass_parity(G, primal_learn_edge_att, epoch):
    """Whole fork step against the restated reference method: primal / dual attention paths (reverse average or lift),
    gumbel_sigmoid dual attention, f1 loss, per-edge prior r, the epoch > 50 mix, loss and every gradient of all four
    modules."""
    p, d = _primal_dual_pair()
    cfg = {'model_name': 'GIN', 'hidden_size': 32, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    sc_p = {'learn_edge_att': primal_learn_edge_att, 'extractor_dropout_p': 0.5, 'precision_k': 5}
    sc_d = {'learn_edge_att': False, 'extractor_dropout_p': 0.5, 'precision_k': 5}
    mc = {'method_name': 'GSAT', 'pred_loss_coef': 1, 'info_loss_coef': 1, 'epochs': 100, 'decay_interval': 10,
          'decay_r': 0.1, 'final_r': 0.5, 'init_r': 0.9}
    torch.manual_seed(0)
    o_mods = (O.get_model(10, 0, 2, False, cfg), O.ExtractorMLP(32, sc_p, 'primal'),
              O.get_model(7, 0, 2, False, cfg), O.ExtractorMLP(32, sc_d, 'dual'))
    g_mods = (G.get_model(10, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(32, sc_p, 'primal').cuda(),
              G.get_model(7, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(32, sc_d, 'dual').cuda())
    for mo, mg in zip(o_mods, g_mods):
        assert set(mg.state_dict().keys()) == set(mo.state_dict().keys())
        mg.load_state_dict(mo.state_dict())
    ms = O.MaskSource(2)
    for m in o_mods + g_mods:
        m.masks = ms
    go = O.DualGSAT(o_mods[0], o_mods[1], o_mods[2], o_mods[3], O.Criterion(2, False), O.Criterion(2, False), mc, sc_p,
                    mc, sc_d)
    gg = G.DualGSAT(g_mods[0], g_mods[1], g_mods[2], g_mods[3], 2, False, 2, False, mc, sc_p, mc, sc_d)
    go64 = copy.deepcopy(go).double()
    for m in (go, gg, go64):
        m.train()
    gen = torch.Generator().manual_seed(1)
    rows_p = p.num_edges if primal_learn_edge_att else p.num_nodes
    noise = {'primal_u': torch.rand(rows_p, 1, generator=gen).clamp(1e-10, 1 - 1e-10),
             'dual_U': torch.rand(d.num_nodes, 1, generator=gen)}
    cast = lambda n, f: {k: f(v) for k, v in n.items()}
    ea_o, loss_o, ld_o, logit_o = go.dual_forward_pass(p, d, epoch, True, noise)
    p64, d64 = p.to('cpu'), d.to('cpu')
    p64.x, d64.x = p64.x.double(), d64.x.double()
    ea_t, loss_t, _, logit_t = go64.dual_forward_pass(p64, d64, epoch, True, cast(noise, lambda v: v.double()))
    ea_g, loss_g, ld_g, logit_g = gg.dual_forward_pass(p.to('cuda'), d.to('cuda'), epoch, True,
                                                       cast(noise, lambda v: v.cuda()))
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
    check(ea_g, ea_o, ea_t, 'primal_edge_att')
    check(logit_g, logit_o, logit_t, 'primal logits')
    check(loss_g, loss_o, loss_t, 'loss')
    assert set(ld_g.keys()) == set(ld_o.keys()) == {'loss', 'pred', 'info'}
    for k in ld_o:
        assert abs(ld_g[k] - ld_o[k]) <= 2e-4 * max(1.0, abs(ld_o[k])), k
    names = ('primal_clf', 'primal_extractor', 'dual_clf', 'dual_extractor')
    for name in names:
        po = dict(getattr(go, name).named_parameters())
        pt = dict(getattr(go64, name).named_parameters())
        pg = dict(getattr(gg, name).named_parameters())
        assert po.keys() == pg.keys()
        for k in po:
            if po[k].grad is None:
                assert pg[k].grad is None or float(pg[k].grad.abs().max()) == 0.0, (name, k)
                continue
            check(pg[k].grad, po[k].grad, pt[k].grad, f'grad {name}.{k}', rtol=1e-3, atol_scale=2e-4)


def test_dual_train_and_eval_one_batch(G):
    """dual_train_one_batch / dual_eval_one_batch (run_gsat.py:610-637) and the 28-argument reference constructor."""
    p, d = _primal_dual_pair(12, seed=2)
    cfg = {'model_name': 'GIN', 'hidden_size': 32, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    sc = {'learn_edge_att': False, 'extractor_dropout_p': 0.5, 'precision_k': 5, 'num_viz_samples': 0, 'viz_interval': 10,
          'viz_norm_att': True}
    mc = {'method_name': 'GSAT', 'pred_loss_coef': 1, 'info_loss_coef': 1, 'epochs': 100, 'decay_interval': 10,
          'decay_r': 0.1, 'final_r': 0.5, 'lr': 1e-3}
    torch.manual_seed(0)
    pc, pe = G.get_model(10, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(32, sc, 'primal').cuda()
    dc, de = G.get_model(7, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(32, sc, 'dual').cuda()
    po = torch.optim.Adam(list(pe.parameters()) + list(pc.parameters()), lr=1e-2)
    do = torch.optim.Adam(list(de.parameters()) + list(dc.parameters()), lr=1e-2)
    gsat = G.DualGSAT.from_reference_args(pc, pe, po, None, None, 'cuda', None, 'mutag', 2, False, 0, mc, sc, cfg,
                                          dc, de, do, None, None, 'cuda', None, 'mutag_dual', 2, False, 0, mc, sc, cfg)
    assert gsat.primal_dataset_name == 'mutag' and gsat.dual_final_r == 0.5 and gsat.primal_learn_edge_att is False
    pd_, dd_ = p.to('cuda'), d.to('cuda')
    w0 = pc.convs[0].nn[0].weight.detach().clone()
    losses = []
    for epoch in range(4):
        att, loss_dict, logits = gsat.dual_train_one_batch(pd_, dd_, epoch)
        assert att.shape == (p.num_edges,) and logits.shape == (p.num_graphs, 1)
        losses.append(loss_dict['loss'])
    assert not torch.equal(w0, pc.convs[0].nn[0].weight.detach())
    assert all(l == l for l in losses)                                          # finite
    att, loss_dict, logits = gsat.dual_eval_one_batch(pd_, dd_, 4)
    att2, _, logits2 = gsat.dual_eval_one_batch(pd_, dd_, 4)
    assert torch.equal(att, att2) and torch.equal(logits, logits2)              # eval: no sampling noise, no dropout
    assert float(att.min()) >= 0.0 and float(att.max()) <= 1.0


# ---------------------------------------------------------------------------------------------------------------
# product directly against the outputs of the reference's OWN class bodies (tests/golden/ref_fork.pt, generated by
# tests/golden/make_golden_fork.py from /root/reference): no oracle in between -- state_dict keys, shapes, outputs
# ---------------------------------------------------------------------------------------------------------------
_FORK_GOLD = {}


def _fork_gold():
    import os
    from tests.conftest import GOLDEN
    if not _FORK_GOLD:
        raw = torch.load(os.path.join(GOLDEN, 'ref_fork.pt'), weights_only=False)
        _FORK_GOLD.update({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in raw.items()})
    return _FORK_GOLD


def _cfg_on_host(cfg):
    return {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in cfg.items()}


def _load(module, state):
    assert set(module.state_dict().keys()) == set(state.keys())
    for k, v in module.state_dict().items():
        assert v.shape == state[k].shape, k
    module.load_state_dict(state)
    return module.cuda()


def test_product_layers_reproduce_the_reference_layers(G):
    fg = _fork_gold()
    ei, x, att = fg['graph/edge_index'], fg['graph/x'], fg['graph/att']
    H = x.shape[1]
    tol = dict(rtol=1e-5, atol=1e-5)
    gin = _load(G.GINConv(G.GIN.MLP(H, H)), fg['ginconv/state']).eval()
    assert torch.allclose(gin(x, ei, edge_atten=att), fg['ginconv/out_att'], **tol)
    assert torch.allclose(gin(x, ei), fg['ginconv/out_noatt'], **tol)
    gine = _load(G.GINEConv(G.GIN.MLP(H, H), edge_dim=5), fg['gineconv/state']).eval()
    assert torch.allclose(gine(x, ei, edge_attr=fg['graph/edge_attr'], edge_atten=att), fg['gineconv/out_att'], **tol)
    le = _load(G.LEConv(H, H), fg['leconv/state'])
    assert torch.allclose(le(x, ei, edge_weight=fg['graph/edge_weight'], edge_atten=att), fg['leconv/out_w_att'], **tol)
    assert torch.allclose(le(x, ei, edge_atten=att), fg['leconv/out_att'], **tol)
    assert torch.allclose(le(x, ei), fg['leconv/out_plain'], **tol)
    Hm = fg['mol/x'].shape[1]
    for tag, aggs, scalers, with_ea, post in (
            ('all_identity', ['mean', 'min', 'max', 'std', 'sum', 'var'], ['identity'], True, 1),
            ('scaled', ['mean', 'min', 'max', 'std'],
             ['identity', 'amplification', 'attenuation', 'linear', 'inverse_linear'], True, 1),
            ('noea', ['mean', 'min', 'max', 'std'], ['identity'], False, 2)):
        conv = _load(G.PNAConvSimple((3 if with_ea else 2) * Hm, Hm, aggs, scalers, fg['mol/deg'].cpu(), post_layers=post),
                     fg[f'pnaconv/{tag}/state'])
        out = conv(fg['mol/x'], fg['mol/edge_index'], fg['mol/edge_feat'] if with_ea else None,
                   edge_atten=fg['mol/att'] if with_ea else None)
        assert torch.allclose(out, fg[f'pnaconv/{tag}/out'], rtol=1e-4, atol=1e-4), tag


@pytest.mark.parametrize('tag', ['pna', 'gin', 'spmotif'])
def test_product_backbones_reproduce_the_reference_classes(G, tag):
    fg = _fork_gold()
    if tag == 'spmotif':
        net = _load(G.get_model(4, 1, 3, False, {'model_name': 'SPMotifNet', 'n_layers': 2, 'hidden_size': 16}, 'cuda'),
                    fg['spmotif/state'])
        ei, batch, w, att, x4 = (fg['graph/edge_index'], fg['graph/batch'], fg['graph/edge_weight'], fg['graph/att'],
                                 fg['spmotif/x'])
        assert torch.allclose(net(x4, ei, batch, w, edge_atten=att), fg['spmotif/logits'], rtol=1e-5, atol=1e-5)
        assert torch.allclose(net.get_emb(x4, ei, batch, w, edge_atten=att), fg['spmotif/emb'], rtol=1e-5, atol=1e-5)
        gx = net.get_graph_rep(x4, ei, w, batch, att)
        assert torch.allclose(net.get_comb_pred(gx, gx), fg['spmotif/comb_pred'], rtol=1e-5, atol=1e-5)
        return
    m = _load(G.get_model(9, 3, 2, False, _cfg_on_host(fg[f'{tag}_model/config']), 'cuda'), fg[f'{tag}_model/state']).eval()
    args = (fg['mol/x_int'], fg['mol/edge_index'], fg['mol/batch'], fg['mol/edge_attr_int'])
    assert torch.allclose(m(*args, edge_atten=fg['mol/att']), fg[f'{tag}_model/logits'], rtol=1e-4, atol=1e-5)
    assert torch.allclose(m.get_emb(*args, edge_atten=fg['mol/att']), fg[f'{tag}_model/emb'], rtol=1e-4, atol=1e-5)
    for fused in (False, True):                     # library lookups and the one-kernel encoders
        m.node_encoder.fused = m.edge_encoder.fused = fused
        m.train()
        m.dropout_p = 0.0
        assert torch.allclose(m(*args, edge_atten=fg['mol/att']), fg[f'{tag}_model/logits_train'], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('epoch', [3, 57])
def test_product_dual_forward_pass_reproduces_the_reference_body(G, epoch):
    fg = _fork_gold()
    cfg, sc, mc = fg['dual/model_config'], fg['dual/shared_config'], fg['dual/method_config']
    pc, pe = G.get_model(10, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(cfg['hidden_size'], sc, 'primal').cuda()
    dc, de = G.get_model(7, 0, 2, False, cfg, 'cuda'), G.ExtractorMLP(cfg['hidden_size'], sc, 'dual').cuda()
    for name, m in (('primal_clf', pc), ('primal_extractor', pe), ('dual_clf', dc), ('dual_extractor', de)):
        _load(m, fg[f'dual/{name}/state']).eval()
    gsat = G.DualGSAT(pc, pe, dc, de, 2, False, 2, False, mc, sc, mc, sc)
    data = {side: G.Batch(fg[f'dual/{side}/x'], fg[f'dual/{side}/edge_index'], fg[f'dual/{side}/batch'],
                          fg[f'dual/{side}/y'], None, fg[f'dual/{side}/edge_label'],
                          int(fg[f'dual/{side}/y'].shape[0])) for side in ('primal', 'dual')}
    noise = {'primal_u': fg[f'dual/epoch{epoch}/primal_u'], 'dual_U': fg[f'dual/epoch{epoch}/dual_U']}
    edge_att, loss, loss_dict, logits = gsat.dual_forward_pass(data['primal'], data['dual'], epoch, True, noise)
    assert torch.allclose(edge_att, fg[f'dual/epoch{epoch}/primal_edge_att'], rtol=1e-4, atol=1e-5)
    assert torch.allclose(logits, fg[f'dual/epoch{epoch}/logits'], rtol=1e-4, atol=1e-5)
    assert torch.allclose(loss, fg[f'dual/epoch{epoch}/loss'], rtol=1e-4, atol=1e-5)
    for k, v in fg[f'dual/epoch{epoch}/loss_dict'].items():
        assert abs(loss_dict[k] - v) <= 1e-4 * max(1.0, abs(v)), k


def test_product_metrics_reproduce_the_reference_bodies(G):
    """On-device precision@k / delta-KL against the outputs of the reference's own get_precision_at_k / get_delta_kl
    (run_gsat.py:783-800, tie-free attention)."""
    fg = _fork_gold()
    att, lab, ei, batch = fg['metrics/att'], fg['metrics/labels'], fg['metrics/edge_index'], fg['metrics/batch']
    ng = int(batch.max().item()) + 1
    for k in (1, 5, 60):
        got = G.get_precision_at_k(att.view(-1, 1), lab, k, batch, ei, ng)
        assert torch.allclose(got.double().cpu(), fg[f'metrics/precision_at_{k}'].cpu(), rtol=0, atol=1e-6), k
    dk = float(G.get_delta_kl(lab, att))
    assert abs(dk - float(fg['metrics/delta_kl'])) < 1e-3 * abs(float(fg['metrics/delta_kl']))


@pytest.mark.parametrize('tag', ['ba2motifs', 'mol'])
def test_product_line_graph_reproduces_the_reference_loops(G, tag):
    """GPU line-graph builder against the dual edges the reference's own loops emit (mutag_dual.py:342-378), bit-exact."""
    fg = _fork_gold()
    ei, batch = fg[f'linegraph/{tag}/edge_index'], fg[f'linegraph/{tag}/batch']
    got_ei, got_b = G.line_graph_dual(ei, batch, halve=False)
    assert torch.equal(got_ei.cpu(), fg[f'linegraph/{tag}/dual_edge_index'].cpu())
    assert torch.equal(got_b.cpu(), batch[ei[0]].cpu())
    if f'linegraph/{tag}/dual_edge_index_halved' in fg:          # the relabelling of mutag_dual.py:535-549
        got_h, _ = G.line_graph_dual(ei, batch, halve=True)
        assert torch.equal(got_h.cpu(), fg[f'linegraph/{tag}/dual_edge_index_halved'].cpu())


def test_dense_dual_reproduces_the_reference_loops(G):
    """line_graph_dual_dense against the dual adjacency the reference's own matrix loops build
    (ba_2motifs_dual.py:26-62 + dense_to_sparse), bit-exact, incl. the motif node labels carried over by und_id."""
    fg = _fork_gold()
    ei, batch = fg['densedual/edge_index'], fg['densedual/batch']
    dual_ei, dual_batch, und = G.line_graph_dual_dense(ei, batch)
    assert torch.equal(dual_ei.cpu(), fg['densedual/dual_edge_index'].cpu())
    src = ei[0]
    assert torch.equal(dual_batch.cpu(), torch.zeros_like(dual_batch).cpu().scatter_(0, und.cpu(), batch[src].cpu()))
    # the reference labels a dual node 1 when both endpoints are motif nodes (>= 20 inside the 25-node graph): :46-47
    motif = ((ei[0] % 25 >= 20) & (ei[1] % 25 >= 20)).float()
    lab = torch.zeros(dual_batch.numel(), device=und.device).scatter_(0, und, motif)
    assert torch.equal(lab.cpu(), fg['densedual/dual_node_label'].cpu())
    dual_x = G.dense_dual_node_features(fg['densedual/x'], ei, und)                 # cat(x[u], x[v]), u < v  (:48)
    assert torch.equal(dual_x.cpu(), fg['densedual/dual_x'].cpu())
    with pytest.raises(ValueError):
        G.line_graph_dual_dense(ei[:, ei[0] < ei[1]].contiguous(), batch)          # not symmetric
