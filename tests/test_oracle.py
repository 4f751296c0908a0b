"""CPU suite, part 1: the oracle against the golden vectors generated from the reference's own function bodies
(tests/golden/make_golden.py), the Mutagenicity topology fixture and analytic known answers."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gsat_oracle as O
from tests.conftest import GOLDEN


@pytest.fixture(scope='module')
def gold():
    return torch.load(os.path.join(GOLDEN, 'ref_functions.pt'))


def test_reorder_like_matches_reference_body(gold):
    for name in ('kat4', 'rand_a', 'rand_b'):
        ei, vals, out = (gold[f'reorder_like/{name}/{k}'] for k in ('edge_index', 'values', 'out'))
        t_idx, t_val = O.transpose(ei, vals)
        assert torch.equal(O.reorder_like(t_idx, ei, t_val), out)
        rev = O.reverse_edge_permutation(ei)
        assert torch.equal(vals[rev], out)                       # reorder_like(transpose(ei,v), ei, v) == v[rev]
        idx = O.build_index_oracle(ei, torch.zeros(int(ei.max()) + 1, dtype=torch.long))
        assert torch.equal(idx['rev'].long(), rev) and idx['symmetric'] and not idx['has_dup']


def test_reorder_like_raises_on_directed():
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]])
    assert not O.is_undirected(ei)
    with pytest.raises(ValueError):
        O.reorder_like(torch.stack([ei[1], ei[0]]), ei, torch.arange(3.))


def test_concrete_sample_get_r_lift_gumbel_f1_against_reference(gold):
    lg, u = gold['concrete/logits'], gold['concrete/u']
    assert torch.allclose(O.concrete_sample(lg, 1, True, u), gold['concrete/train'], rtol=0, atol=0)
    assert torch.equal(O.concrete_sample(lg, 1, False), gold['concrete/eval'])
    for e, r in gold['get_r'].tolist():
        assert O.get_r(10, 0.1, int(e), final_r=0.5) == pytest.approx(r, abs=0)
    for e, r in gold['get_r_init07'].tolist():
        assert O.get_r(10, 0.1, int(e), init_r=0.9, final_r=0.7) == pytest.approx(r, abs=0)
    assert torch.equal(O.lift_node_att_to_edge_att(gold['lift/node_att'][:4], gold['lift/edge_index']), gold['lift/out'])
    assert torch.equal(O.gumbel_sigmoid(lg, tau=0.1, noise_u=gold['gumbel/U']), gold['gumbel/out_tau0.1'])
    assert torch.equal(O.f1_sparsity_loss(gold['f1/p'], gold['f1/y']), gold['f1/out'])


def test_get_r_analytic():
    assert O.get_r(10, 0.1, 25, final_r=0.5) == pytest.approx(0.7)
    assert O.get_r(10, 0.1, 40, final_r=0.5) == pytest.approx(0.5)
    assert O.get_r(10, 0.1, 400, final_r=0.5) == 0.5


def test_criterion_mlp_extractor_loss_against_reference(gold):
    crit = O.Criterion(2, False)
    assert torch.equal(crit(gold['criterion/logits'], gold['criterion/y']), gold['criterion/out'])
    mlp = O.MLP([16, 32, 8, 1], dropout=0.5)
    mlp.load_state_dict(gold['mlp/state'])           # identical state_dict keys ('0','4','8')
    mlp.eval()
    assert torch.allclose(mlp(gold['mlp/x'], gold['mlp/batch']), gold['mlp/out_eval'], rtol=0, atol=0)
    ex = O.ExtractorMLP(8, True)
    ex.load_state_dict(gold['extractor/state'])
    ex.eval()
    out = ex(gold['extractor/emb'], gold['extractor/edge_index'], gold['extractor/batch'])
    assert torch.allclose(out, gold['extractor/out_eval'], rtol=0, atol=0)
    g = O.GSAT(None, None, crit, final_r=0.7, decay_interval=10, decay_r=0.1)
    loss, ld = g.__loss__(gold['loss/att'], gold['criterion/logits'], gold['criterion/y'], int(gold['loss/epoch']))
    assert torch.equal(loss, gold['loss/total'])
    assert ld['pred'] == pytest.approx(float(gold['loss/pred']), abs=0)
    assert ld['info'] == pytest.approx(float(gold['loss/info']), abs=0)


def test_mutag_fixture_reverse_is_xor1():
    from dp_gsat_b200.data import load_mutag_fixture
    src, dst, ng = load_mutag_fixture(os.path.join(GOLDEN, 'mutag_slice.npz'))
    ei = torch.from_numpy(np.stack([src, dst]))
    assert O.is_undirected(ei)
    rev = O.reverse_edge_permutation(ei)
    assert torch.equal(rev, torch.arange(ei.shape[1]) ^ 1)
    idx = O.build_index_oracle(ei, torch.from_numpy(ng))
    assert idx['symmetric'] and idx['graph_contiguous'] and not idx['has_dup']
    summ = json.load(open(os.path.join(GOLDEN, 'mutag_full_summary.json')))
    assert summ['rev_is_xor1'] and summ['is_undirected'] and summ['slice_edges'] == ei.shape[1]


def test_analytic_known_answers():
    torch.manual_seed(0)
    # info loss at att == r is ~0 ; concrete sample with u = 0.5 has no noise
    a = torch.full((10, 1), 0.7)
    assert abs(float(O.info_loss(a, 0.7))) < 1e-5
    lg = torch.randn(9, 1)
    assert torch.equal(O.concrete_sample(lg, 1, True, torch.full_like(lg, 0.5)), lg.sigmoid())
    # GINConv with att == 1 equals plain GIN aggregation
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    x = torch.randn(3, 4)
    conv = O.GINConv(torch.nn.Identity())
    assert torch.allclose(conv(x, ei, edge_atten=torch.ones(4, 1)), conv(x, ei))
    # PNA: std of a constant segment is sqrt(1e-5); empty rows give 0 for min / max / mean
    src = torch.ones(4, 2)
    idx = torch.tensor([0, 0, 2, 2])
    assert torch.allclose(O.aggregate_std(src, idx, 4)[0], torch.full((2,), 1e-5).sqrt())
    assert torch.equal(O.scatter_min(src, idx, 4)[1], torch.zeros(2))
    assert torch.equal(O.scatter_max(src, idx, 4)[3], torch.zeros(2))
    assert torch.equal(O.scatter_mean(src, idx, 4)[1], torch.zeros(2))
    # InstanceNorm: per-graph zero mean / unit (biased) variance
    xn = O.InstanceNorm(3)(torch.randn(20, 3), torch.tensor([0] * 8 + [1] * 12))
    assert torch.allclose(xn[:8].mean(0), torch.zeros(3), atol=1e-6)
    assert torch.allclose(xn[8:].var(0, unbiased=False), torch.ones(3), atol=1e-3)


def test_oracle_gradcheck_fp64():
    """fp64 gradcheck of the oracle's differentiable pieces."""
    torch.manual_seed(0)
    ei = torch.tensor([[0, 1, 1, 2, 2, 3, 3, 0], [1, 0, 2, 1, 3, 2, 0, 3]])
    x = torch.randn(4, 4, dtype=torch.double, requires_grad=True)
    att = torch.rand(8, 1, dtype=torch.double, requires_grad=True)
    conv = O.GINConv(torch.nn.Identity()).double()
    assert torch.autograd.gradcheck(lambda a, b: conv(a, ei, edge_atten=b), (x, att))
    lg = torch.randn(8, 1, dtype=torch.double, requires_grad=True)
    u = torch.rand(8, 1, dtype=torch.double).clamp(1e-3, 1 - 1e-3)
    f = lambda l: O.info_loss(O.undirected_average(O.concrete_sample(l, 1, True, u), ei), 0.7)
    assert torch.autograd.gradcheck(f, (lg,))
    inorm = O.InstanceNorm(4)
    seg = torch.tensor([0, 0, 0, 1, 1, 1, 1, 1])
    z = torch.randn(8, 4, dtype=torch.double, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: inorm(t, seg), (z,))


def test_oracle_step_runs_and_is_deterministic():
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(8, seed=0)
    cfg = {'model_name': 'GIN', 'hidden_size': 16, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    outs = []
    for _ in range(2):
        torch.manual_seed(0)
        clf = O.get_model(10, 0, 2, False, cfg)
        ext = O.ExtractorMLP(16, {'learn_edge_att': True, 'extractor_dropout_p': 0.5})
        ms = O.MaskSource(2)
        clf.masks = ms
        ext.masks = ms
        g = O.GSAT(clf, ext, O.Criterion(2, False), learn_edge_att=True, final_r=0.5)
        g.train()
        u = torch.rand(b.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-10, 1 - 1e-10)
        edge_att, loss, ld, logits = g.forward_pass(b, 0, True, noise_u=u)
        loss.backward()
        outs.append((edge_att.detach(), loss.detach(), clf.convs[0].nn[0].weight.grad.clone()))
    assert all(torch.equal(a, c) for a, c in zip(*outs))
    assert outs[0][0].shape == (b.num_edges, 1) and torch.isfinite(outs[0][1])


def test_line_graph_dual_oracle_matches_vectorised_builder_and_known_answer():
    """The dual-graph oracle (reference mutag_dual.py:342-378 restated with its dict loops) against (1) the 4-node
    example of the comment at mutag_dual.py:181-193 and (2) the vectorised numpy builder of data.py on the committed
    Mutagenicity slice."""
    import numpy as np
    from oracle import gsat_oracle as O
    # edges (1,2),(2,1),(1,3),(3,1),(2,4),(4,2),(1,4),(4,1),(2,3),(3,2) -> 0-based
    ei = torch.tensor([[0, 1, 0, 2, 1, 3, 0, 3, 1, 2], [1, 0, 2, 0, 3, 1, 3, 0, 2, 1]])
    batch = torch.zeros(4, dtype=torch.int64)
    dei, db = O.line_graph_dual(ei, batch)
    # groups by first endpoint in order of first appearance: 0:[0,2,6] 1:[1,4,8] 2:[3,9] 3:[5,7]
    exp = [(0, 2), (2, 0), (0, 6), (6, 0), (2, 6), (6, 2), (1, 4), (4, 1), (1, 8), (8, 1), (4, 8), (8, 4), (3, 9), (9, 3),
           (5, 7), (7, 5)]
    assert dei.t().tolist() == [list(p) for p in exp]
    assert db.tolist() == [0] * 10
    deh, dbh = O.line_graph_dual(ei, batch, halve=True)
    assert deh.t().tolist() == [[a // 2, b // 2] for a, b in exp] and dbh.numel() == 5
    # E_d = sum_v d(v) (d(v) - 1)
    deg = torch.bincount(ei[0], minlength=4)
    assert dei.shape[1] == int((deg * (deg - 1)).sum())
    from dp_gsat_b200.data import load_mutag_fixture, line_graph_dual as np_dual
    import os
    from tests.conftest import GOLDEN
    src, dst, ng = load_mutag_fixture(os.path.join(GOLDEN, 'mutag_slice.npz'))
    keep = ng[src] < 40
    ds, dd, dng = np_dual(src[keep], dst[keep], ng)
    dei2, db2 = O.line_graph_dual(torch.from_numpy(np.stack([src[keep], dst[keep]])), torch.from_numpy(ng))
    assert np.array_equal(dei2[0].numpy(), ds) and np.array_equal(dei2[1].numpy(), dd)
    assert np.array_equal(db2.numpy(), dng)


def test_leconv_and_spmotifnet_oracle_known_answers():
    """LEConv restatement (conv_layers.py:69-92): with no weights / attention the layer is the dense identity
    A^T lin1(x) - deg_in * lin2(x) + lin3(x); attention == 0 leaves only lin3(x); fp64 gradcheck of the message;
    SPMotifNet (spmotif_gnn.py) state_dict keys as in the reference class."""
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(3, seed=0)
    N, E = b.num_nodes, b.num_edges
    torch.manual_seed(0)
    conv = O.LEConv(8, 8).double()
    x = torch.randn(N, 8, dtype=torch.float64)
    A = torch.zeros(N, N, dtype=torch.float64)
    A.index_put_((b.edge_index[1], b.edge_index[0]), torch.ones(E, dtype=torch.float64), accumulate=True)
    dense = A @ conv.lin1(x) - A.sum(1, keepdim=True) * conv.lin2(x) + conv.lin3(x)
    assert torch.allclose(conv(x, b.edge_index), dense, atol=1e-12)
    assert torch.allclose(conv(x, b.edge_index, edge_atten=torch.zeros(E, 1, dtype=torch.float64)), conv.lin3(x))
    w = torch.rand(E, 1, dtype=torch.float64)
    assert torch.allclose(conv(x, b.edge_index, edge_weight=w, edge_atten=2 * torch.ones(E, 1, dtype=torch.float64)),
                          conv(x, b.edge_index, edge_weight=2 * w), atol=1e-12)
    xs = x[:, :8].clone().requires_grad_(True)
    ws, ts = w.clone().requires_grad_(True), torch.rand(E, 1, dtype=torch.float64).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a, c, d: conv(a, b.edge_index, edge_weight=c, edge_atten=d), (xs, ws, ts))
    net = O.get_model(4, 1, 3, False, {'model_name': 'SPMotifNet', 'hidden_size': 16, 'n_layers': 2})
    keys = set(net.state_dict().keys())
    assert {'node_emb.weight', 'convs.0.lin1.bias', 'convs.0.lin2.weight', 'convs.1.lin3.weight', 'fc_out.0.weight',
            'fc_out.2.bias', 'conf_mlp.2.weight', 'cq.bias', 'conf_fw.0.0.weight', 'conf_fw.1.bias'} <= keys
    assert 'convs.0.lin2.bias' not in keys
    out = net(torch.rand(N, 4), b.edge_index, b.batch, torch.ones(E, 1))
    assert out.shape == (3, 3)


# ---------------------------------------------------------------------------------------------------------------
# fork-specific classes against the reference's own bodies (tests/golden/ref_fork.pt, made by make_golden_fork.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def fork_gold():
    return torch.load(os.path.join(GOLDEN, 'ref_fork.pt'), weights_only=False)


def test_conv_layers_against_reference_bodies(fork_gold):
    """GINConv / GINEConv / LEConv forward + message (conv_layers.py:14-92) executed from the reference source."""
    fg = fork_gold
    ei, x, att = fg['graph/edge_index'], fg['graph/x'], fg['graph/att']
    H = x.shape[1]
    gin = O.GINConv(O.gin_mlp(H, H))
    gin.load_state_dict(fg['ginconv/state'])
    gin.eval()
    assert torch.allclose(gin(x, ei, edge_atten=att), fg['ginconv/out_att'], rtol=1e-6, atol=1e-6)
    assert torch.allclose(gin(x, ei), fg['ginconv/out_noatt'], rtol=1e-6, atol=1e-6)
    gine = O.GINEConv(O.gin_mlp(H, H), edge_dim=5)
    gine.load_state_dict(fg['gineconv/state'])
    gine.eval()
    assert torch.allclose(gine(x, ei, edge_attr=fg['graph/edge_attr'], edge_atten=att), fg['gineconv/out_att'],
                          rtol=1e-6, atol=1e-6)
    le = O.LEConv(H, H)
    le.load_state_dict(fg['leconv/state'])
    w = fg['graph/edge_weight']
    assert torch.allclose(le(x, ei, edge_weight=w, edge_atten=att), fg['leconv/out_w_att'], rtol=1e-6, atol=1e-6)
    assert torch.allclose(le(x, ei, edge_atten=att), fg['leconv/out_att'], rtol=1e-6, atol=1e-6)
    assert torch.allclose(le(x, ei), fg['leconv/out_plain'], rtol=1e-6, atol=1e-6)


def test_spmotifnet_against_reference_class(fork_gold):
    """SPMotifNet (spmotif_gnn.py:9-87), the whole reference class on the reference LEConv."""
    fg = fork_gold
    net = O.get_model(4, 1, 3, False, {'model_name': 'SPMotifNet', 'n_layers': 2, 'hidden_size': 16})
    assert set(net.state_dict().keys()) == set(fg['spmotif/state'].keys())
    net.load_state_dict(fg['spmotif/state'])
    ei, batch, w, att, x4 = (fg['graph/edge_index'], fg['graph/batch'], fg['graph/edge_weight'], fg['graph/att'],
                             fg['spmotif/x'])
    assert torch.allclose(net(x4, ei, batch, w, edge_atten=att), fg['spmotif/logits'], rtol=1e-6, atol=1e-6)
    assert torch.allclose(net.get_emb(x4, ei, batch, w, edge_atten=att), fg['spmotif/emb'], rtol=1e-6, atol=1e-6)
    gx = net.get_graph_rep(x4, ei, w, batch, att)
    assert torch.allclose(net.get_comb_pred(gx, gx), fg['spmotif/comb_pred'], rtol=1e-6, atol=1e-6)
    assert torch.allclose(net.get_conf_pred(gx), fg['spmotif/conf_pred'], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('epoch', [3, 57])
def test_dual_forward_pass_against_reference_body(fork_gold, epoch):
    """The fork's GSAT.__loss__ + dual_forward_pass (run_gsat.py:121-149, 189-283) executed from the reference source
    with the reference ExtractorMLP / Criterion / MLP / reorder_like, on both sides of the epoch > 50 mix."""
    from types import SimpleNamespace
    fg = fork_gold
    cfg, sc, mc = fg['dual/model_config'], fg['dual/shared_config'], fg['dual/method_config']
    pc, pe = O.get_model(10, 0, 2, False, cfg), O.ExtractorMLP(cfg['hidden_size'], sc, 'primal')
    dc, de = O.get_model(7, 0, 2, False, cfg), O.ExtractorMLP(cfg['hidden_size'], sc, 'dual')
    for name, m in (('primal_clf', pc), ('primal_extractor', pe), ('dual_clf', dc), ('dual_extractor', de)):
        assert set(m.state_dict().keys()) == set(fg[f'dual/{name}/state'].keys()), name
        m.load_state_dict(fg[f'dual/{name}/state'])
        m.eval()
    g = O.DualGSAT(pc, pe, dc, de, O.Criterion(2, False), O.Criterion(2, False), mc, sc, mc, sc)
    data = {side: SimpleNamespace(edge_attr=None, **{k: fg[f'dual/{side}/{k}'] for k in
                                                     ('x', 'edge_index', 'batch', 'y', 'edge_label')})
            for side in ('primal', 'dual')}
    noise = {'primal_u': fg[f'dual/epoch{epoch}/primal_u'], 'dual_U': fg[f'dual/epoch{epoch}/dual_U']}
    edge_att, loss, loss_dict, logits = g.dual_forward_pass(data['primal'], data['dual'], epoch, True, noise)
    assert torch.allclose(edge_att, fg[f'dual/epoch{epoch}/primal_edge_att'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(logits, fg[f'dual/epoch{epoch}/logits'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(loss, fg[f'dual/epoch{epoch}/loss'], rtol=1e-5, atol=1e-6)
    for k, v in fg[f'dual/epoch{epoch}/loss_dict'].items():
        assert abs(loss_dict[k] - v) <= 1e-5 * max(1.0, abs(v)), k


@pytest.mark.parametrize('tag,aggs,scalers,with_ea,post_layers', [
    ('all_identity', ['mean', 'min', 'max', 'std', 'sum', 'var'], ['identity'], True, 1),
    ('scaled', ['mean', 'min', 'max', 'std'], ['identity', 'amplification', 'attenuation', 'linear', 'inverse_linear'],
     True, 1),
    ('noea', ['mean', 'min', 'max', 'std'], ['identity'], False, 2)])
def test_pnaconv_against_reference_class(fork_gold, tag, aggs, scalers, with_ea, post_layers):
    """PNAConvSimple.forward / message / aggregate, all six aggregators and five scalers (conv_layers.py:96-259)."""
    fg = fork_gold
    H = fg['mol/x'].shape[1]
    conv = O.PNAConvSimple((3 if with_ea else 2) * H, H, aggs, scalers, fg['mol/deg'], post_layers=post_layers)
    assert set(conv.state_dict().keys()) == set(fg[f'pnaconv/{tag}/state'].keys())
    conv.load_state_dict(fg[f'pnaconv/{tag}/state'])
    out = conv(fg['mol/x'], fg['mol/edge_index'], fg['mol/edge_feat'] if with_ea else None,
               edge_atten=fg['mol/att'] if with_ea else None)
    assert torch.allclose(out, fg[f'pnaconv/{tag}/out'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('tag', ['pna', 'gin'])
def test_backbones_against_reference_classes(fork_gold, tag):
    """The whole GIN (gin.py:12-81) and PNA (pna.py:12-78) classes executed from the reference source: identical
    state_dict keys (incl. PNA's ``batch_norms.{i}.module.*``), eval and training-mode outputs."""
    fg = fork_gold
    cfg = dict(fg[f'{tag}_model/config'])
    m = O.get_model(9, 3, 2, False, cfg)
    assert set(m.state_dict().keys()) == set(fg[f'{tag}_model/state'].keys())
    m.load_state_dict(fg[f'{tag}_model/state'])
    m.eval()
    args = (fg['mol/x_int'], fg['mol/edge_index'], fg['mol/batch'], fg['mol/edge_attr_int'])
    assert torch.allclose(m(*args, edge_atten=fg['mol/att']), fg[f'{tag}_model/logits'], rtol=1e-5, atol=1e-5)
    assert torch.allclose(m.get_emb(*args, edge_atten=fg['mol/att']), fg[f'{tag}_model/emb'], rtol=1e-5, atol=1e-5)
    m.train()
    m.dropout_p = 0.0
    assert torch.allclose(m(*args, edge_atten=fg['mol/att']), fg[f'{tag}_model/logits_train'], rtol=1e-4, atol=1e-5)


def test_metrics_against_reference_bodies(fork_gold):
    """get_precision_at_k / get_delta_kl (run_gsat.py:783-800) executed from the reference source (tie-free attention)."""
    fg = fork_gold
    att, lab, ei, batch = fg['metrics/att'], fg['metrics/labels'], fg['metrics/edge_index'], fg['metrics/batch']
    for k in (1, 5, 60):
        got = torch.tensor(O.get_precision_at_k(att, lab, k, batch, ei), dtype=torch.float64)
        assert torch.allclose(got, fg[f'metrics/precision_at_{k}'], rtol=0, atol=1e-12), k
    assert abs(O.get_delta_kl(lab, att) - float(fg['metrics/delta_kl'])) < 1e-4 * abs(float(fg['metrics/delta_kl']))


@pytest.mark.parametrize('tag', ['ba2motifs', 'mol'])
def test_line_graph_against_reference_loops(fork_gold, tag):
    """The fork's dual construction (mutag_dual.py:342-378) executed from the reference source: same dual edges, same
    order, as the oracle restatement and the vectorised host builder."""
    from dp_gsat_b200.data import line_graph_dual
    fg = fork_gold
    ei, batch = fg[f'linegraph/{tag}/edge_index'], fg[f'linegraph/{tag}/batch']
    want = fg[f'linegraph/{tag}/dual_edge_index']
    got, got_b = O.line_graph_dual(ei, batch, halve=False)
    assert torch.equal(got, want) and torch.equal(got_b, batch[ei[0]])
    ds, dd, _ = line_graph_dual(ei[0].numpy(), ei[1].numpy(), batch.numpy())
    assert torch.equal(torch.from_numpy(ds), want[0]) and torch.equal(torch.from_numpy(dd), want[1])
    if f'linegraph/{tag}/dual_edge_index_halved' in fg:          # the one-id-per-undirected-edge relabelling, :535-549
        got_h, _ = O.line_graph_dual(ei, batch, halve=True)
        assert torch.equal(got_h, fg[f'linegraph/{tag}/dual_edge_index_halved'])
