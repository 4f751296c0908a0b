"""Generate tests/golden/ref_fork.pt: outputs of the REFERENCE'S OWN class bodies for the fork-specific parts of the
path, extracted with ``ast`` from /root/reference and executed unmodified.  Run ONCE in the build container, where
/root/reference exists:   python tests/golden/make_golden_fork.py

Pinned here (all first-party code of mihikamd/DP-GSAT):

* ``GINConv`` / ``GINEConv`` / ``LEConv`` of src/models/conv_layers.py:14-92 -- their ``forward`` and ``message``
  bodies, on stub base classes that supply what torch_geometric's bases supply (``nn``, ``eps``, ``lin``, ``lin1..3``
  and a ``propagate`` that gathers x_j = x[edge_index[0]], x_i = x[edge_index[1]], calls the reference ``message`` and
  scatter-adds into edge_index[1] -- SURVEY App. A.1; that third-party part stays UNPINNED);
* ``SPMotifNet`` (src/models/spmotif_gnn.py:9-87), whole class, on the reference LEConv above;
* ``PNAConvSimple`` with its aggregators and scalers (src/models/conv_layers.py:96-259; torch_scatter's ``scatter`` and
  PyG's ``degree`` bound to the restatements) and the whole ``GIN`` / ``PNA`` backbone classes (src/models/gin.py,
  pna.py) on those layers, with ogb's encoders and PyG's BatchNorm / pooling bound to the restatements;
* the fork's ``GSAT`` class (src/run_gsat.py:33-283): ``__init__``, ``__loss__``, ``dual_forward_pass`` and the helpers
  they call, with the reference's ``ExtractorMLP`` (:886-927), ``Criterion`` / ``MLP`` (src/utils/get_model.py) and
  ``reorder_like`` (src/utils/utils.py); the GNN backbones inside are the oracle's GIN (PyG-dependent, unpinned).
  Run with every module in eval mode (no dropout draws, BatchNorm on its running statistics) and ``training=True``, so
  the only random draws are the sampler's uniform and gumbel_sigmoid's ``rand_like``, which are replayed from the same
  seed and stored; epochs 3 and 57 (not multiples of 10: the plotting block of :394-426 stays off) cover both sides of
  the ``epoch > 50`` mix.  ``primal_learn_edge_att`` is False in both (with True the reference body raises NameError
  on ``old_primal_edge_att``);
* the line-graph construction loops of src/datasets/mutag_dual.py:342-378 (``group_by_first`` / ``add_pairs_from_group``),
  executed as they stand on a primal edge list;
* the dense-matrix dual of src/datasets/ba_2motifs_dual.py:26-62 (edge numbering by adjacency scan, shared-node adjacency),
  executed as it stands on BA-2Motifs-shaped adjacency matrices;
* the trainer's per-batch explanation metrics ``get_precision_at_k`` / ``get_delta_kl`` (src/run_gsat.py:783-800).
"""
import ast
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import gsat_oracle as O  # noqa: E402
from make_golden import _extract, _compile  # noqa: E402

REF = '/root/reference'


def _strip_annotations(node):
    """The reference annotates with torch_geometric.typing names (Adj, OptTensor ...): drop annotations, keep bodies."""
    for n in ast.walk(node):
        if isinstance(n, ast.FunctionDef):
            n.returns = None
            for a in n.args.args + n.args.kwonlyargs:
                a.annotation = None
        if isinstance(n, ast.AnnAssign):      # `x: OptPairTensor = (x, x)` -> `x = (x, x)`
            pass
    class T(ast.NodeTransformer):
        def visit_AnnAssign(self, n):
            return ast.copy_location(ast.Assign(targets=[n.target], value=n.value), n) if n.value is not None else None
    return ast.fix_missing_locations(T().visit(node))


class _Propagate(nn.Module):
    """What torch_geometric.nn.MessagePassing.propagate does for these layers (flow source_to_target, aggr 'add')."""

    def propagate(self, edge_index, size=None, **kw):
        args = {}
        for k, v in kw.items():
            if k in ('x', 'a', 'b'):
                xs = v if isinstance(v, tuple) else (v, v)
                args[k + '_j'] = xs[0].index_select(0, edge_index[0])
                args[k + '_i'] = xs[1].index_select(0, edge_index[1])
            else:
                args[k] = v
        import inspect
        want = inspect.signature(self.message).parameters
        msg = self.message(**{k: v for k, v in args.items() if k in want})
        n = (kw['x'] if 'x' in kw else kw['b'])
        n = (n[1] if isinstance(n, tuple) else n).shape[0]
        return O.scatter_sum(msg, edge_index[1], n)


class BaseGINConv(_Propagate):
    def __init__(self, nn_, eps=0.0):
        super().__init__()
        self.nn = nn_
        self.register_buffer('eps', torch.tensor([eps]))


class BaseGINEConv(_Propagate):
    def __init__(self, nn_, eps=0.0, edge_dim=None):
        super().__init__()
        self.nn = nn_
        self.register_buffer('eps', torch.tensor([eps]))
        self.lin = nn.Linear(edge_dim, nn_[0].in_features) if edge_dim is not None else None


class BaseLEConv(_Propagate):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.lin1 = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin2 = nn.Linear(in_channels, out_channels, bias=False)
        self.lin3 = nn.Linear(in_channels, out_channels, bias=bias)


def main():
    gold = {}
    quiet = lambda *a, **k: None
    g = torch.Generator().manual_seed(0)

    # ---- conv layers ------------------------------------------------------------------------------------------
    convs = _extract(f'{REF}/src/models/conv_layers.py', ['GINConv', 'GINEConv', 'LEConv'])
    ns = _compile([_strip_annotations(n) for n in convs.values()],
                  {'torch': torch, 'Tensor': torch.Tensor, 'BaseGINConv': BaseGINConv, 'BaseGINEConv': BaseGINEConv,
                   'BaseLEConv': BaseLEConv, 'print': quiet})
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(4, seed=1)
    N, E, H = b.num_nodes, b.num_edges, 8
    x = torch.randn(N, H, generator=g)
    att = torch.rand(E, 1, generator=g)
    ea = torch.randn(E, 5, generator=g)
    w = torch.rand(E, 1, generator=g) + 0.5
    gold['graph/edge_index'], gold['graph/batch'] = b.edge_index, b.batch
    gold['graph/x'], gold['graph/att'], gold['graph/edge_attr'], gold['graph/edge_weight'] = x, att, ea, w
    torch.manual_seed(1)
    gin = ns['GINConv'](O.gin_mlp(H, H))
    gin.eval()
    gold['ginconv/state'] = {k: v.clone() for k, v in gin.state_dict().items()}
    gold['ginconv/out_att'] = gin(x, b.edge_index, edge_atten=att).detach()
    gold['ginconv/out_noatt'] = gin(x, b.edge_index).detach()
    gine = ns['GINEConv'](O.gin_mlp(H, H), edge_dim=5)
    gine.eval()
    gold['gineconv/state'] = {k: v.clone() for k, v in gine.state_dict().items()}
    gold['gineconv/out_att'] = gine(x, b.edge_index, edge_attr=ea, edge_atten=att).detach()
    le = ns['LEConv'](H, H)
    gold['leconv/state'] = {k: v.clone() for k, v in le.state_dict().items()}
    gold['leconv/out_w_att'] = le(x, b.edge_index, edge_weight=w, edge_atten=att).detach()
    gold['leconv/out_att'] = le(x, b.edge_index, edge_atten=att).detach()
    gold['leconv/out_plain'] = le(x, b.edge_index).detach()

    # ---- SPMotifNet -------------------------------------------------------------------------------------------
    sp = _extract(f'{REF}/src/models/spmotif_gnn.py', ['SPMotifNet'])
    ns_sp = _compile(sp.values(), {'torch': torch, 'Linear': nn.Linear, 'ReLU': nn.ReLU, 'ModuleList': nn.ModuleList,
                                   'global_mean_pool': O.global_mean_pool, 'LEConv': ns['LEConv']})
    torch.manual_seed(2)
    net = ns_sp['SPMotifNet'](4, 1, 3, False, {'n_layers': 2, 'hidden_size': 16})
    x4 = torch.rand(N, 4, generator=g)
    gold['spmotif/state'] = {k: v.clone() for k, v in net.state_dict().items()}
    gold['spmotif/x'] = x4
    gold['spmotif/logits'] = net(x4, b.edge_index, b.batch, w, edge_atten=att).detach()
    gold['spmotif/emb'] = net.get_emb(x4, b.edge_index, b.batch, w, edge_atten=att).detach()
    gx = net.get_graph_rep(x4, b.edge_index, w, b.batch, att).detach()
    gold['spmotif/comb_pred'] = net.get_comb_pred(gx, gx).detach()
    gold['spmotif/conf_pred'] = net.get_conf_pred(gx).detach()

    # ---- PNAConvSimple with its aggregators / scalers, and the GIN / PNA backbones ------------------------------
    tree = ast.parse(open(f'{REF}/src/models/conv_layers.py').read())
    pna_nodes = [n for n in tree.body
                 if (isinstance(n, (ast.FunctionDef, ast.ClassDef)) and (n.name.startswith(('aggregate_', 'scale_'))
                                                                         or n.name == 'PNAConvSimple'))
                 or (isinstance(n, ast.Assign) and getattr(n.targets[0], 'id', '') in ('AGGREGATORS', 'SCALERS'))]

    class MessagePassing(_Propagate):
        def __init__(self, aggr=None, node_dim=0, **kwargs):
            super().__init__()

        def propagate(self, edge_index, size=None, **kw):
            x = kw['x']
            msg = self.message(x_i=x.index_select(0, edge_index[1]), x_j=x.index_select(0, edge_index[0]),
                               edge_attr=kw.get('edge_attr'), edge_atten=kw.get('edge_atten'))
            return self.aggregate(msg, edge_index[1], dim_size=x.shape[0])

    def scatter(src, index, dim, out, dim_size, reduce):          # torch_scatter.scatter (restated, unpinned)
        assert dim == 0 and out is None
        return {'sum': O.scatter_sum, 'mean': O.scatter_mean, 'min': O.scatter_min, 'max': O.scatter_max}[reduce](
            src, index, dim_size)

    def reset(module):
        for m in module.modules():
            if hasattr(m, 'reset_parameters') and m is not module:
                m.reset_parameters()
    ns_pna = _compile([_strip_annotations(n) for n in pna_nodes],
                      {'torch': torch, 'MessagePassing': MessagePassing, 'scatter': scatter, 'degree': O.degree,
                       'reset': reset, 'Linear': nn.Linear, 'ReLU': nn.ReLU, 'Sequential': nn.Sequential,
                       'print': quiet})
    from dp_gsat_b200.data import molhiv_like_batch, in_degree_histogram
    mb = molhiv_like_batch(6, seed=4, with_edge_attr=True)
    deg = in_degree_histogram(mb)
    Hp = 8
    xm = torch.randn(mb.num_nodes, Hp, generator=g)
    eam = torch.randn(mb.num_edges, Hp, generator=g)
    attm = torch.rand(mb.num_edges, 1, generator=g)
    gold['mol/edge_index'], gold['mol/batch'], gold['mol/x_int'], gold['mol/edge_attr_int'] = (
        mb.edge_index, mb.batch, mb.x, mb.edge_attr)
    gold['mol/x'], gold['mol/edge_feat'], gold['mol/att'], gold['mol/deg'] = xm, eam, attm, deg
    for tag, aggs, scalers in (('all_identity', ['mean', 'min', 'max', 'std', 'sum', 'var'], ['identity']),
                               ('scaled', ['mean', 'min', 'max', 'std'],
                                ['identity', 'amplification', 'attenuation', 'linear', 'inverse_linear'])):
        torch.manual_seed(5)
        conv = ns_pna['PNAConvSimple'](3 * Hp, Hp, aggs, scalers, deg, post_layers=1)
        gold[f'pnaconv/{tag}/state'] = {k: v.clone() for k, v in conv.state_dict().items()}
        gold[f'pnaconv/{tag}/out'] = conv(xm, mb.edge_index, eam, edge_atten=attm).detach()
    torch.manual_seed(5)
    conv = ns_pna['PNAConvSimple'](2 * Hp, Hp, ['mean', 'min', 'max', 'std'], ['identity'], deg, post_layers=2)
    gold['pnaconv/noea/state'] = {k: v.clone() for k, v in conv.state_dict().items()}
    gold['pnaconv/noea/out'] = conv(xm, mb.edge_index, None, edge_atten=None).detach()

    model_ns = {'torch': torch, 'nn': nn, 'F': F, 'Linear': nn.Linear, 'ReLU': nn.ReLU, 'Sequential': nn.Sequential,
                'ModuleList': nn.ModuleList, 'AtomEncoder': O.AtomEncoder, 'BondEncoder': O.BondEncoder,
                'BatchNorm': O.BatchNorm, 'global_mean_pool': O.global_mean_pool, 'global_add_pool': O.global_add_pool,
                'PNAConvSimple': ns_pna['PNAConvSimple'], 'GINConv': ns['GINConv'], 'GINEConv': ns['GINEConv'],
                'input': quiet, 'print': quiet}
    RefPNA = _compile(_extract(f'{REF}/src/models/pna.py', ['PNA']).values(), dict(model_ns))['PNA']
    RefGIN = _compile(_extract(f'{REF}/src/models/gin.py', ['GIN']).values(), dict(model_ns))['GIN']
    pna_cfg = {'model_name': 'PNA', 'hidden_size': 16, 'n_layers': 2, 'dropout_p': 0.3, 'atom_encoder': True,
               'use_edge_attr': True, 'aggregators': ['mean', 'min', 'max', 'std'], 'scalers': False, 'deg': deg}
    gin_cfg = {'model_name': 'GIN', 'hidden_size': 16, 'n_layers': 2, 'dropout_p': 0.3, 'atom_encoder': True,
               'use_edge_attr': True}
    for tag, Ref, cfg_ in (('pna', RefPNA, pna_cfg), ('gin', RefGIN, gin_cfg)):
        torch.manual_seed(6)
        m = Ref(9, 3, 2, False, cfg_)
        m.eval()
        gold[f'{tag}_model/config'] = cfg_
        gold[f'{tag}_model/state'] = {k: v.clone() for k, v in m.state_dict().items()}
        gold[f'{tag}_model/logits'] = m(mb.x, mb.edge_index, mb.batch, mb.edge_attr, edge_atten=attm).detach()
        gold[f'{tag}_model/emb'] = m.get_emb(mb.x, mb.edge_index, mb.batch, mb.edge_attr, edge_atten=attm).detach()
        m.train()                                     # training-mode BatchNorm (batch statistics); dropout off
        m.dropout_p = 0.0
        gold[f'{tag}_model/logits_train'] = m(mb.x, mb.edge_index, mb.batch, mb.edge_attr, edge_atten=attm).detach()

    # ---- the fork's GSAT class: __loss__ + dual_forward_pass --------------------------------------------------
    gm = _extract(f'{REF}/src/utils/get_model.py', ['Criterion', 'BatchSequential', 'MLP'])
    ns_gm = _compile(gm.values(), {'nn': nn, 'F': F, 'InstanceNorm': O.InstanceNorm, 'print': quiet})
    ns_u = {'torch': torch, 'sort_edge_index': O.sort_edge_index}
    _compile(_extract(f'{REF}/src/utils/utils.py', ['reorder_like']).values(), ns_u)
    run = _extract(f'{REF}/src/run_gsat.py', ['GSAT', 'ExtractorMLP'])
    ns_run = _compile(run.values(), {
        'torch': torch, 'nn': nn, 'F': F, 'np': np, 'Criterion': ns_gm['Criterion'], 'MLP': ns_gm['MLP'],
        'is_undirected': O.is_undirected, 'transpose': O.transpose, 'reorder_like': ns_u['reorder_like'],
        'print': quiet, 'input': quiet, 'plt': None, 'sns': None})
    RefGSAT, RefExtractor = ns_run['GSAT'], ns_run['ExtractorMLP']

    from tests.test_gpu_z_next_rows import _primal_dual_pair
    p, d = _primal_dual_pair(6, seed=8)
    cfg = {'model_name': 'GIN', 'hidden_size': 16, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    sc = {'learn_edge_att': False, 'extractor_dropout_p': 0.5, 'precision_k': 5, 'num_viz_samples': 0,
          'viz_interval': 10, 'viz_norm_att': True}
    mc = {'method_name': 'GSAT', 'pred_loss_coef': 1, 'info_loss_coef': 1, 'epochs': 100, 'decay_interval': 10,
          'decay_r': 0.1, 'final_r': 0.5, 'init_r': 0.9}
    torch.manual_seed(3)
    pc, pe = O.get_model(10, 0, 2, False, cfg), RefExtractor(16, sc, 'primal')
    dc, de = O.get_model(7, 0, 2, False, cfg), RefExtractor(16, sc, 'dual')
    ref = RefGSAT(pc, pe, None, None, None, 'cpu', None, 'mutag', 2, False, 0, mc, sc, cfg,
                  dc, de, None, None, None, 'cpu', None, 'mutag_dual', 2, False, 0, mc, sc, cfg)
    for m in (pc, pe, dc, de):                  # (the reference class re-defines .train() as its training loop)
        m.eval()
    for name, m in (('primal_clf', pc), ('primal_extractor', pe), ('dual_clf', dc), ('dual_extractor', de)):
        gold[f'dual/{name}/state'] = {k: v.clone() for k, v in m.state_dict().items()}
    for side, data in (('primal', p), ('dual', d)):
        for k in ('x', 'edge_index', 'batch', 'y', 'edge_label'):
            gold[f'dual/{side}/{k}'] = getattr(data, k)
    gold['dual/method_config'], gold['dual/shared_config'], gold['dual/model_config'] = mc, sc, cfg
    for epoch in (3, 57):
        seed = 100 + epoch
        torch.manual_seed(seed)                       # the draws the body will make, in its order (:204, :222, :229)
        u_primal = torch.empty(p.num_nodes, 1).uniform_(1e-10, 1 - 1e-10)
        U_dual = torch.rand(d.num_nodes, 1)
        torch.manual_seed(seed)
        edge_att, loss, loss_dict, logits = ref.dual_forward_pass(p, d, epoch, True)
        gold[f'dual/epoch{epoch}/primal_u'], gold[f'dual/epoch{epoch}/dual_U'] = u_primal, U_dual
        gold[f'dual/epoch{epoch}/primal_edge_att'] = edge_att.detach()
        gold[f'dual/epoch{epoch}/loss'] = loss.detach()
        gold[f'dual/epoch{epoch}/logits'] = logits.detach()
        gold[f'dual/epoch{epoch}/loss_dict'] = dict(loss_dict)
    # ---- per-batch explanation metrics of the trainer (run_gsat.py:783-800), tie-free attention ----------------------
    gm_ = torch.Generator().manual_seed(9)
    m_att = torch.rand(p.num_edges, generator=gm_)
    m_lab = (torch.rand(p.num_edges, generator=gm_) < 0.3).float()
    gold['metrics/att'], gold['metrics/labels'] = m_att, m_lab
    gold['metrics/edge_index'], gold['metrics/batch'] = p.edge_index, p.batch
    for k in (1, 5, 60):
        gold[f'metrics/precision_at_{k}'] = torch.tensor(
            RefGSAT.get_precision_at_k(None, m_att, m_lab, k, p.batch, p.edge_index), dtype=torch.float64)
    gold['metrics/delta_kl'] = torch.tensor(RefGSAT.get_delta_kl(None, m_lab, m_att), dtype=torch.float64)

    # ---- the fork's line-graph ("dual") construction loops, src/datasets/mutag_dual.py:342-378 -------------------------
    # statements of the dataset-reading method between those lines, executed as they stand on `dual_nodes` = the primal
    # edge list; the (a, b) node pairs they emit are mapped back to primal-edge indices (unique: no duplicate edges)
    tree = ast.parse(open(f'{REF}/src/datasets/mutag_dual.py').read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.lineno < 342 and n.end_lineno > 378][-1]
    stmts = [st_ for st_ in fn.body if 342 <= st_.lineno and st_.end_lineno <= 378]
    mod = ast.Module(body=stmts, type_ignores=[])
    ast.fix_missing_locations(mod)
    code = compile(mod, '<mutag_dual.py:342-378>', 'exec')
    for tag, batch_obj in (('ba2motifs', p), ('mol', mb)):
        prim = batch_obj.edge_index.t().contiguous().numpy()
        ns_lg = {'np': np, 'dual_nodes': [tuple(r) for r in prim.tolist()], 'print': quiet, 'input': quiet}
        exec(code, ns_lg)
        lut = {tuple(r): i for i, r in enumerate(prim.tolist())}
        assert len(lut) == prim.shape[0]
        de = [[lut[tuple(np.asarray(e1).tolist())], lut[tuple(np.asarray(e2).tolist())]] for e1, e2 in ns_lg['dual_edges']]
        gold[f'linegraph/{tag}/edge_index'], gold[f'linegraph/{tag}/batch'] = batch_obj.edge_index, batch_obj.batch
        gold[f'linegraph/{tag}/dual_edge_index'] = torch.tensor(de, dtype=torch.int64).t().contiguous()
        if tag == 'mol':          # both directions of every edge are consecutive rows: the relabelling of :535-549 applies
            stmts2 = [st_ for st_ in fn.body if 535 <= st_.lineno and st_.end_lineno <= 549]
            mod2 = ast.Module(body=stmts2, type_ignores=[])
            ast.fix_missing_locations(mod2)
            ns2 = {'dual_node_lists': [ns_lg['dual_nodes']], 'dual_edge_lists': [ns_lg['dual_edges']], 'print': quiet,
                   'input': quiet}
            exec(compile(mod2, '<mutag_dual.py:535-549>', 'exec'), ns2)
            halved = torch.tensor(ns2['dual_single_edge_lists'][0], dtype=torch.int64).t().contiguous() - 1   # 1-based ids
            gold[f'linegraph/{tag}/dual_edge_index_halved'] = halved

    # ---- the dense-matrix dual of src/datasets/ba_2motifs_dual.py:33-62 (+ dense_to_sparse of :69) ------------------------
    tree = ast.parse(open(f'{REF}/src/datasets/ba_2motifs_dual.py').read())
    fn2 = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == 'read_ba2motif_data'][0]
    stmts3 = [st_ for st_ in fn2.body if 26 <= st_.lineno and st_.end_lineno <= 62]
    mod3 = ast.Module(body=stmts3, type_ignores=[])
    ast.fix_missing_locations(mod3)
    from dp_gsat_b200.data import ba2motifs_batch as _ba
    bb = _ba(5, seed=11)
    n_per = 25
    dense = np.zeros((bb.num_graphs, n_per, n_per))
    gidx = bb.batch[bb.edge_index[0]].numpy()
    dense[gidx, bb.edge_index[0].numpy() - gidx * n_per, bb.edge_index[1].numpy() - gidx * n_per] = 1.0
    feats = np.random.default_rng(3).random((bb.num_graphs, n_per, 10)).astype(np.float32)
    ns3 = {'np': np, 'torch': torch, 'dense_edges': dense, 'node_features': feats,
           'print': quiet, 'input': quiet}
    exec(compile(mod3, '<ba_2motifs_dual.py:26-62>', 'exec'), ns3)
    duals, off = [], 0
    for dd in ns3['dual_dense_edges_list']:
        nz = torch.from_numpy(dd).nonzero().t().contiguous()          # dense_to_sparse: row-major non-zeros (:69)
        duals.append(nz + off)                                        # collate: cumulative dual-node offsets
        off += dd.shape[0]
    gold['densedual/edge_index'], gold['densedual/batch'] = bb.edge_index, bb.batch
    gold['densedual/dual_edge_index'] = torch.cat(duals, 1)
    gold['densedual/dual_node_label'] = torch.cat(ns3['dual_node_label_lists'])
    gold['densedual/x'] = torch.from_numpy(feats.reshape(-1, 10))
    gold['densedual/dual_x'] = torch.from_numpy(np.concatenate(ns3['dual_node_features_list'], 0)).float()

    torch.save(gold, os.path.join(HERE, 'ref_fork.pt'))
    print('golden keys:', len(gold), 'size', os.path.getsize(os.path.join(HERE, 'ref_fork.pt')))


if __name__ == '__main__':
    main()
