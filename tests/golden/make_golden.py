"""Generate the committed golden fixtures under tests/golden/.  Run ONCE in the build container, where
/root/reference exists:   python tests/golden/make_golden.py

What is pinned, and by what:

* ``ref_functions.pt`` -- outputs of the REFERENCE'S OWN function bodies, extracted with ``ast`` from the files
  under /root/reference and executed unmodified: ``reorder_like`` (src/utils/utils.py:19-25), ``concrete_sample``,
  ``get_r``, ``lift_node_att_to_edge_att``, ``gumbel_sigmoid``, ``f1_sparsity_loss`` (src/run_gsat.py:151-187,
  860-885), ``GSAT.__loss__`` (example/gsat.py:27-35), ``Criterion`` / ``BatchSequential`` / ``MLP``
  (src/utils/get_model.py:19-68) and ``ExtractorMLP.forward`` (example/gsat.py:120-139).  The reference modules
  cannot be imported whole (they import torch_geometric / rdkit / matplotlib at module top, none installed), so the
  third-party names those bodies call (``sort_edge_index``, ``InstanceNorm``) are bound to the oracle's
  restatements -- those third-party semantics remain UNPINNED.
* ``mutag_slice.npz`` -- the first 512 graphs of the reference's in-tree Mutagenicity topology
  (data/mutag_dual/raw/Mutagenicity_A.txt + Mutagenicity_graph_indicator.txt): real edge order, unsorted,
  consecutive rows mutual reverses (=> rev[e] == e ^ 1).
* ``mutag_full_summary.json`` -- size-independent facts of the FULL file computed here with the oracle
  (E, N, G, symmetric, rev == e^1, checksums), for the record.
"""
import ast
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import gsat_oracle as O  # noqa: E402

REF = '/root/reference'


def _extract(path, names, cls=None):
    """Return {name: ast node} for top-level (or class-level, if ``cls``) defs in a reference file."""
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = [n for n in body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    out = {}
    for n in body:
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names:
            if isinstance(n, ast.FunctionDef):
                n.decorator_list = []
            out[n.name] = n
    return out


def _compile(nodes, ns):
    mod = ast.Module(body=list(nodes), type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, '<reference>', 'exec'), ns)
    return ns


def main():
    torch.manual_seed(0)
    gold = {}

    # ---- reorder_like (reference body) on several edge lists -------------------------------------------
    ns = {'torch': torch, 'sort_edge_index': O.sort_edge_index}
    _compile(_extract(f'{REF}/src/utils/utils.py', ['reorder_like']).values(), ns)
    reorder_like = ns['reorder_like']
    # 4-node graph from the comment at src/datasets/mutag_dual.py:181-193 (0-based)
    kat = torch.tensor([[1, 2], [2, 1], [1, 3], [3, 1], [2, 4], [4, 2], [1, 4], [4, 1], [2, 3], [3, 2]]).t() - 1
    cases = {'kat4': kat}
    g = torch.Generator().manual_seed(5)
    for name, n, m in (('rand_a', 12, 20), ('rand_b', 40, 90)):
        und = set()
        while len(und) < m:
            a, b = torch.randint(0, n, (2,), generator=g).tolist()
            if a != b:
                und.add((min(a, b), max(a, b)))
        und = torch.tensor(sorted(und)).t()
        ei = torch.cat([und, und.flip(0)], dim=1)
        ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
        cases[name] = ei
    for name, ei in cases.items():
        vals = torch.arange(ei.shape[1], dtype=torch.float32).view(-1, 1) * 0.5 + 1.0
        t_idx = torch.stack([ei[1], ei[0]])
        out = reorder_like(t_idx, ei, vals)
        gold[f'reorder_like/{name}/edge_index'] = ei
        gold[f'reorder_like/{name}/values'] = vals
        gold[f'reorder_like/{name}/out'] = out

    # ---- GSAT static pieces from src/run_gsat.py --------------------------------------------------------
    fns = _extract(f'{REF}/src/run_gsat.py',
                   ['concrete_sample', 'get_r', 'lift_node_att_to_edge_att', 'gumbel_sigmoid', 'f1_sparsity_loss'],
                   cls='GSAT')
    ns = _compile(fns.values(), {'torch': torch, 'input': lambda *a: None})

    class Dummy:
        pass
    logits = torch.randn(64, 1, generator=torch.Generator().manual_seed(11)) * 2
    torch.manual_seed(123)
    u = torch.empty_like(logits).uniform_(1e-10, 1 - 1e-10)       # the draw concrete_sample will make
    torch.manual_seed(123)
    gold['concrete/logits'] = logits
    gold['concrete/u'] = u
    gold['concrete/train'] = ns['concrete_sample'](logits, 1, True)
    gold['concrete/eval'] = ns['concrete_sample'](logits, 1, False)
    gold['get_r'] = torch.tensor([[e, ns['get_r'](Dummy(), 10, 0.1, e, final_r=0.5)] for e in range(0, 80, 5)],
                                 dtype=torch.float64)
    gold['get_r_init07'] = torch.tensor([[e, ns['get_r'](Dummy(), 10, 0.1, e, init_r=0.9, final_r=0.7)]
                                         for e in range(0, 80, 5)], dtype=torch.float64)
    node_att = torch.rand(5, 1, generator=torch.Generator().manual_seed(3))
    gold['lift/node_att'] = node_att
    gold['lift/edge_index'] = kat
    gold['lift/out'] = ns['lift_node_att_to_edge_att'](node_att[:4], kat)
    torch.manual_seed(77)
    U = torch.rand_like(logits)
    torch.manual_seed(77)
    gold['gumbel/U'] = U
    gold['gumbel/out_tau0.1'] = ns['gumbel_sigmoid'](Dummy(), logits, tau=0.1)
    p = torch.rand(64, 1, generator=torch.Generator().manual_seed(8))
    yv = (torch.rand(64, generator=torch.Generator().manual_seed(9)) < 0.3).float()
    gold['f1/p'], gold['f1/y'] = p, yv
    gold['f1/out'] = ns['f1_sparsity_loss'](Dummy(), p, yv)

    # ---- Criterion / BatchSequential / MLP from src/utils/get_model.py ----------------------------------
    cls = _extract(f'{REF}/src/utils/get_model.py', ['Criterion', 'BatchSequential', 'MLP'])
    ns = _compile(cls.values(), {'nn': nn, 'F': F, 'InstanceNorm': O.InstanceNorm, 'print': lambda *a, **k: None})
    crit = ns['Criterion'](2, False)
    cl = torch.randn(16, 1, generator=torch.Generator().manual_seed(21))
    cy = (torch.rand(16, 1, generator=torch.Generator().manual_seed(22)) < 0.5).float()
    gold['criterion/logits'], gold['criterion/y'] = cl, cy
    gold['criterion/out'] = crit(cl, cy)
    torch.manual_seed(4)
    mlp = ns['MLP']([16, 32, 8, 1], dropout=0.5)
    mlp.eval()
    feats = torch.randn(30, 16, generator=torch.Generator().manual_seed(31))
    seg = torch.tensor([0] * 7 + [1] * 11 + [2] * 12)
    gold['mlp/state'] = {k: v.clone() for k, v in mlp.state_dict().items()}
    gold['mlp/x'], gold['mlp/batch'] = feats, seg
    gold['mlp/out_eval'] = mlp(feats, seg).detach()

    # ---- example GSAT.__loss__ and ExtractorMLP.forward (example/gsat.py) --------------------------------
    loss_fn = _extract(f'{REF}/example/gsat.py', ['__loss__', 'get_r'], cls='GSAT')
    ns2 = _compile(loss_fn.values(), {'torch': torch})
    d = Dummy()
    d.criterion, d.decay_interval, d.decay_r, d.final_r = crit, 10, 0.1, 0.7
    d.get_r = ns2['get_r']
    att = torch.rand(64, 1, generator=torch.Generator().manual_seed(41)) * 0.98 + 0.01
    loss, ld = ns2['__loss__'](d, att, cl, cy, 25)
    gold['loss/att'] = att
    gold['loss/epoch'] = torch.tensor(25)
    gold['loss/total'] = loss
    gold['loss/pred'] = torch.tensor(ld['pred'])
    gold['loss/info'] = torch.tensor(ld['info'])
    ext = _extract(f'{REF}/example/gsat.py', ['ExtractorMLP'])
    ns3 = _compile(ext.values(), {'torch': torch, 'nn': nn, 'MLP': ns['MLP']})
    torch.manual_seed(6)
    ex = ns3['ExtractorMLP'](8, True)
    ex.eval()
    emb = torch.randn(4, 8, generator=torch.Generator().manual_seed(51))
    gold['extractor/state'] = {k: v.clone() for k, v in ex.state_dict().items()}
    gold['extractor/emb'] = emb
    gold['extractor/edge_index'] = kat
    gold['extractor/batch'] = torch.zeros(4, dtype=torch.long)
    gold['extractor/out_eval'] = ex(emb, kat, torch.zeros(4, dtype=torch.long)).detach()

    torch.save(gold, os.path.join(HERE, 'ref_functions.pt'))

    # ---- Mutagenicity topology ----------------------------------------------------------------------------
    A = np.loadtxt(f'{REF}/data/mutag_dual/raw/Mutagenicity_A.txt', delimiter=',', dtype=np.int64) - 1
    gi = np.loadtxt(f'{REF}/data/mutag_dual/raw/Mutagenicity_graph_indicator.txt', dtype=np.int64) - 1
    K = 512
    n_keep = int(np.searchsorted(gi, K, side='left'))
    m = (A[:, 0] < n_keep) & (A[:, 1] < n_keep)
    np.savez_compressed(os.path.join(HERE, 'mutag_slice.npz'), src=A[m, 0].astype(np.int32),
                        dst=A[m, 1].astype(np.int32), node_graph=gi[:n_keep].astype(np.int32))
    ei = torch.from_numpy(A.T.copy())
    rev = O.reverse_edge_permutation(ei)
    E = ei.shape[1]
    idx = O.build_index_oracle(ei, torch.from_numpy(gi))
    summary = {
        'E': int(E), 'N': int(gi.shape[0]), 'G': int(gi.max()) + 1,
        'is_undirected': O.is_undirected(ei),
        'rev_is_xor1': bool((rev == (torch.arange(E) ^ 1)).all()),
        'rev_matches_index_builder': bool((idx['rev'].long() == rev).all()),
        'checksum_eid_by_dst': int((idx['eid_by_dst'].long() * (torch.arange(E) % 1000003 + 1)).sum() % (2 ** 61 - 1)),
        'checksum_rowptr_dst': int(idx['rowptr_dst'].long().sum()),
        'slice_graphs': K, 'slice_nodes': n_keep, 'slice_edges': int(m.sum()),
    }
    json.dump(summary, open(os.path.join(HERE, 'mutag_full_summary.json'), 'w'), indent=1)
    print(json.dumps(summary, indent=1))
    print('golden keys:', len(gold))


if __name__ == '__main__':
    main()
