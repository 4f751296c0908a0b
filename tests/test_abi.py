"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/gsat_b200.h declares; host logic
(header parser, sharding, generators).  No compute calls: there is no GPU here."""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests.conftest import ROOT


def test_library_exports_every_declared_symbol():
    from dp_gsat_b200._lib import LIB_PATH, parse_header
    assert os.path.exists(LIB_PATH)
    protos = parse_header()
    assert len(protos) >= 20
    cdll = ctypes.CDLL(LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f'{name} declared in include/gsat_b200.h but not exported'
    cdll.gsatb_version.restype = ctypes.c_int
    assert cdll.gsatb_version() == 100
    cdll.gsatb_strerror.restype = ctypes.c_char_p
    assert b'workspace' in cdll.gsatb_strerror(-4)
    cdll.gsatb_index_build_workspace.restype = ctypes.c_size_t
    cdll.gsatb_index_build_workspace.argtypes = [ctypes.c_int64] * 3
    assert cdll.gsatb_index_build_workspace(100, 1000, 4) >= 7 * 4000


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle nor run without CUDA."""
    import dp_gsat_b200
    pkg = os.path.join(ROOT, 'dp_gsat_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            txt = open(os.path.join(pkg, fn)).read()
            assert 'import oracle' not in txt and 'from oracle' not in txt, fn
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            dp_gsat_b200.get_graph_index(torch.zeros((2, 3), dtype=torch.int64), torch.zeros(2, dtype=torch.int64))
        with pytest.raises(RuntimeError):
            dp_gsat_b200.ops.gin_aggregate(torch.zeros(2, 4), None, None, 0.0)
        with pytest.raises(RuntimeError):
            dp_gsat_b200.ops.embedding_sum(torch.zeros((2, 1), dtype=torch.int64), [torch.zeros(3, 4)])
        with pytest.raises(RuntimeError):
            dp_gsat_b200.ops.le_aggregate(torch.zeros(2, 4), torch.zeros(2, 4), None, None, None)
    for fn in os.listdir(pkg):          # the host SIMT emulator is test infrastructure: the product never references it
        if fn.endswith('.py') or fn.endswith('.cu'):
            assert 'tests.simt' not in open(os.path.join(pkg, fn)).read(), fn


def test_generators_shapes():
    from dp_gsat_b200.data import ba2motifs_batch, molhiv_like_batch, in_degree_histogram
    from oracle import gsat_oracle as O
    b = ba2motifs_batch(128, seed=0)
    assert b.num_nodes == 3200 and b.x.shape == (3200, 10) and b.y.shape == (128, 1)
    assert b.num_edges == 64 * 50 + 64 * 52
    idx = O.build_index_oracle(b.edge_index, b.batch)
    assert idx['symmetric'] and idx['graph_contiguous'] and not idx['has_dup']
    key = b.edge_index[0] * 3200 + b.edge_index[1]
    assert bool((key[1:] > key[:-1]).all())                      # dense_to_sparse (row-major) order
    assert float(b.edge_label.sum()) == 64 * 10 + 64 * 12
    m = molhiv_like_batch(64, seed=0)
    idx = O.build_index_oracle(m.edge_index, m.batch)
    assert idx['symmetric'] and idx['graph_contiguous'] and not idx['has_dup']
    assert m.x.dtype == torch.int64 and m.x.shape[1] == 9 and m.edge_attr.shape[1] == 3
    assert in_degree_histogram(m).sum() == m.num_nodes


def test_line_graph_dual_matches_definition():
    from dp_gsat_b200.data import line_graph_dual
    # 4-node graph of the comment at reference src/datasets/mutag_dual.py:181-193
    prim = np.array([[1, 2], [2, 1], [1, 3], [3, 1], [2, 4], [4, 2], [1, 4], [4, 1], [2, 3], [3, 2]]) - 1
    ds, dd, ng = line_graph_dual(prim[:, 0], prim[:, 1], np.zeros(4, dtype=np.int64))
    deg = np.bincount(prim[:, 0], minlength=4)
    assert ds.shape[0] == int((deg * (deg - 1)).sum())
    assert np.array_equal(ds[0::2], dd[1::2]) and np.array_equal(dd[0::2], ds[1::2])     # mutual reverses, adjacent
    assert all(prim[a, 0] == prim[b, 0] for a, b in zip(ds, dd))                          # share the first endpoint


def test_shard_bounds_balance_edges():
    from dp_gsat_b200.data import ba2motifs_batch, shard_batch
    b = ba2motifs_batch(64, seed=3)
    parts = [shard_batch(b, r, 4) for r in range(4)]
    assert sum(p.num_graphs for p in parts) == 64
    assert sum(p.num_edges for p in parts) == b.num_edges
    assert max(p.num_edges for p in parts) - min(p.num_edges for p in parts) <= 52
    for p in parts:
        assert int(p.edge_index.max()) < p.num_nodes and int(p.batch.max()) == p.num_graphs - 1


def test_every_entry_point_rejects_null_pointers_without_touching_the_device():
    """Error behaviour of the C ABI (include/gsat_b200.h: 'return a negative GSATB_E* code, nothing throws across the
    ABI'): every int-returning entry point called with NULL pointers and sizes of 1 must come back with GSATB_EINVAL
    before any CUDA work -- on this GPU-less machine too.  Runs in a child process so that a crash is a test failure,
    not the end of the test session."""
    import subprocess
    import sys
    code = r'''
import ctypes, importlib.util, sys
spec = importlib.util.spec_from_file_location('_b', sys.argv[1])
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
lib = ctypes.CDLL(m.LIB_PATH)
bad = []
for name, (restype, argtypes) in sorted(m.parse_header().items()):
    if restype is not ctypes.c_int or name in ('gsatb_version', 'gsatb_check_device', 'gsatb_set_step_counter',
                                               'gsatb_tc_set_profile_buffer', 'gsatb_ext_tile_slots',
                                                   'gsatb_gin_rows_supported'):      # (the last two are pure shape queries)
        continue
    fn = getattr(lib, name); fn.restype, fn.argtypes = restype, argtypes
    args = [None if t is ctypes.c_void_p else (0.0 if t is ctypes.c_float else 1) for t in argtypes]
    rc = fn(*args)
    if rc != -1:
        bad.append((name, rc))
print('CHECKED', bad)
sys.exit(1 if bad else 0)
'''
    r = subprocess.run([sys.executable, '-c', code, os.path.join(ROOT, 'dp_gsat_b200', '_lib.py')],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and 'CHECKED []' in r.stdout, r.stdout[-2000:]


def test_bench_roofline_table_from_recorded_events():
    """bench.build_roofline on a recorded event table: dominant kernel = largest share of the step, achieved = algorithmic
    bytes / launch time, fractions of the measured peak and of the nominal 8 TB/s, kernels without a byte model listed
    without a fraction."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Ev:
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t
    N, E, G, H = 4900000, 9996000, 196000, 128
    timer = {'gsatb_gin_aggregate_fwd': [(Ev(0), Ev(1.0), 'att'), (Ev(2), Ev(3.0), 'noatt')],
             'gsatb_tc_ext_bwd1': [(Ev(0), Ev(13.4), '')], 'gsatb_something_new': [(Ev(0), Ev(0.2), '')]}
    r = bench.build_roofline(timer, N, E, G, H, 1, 92.0)
    assert r['kernel'] == 'gsatb_tc_ext_bwd1' and r['bound'] == 'hbm' and r['unit'] == 'GB/s'
    nbytes = bench.kernel_models(N, E, G, H)['gsatb_tc_ext_bwd1'][0]
    assert abs(r['achieved'] - nbytes / 13.4e-3 / 1e9) < 1e-6 * r['achieved']
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12 and abs(r['frac_of_nominal_8TBs'] - r['achieved'] / 8000) < 1e-12
    rows = {k['kernel']: k for k in r['kernels']}
    assert rows['gsatb_gin_aggregate_fwd:att']['algorithmic_bytes_per_launch'] == 8.0 * N * H + 8.0 * E + 4.0 * N
    assert 'frac' not in rows['gsatb_something_new']
    assert abs(r['ours_share_of_step'] - (13.4 + 1.0 + 1.0 + 0.2) / 92.0) < 1e-9
