"""Kernel-logic tests on the host SIMT emulator (tests/simt): the SAME .cu sources the product library is built from,
compiled by g++ with -DGSATB_HOST_SIM, called through the same C ABI with host pointers, against the CPU oracle.
They check indexing, masks, sub-warp shuffles and reduction order without a GPU; the `-m gpu` suite repeats the
comparisons on the real device (tests/test_gpu_z_next_rows.py)."""
import pytest
import torch

from oracle import gsat_oracle as O
from tests.simt.simlib import sim, guarded, intact

OK = 0


def _graph(n_graphs=6, seed=0):
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(n_graphs, seed=seed)
    return b, O.build_index_oracle(b.edge_index, b.batch)


def test_emulator_known_answers():
    """The emulator itself (tests/simt/selftest.cpp): barriers + shared memory, sub-warp shuffles with divergent trip
    counts, ballot / up / down / broadcast shuffles in a partial last warp, 3-D grids."""
    import ctypes
    c = sim().cdll
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    x = torch.arange(1000, dtype=torch.int32)
    out = torch.zeros(3, dtype=torch.int32)
    c.simt_selftest_block_sum(P(x), P(out), 1000, 3, 256)
    assert int(out.sum()) == 999 * 1000 // 2
    for lpr in (1, 2, 4, 8, 16, 32):
        v = torch.arange(128, dtype=torch.float32) + 1
        o = torch.zeros(128 // lpr)
        c.simt_selftest_subwarp(P(v), P(o), lpr, 2, 64)
        grp = v.view(-1, lpr).sum(1)
        trips = torch.tensor([sum(range(1, g % 3 + 2)) for g in range(128 // lpr)], dtype=torch.float32)
        assert torch.equal(o, grp * trips), lpr
    o = torch.zeros(160, dtype=torch.int32)
    c.simt_selftest_ballot(P(o))
    o = o.view(40, 4)
    full = sum(1 << l for l in range(32) if l % 3 == 0)
    last = sum(1 << l for l in range(8) if (32 + l) % 3 == 0)
    for t in range(40):
        lane, base, width = t % 32, t - t % 32, (32 if t < 32 else 8)
        assert (int(o[t, 0]) & 0xffffffff) == (full if t < 32 else last)
        assert int(o[t, 1]) == (t - 1 if lane >= 1 else t)
        assert int(o[t, 2]) == (t + 2 if lane + 2 < width else t) or (t >= 32 and lane + 2 < 32)
        assert int(o[t, 3]) == base + 5
    o = torch.zeros(12, dtype=torch.int32)
    c.simt_selftest_grid3(P(o))
    assert torch.equal(o, torch.full((12,), sum(range(8)), dtype=torch.int32))
    for name in ('gsatb_le_aggregate_fwd', 'gsatb_le_aggregate_bwd'):
        assert sim().has(name)


@pytest.mark.parametrize('H', [4, 16, 64, 132, 300])
@pytest.mark.parametrize('with_w,with_att,with_add', [(True, True, True), (False, True, False), (True, False, True),
                                                      (False, False, False)])
def test_leconv_aggregate_on_emulator(H, with_w, with_att, with_add):
    b, ix = _graph(5 if H > 64 else 9, seed=H)
    N, E = b.num_nodes, b.num_edges
    g = torch.Generator().manual_seed(H + 1)
    a, bb = torch.randn(N, H, generator=g), torch.randn(N, H, generator=g)
    w = torch.rand(E, generator=g) + 0.5 if with_w else None
    att = torch.rand(E, 1, generator=g) if with_att else None
    add = torch.randn(N, H, generator=g) if with_add else None
    gout = torch.randn(N, H, generator=g)
    # oracle: the message / aggregation lines of conv_layers.py:69-92 with identity linears
    ar, br = a.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True) if with_w else None
    tr = att.clone().requires_grad_(True) if with_att else None
    m = ar[b.edge_index[0]] - br[b.edge_index[1]]
    if with_w:
        m = m * wr.view(-1, 1)
    if with_att:
        m = m * tr
    ref = O.scatter_sum(m, b.edge_index[1], N)
    if with_add:
        ref = ref + add
    (ref * gout).sum().backward()

    out = guarded((N, H))
    rc = sim().call('gsatb_le_aggregate_fwd', a, bb, w, None if att is None else att.view(-1).contiguous(),
                    ix['rowptr_dst'], ix['eid_by_dst'], ix['src_by_dst'], add, out, N, E, H, None)
    assert rc == OK and intact(out)
    assert torch.allclose(out, ref.detach(), rtol=1e-5, atol=1e-5)

    da, db = guarded((N, H)), guarded((N, H))
    dw = guarded((E,)) if with_w else None
    datt = guarded((E,)) if with_att else None
    rc = sim().call('gsatb_le_aggregate_bwd', gout, a, bb, w, None if att is None else att.view(-1).contiguous(),
                    ix['rowptr_src'], ix['eid_by_src'], ix['dst_by_src'], ix['rowptr_dst'], ix['eid_by_dst'],
                    da, db, dw, datt, N, E, H, None)
    assert rc == OK and intact(da) and intact(db)
    assert torch.allclose(da, ar.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(db, br.grad, rtol=1e-5, atol=1e-5)
    if with_w:
        assert intact(dw) and torch.allclose(dw, wr.grad, rtol=1e-5, atol=2e-5 * H ** 0.5)
    if with_att:
        assert intact(datt) and torch.allclose(datt, tr.grad.view(-1), rtol=1e-5, atol=2e-5 * H ** 0.5)


def test_leconv_argument_checks_on_emulator():
    b, ix = _graph(2)
    N, E = b.num_nodes, b.num_edges
    x = torch.zeros(N, 6)
    out = torch.zeros(N, 6)
    rc = sim().call('gsatb_le_aggregate_fwd', x, x, None, None, ix['rowptr_dst'], ix['eid_by_dst'], ix['src_by_dst'],
                    None, out, N, E, 6, None)
    assert rc == -2                       # GSATB_ESHAPE: width not a multiple of 4
    rc = sim().call('gsatb_le_aggregate_fwd', None, x, None, None, ix['rowptr_dst'], ix['eid_by_dst'],
                    ix['src_by_dst'], None, out, N, E, 8, None)
    assert rc == -1                       # GSATB_EINVAL
    rc = sim().call('gsatb_le_aggregate_fwd', None, None, None, None, None, None, None, None, None, 0, 0, 8, None)
    assert rc == OK                       # empty batch


# ---------------------------------------------------------------------------------------------------------------
# the Python autograd wrappers on top of the emulated library: argument order of the ctypes calls (every pointer is a
# void* there, nothing else would catch a swapped pair before the GPU run), saved tensors, gradient routing
# ---------------------------------------------------------------------------------------------------------------
class _FakeIndex:
    """The GraphIndex attributes the wrappers read, filled from the oracle's index specification."""

    def __init__(self, batch_obj):
        ix = O.build_index_oracle(batch_obj.edge_index, batch_obj.batch)
        for k, v in ix.items():
            setattr(self, k, v)
        self.N, self.E, self.G = batch_obj.num_nodes, batch_obj.num_edges, batch_obj.num_graphs


@pytest.fixture
def sim_ops(monkeypatch):
    """dp_gsat_b200.ops (and the rest of the package) on top of the emulated library, host tensors allowed."""
    from tests.simt import emulate
    emulate.patch_product(monkeypatch.setattr)
    import dp_gsat_b200.ops as ops
    return ops


@pytest.mark.parametrize('with_w,with_att,with_add', [(True, True, True), (False, True, False), (True, False, True)])
def test_le_aggregate_autograd_wrapper_on_emulator(sim_ops, with_w, with_att, with_add):
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(7, seed=4)
    gi = _FakeIndex(b)
    N, E, H = b.num_nodes, b.num_edges, 32
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: torch.randn(*s, generator=g)
    a, bb, add, gout = mk(N, H), mk(N, H), (mk(N, H) if with_add else None), mk(N, H)
    w = torch.rand(E, 1, generator=g) + 0.5 if with_w else None
    att = torch.rand(E, 1, generator=g) if with_att else None
    leaf = lambda t: None if t is None else t.clone().requires_grad_(True)
    ar, br, dr, wr, tr = leaf(a), leaf(bb), leaf(add), leaf(w), leaf(att)
    m = ar[b.edge_index[0]] - br[b.edge_index[1]]
    if with_w:
        m = m * wr.view(-1, 1)
    if with_att:
        m = m * tr
    ref = O.scatter_sum(m, b.edge_index[1], N)
    if with_add:
        ref = ref + dr
    (ref * gout).sum().backward()
    a2, b2, d2, w2, t2 = leaf(a), leaf(bb), leaf(add), leaf(w), leaf(att)
    out = sim_ops.le_aggregate(a2, b2, w2, t2, gi, add=d2)
    (out * gout).sum().backward()
    tol = dict(rtol=1e-5, atol=2e-5)
    assert torch.allclose(out, ref, **tol)
    assert torch.allclose(a2.grad, ar.grad, **tol) and torch.allclose(b2.grad, br.grad, **tol)
    if with_add:
        assert torch.allclose(d2.grad, dr.grad, **tol)
    if with_w:
        assert w2.grad.shape == w.shape and torch.allclose(w2.grad, wr.grad, rtol=1e-5, atol=2e-4)
    if with_att:
        assert t2.grad.shape == att.shape and torch.allclose(t2.grad, tr.grad, rtol=1e-5, atol=2e-4)


def test_leconv_module_on_emulator(sim_ops):
    """nn.LEConv (reference conv_layers.py:69-92) end to end on the emulated kernels: same parameters as the oracle
    layer, forward and parameter gradients."""
    import dp_gsat_b200 as G
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(5, seed=1)
    gi = _FakeIndex(b)
    torch.manual_seed(0)
    ref = O.LEConv(16, 16)
    dev = G.LEConv(16, 16)
    dev.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(2)
    x = torch.randn(b.num_nodes, 16, generator=g)
    w, att = torch.rand(b.num_edges, 1, generator=g), torch.rand(b.num_edges, 1, generator=g)
    want = ref(x, b.edge_index, edge_weight=w, edge_atten=att)
    got = dev(x, b.edge_index, edge_weight=w, edge_atten=att, _index=gi)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    want.square().sum().backward()
    got.square().sum().backward()
    for (k, p), (_, q) in zip(dev.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-4), k


# ---------------------------------------------------------------------------------------------------------------
# fused atom / bond encoders, device-side collate
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,H,dims', [(1, 4, [3]), (37, 16, [5, 6, 2]), (700, 80, [119, 4, 12, 12, 10, 6, 6, 2, 2]),
                                      (300, 132, [200, 7]), (0, 8, [4, 4])])
def test_embedding_sum_on_emulator(M, H, dims):
    """out[m] = sum_k table_k[idx[m, k]]: forward BIT-exact against the reference's loop (ogb AtomEncoder.forward:
    x_embedding = 0; x_embedding += emb_k(x[:, k])), backward against autograd of the same loop; tables larger than one
    shared-memory window (200 rows) and widths that are not a multiple of the 64-channel slab included."""
    g = torch.Generator().manual_seed(M + H)
    K = len(dims)
    idx = torch.stack([torch.randint(0, d, (M,), generator=g) for d in dims], dim=1).contiguous()
    tables = [torch.randn(d, H, generator=g).requires_grad_(True) for d in dims]
    ref = 0
    for k in range(K):
        ref = ref + tables[k][idx[:, k]]
    gout = torch.randn(M, H, generator=g)
    if M:
        (ref * gout).sum().backward()
    offs = torch.zeros(K + 1, dtype=torch.int32)
    offs[1:] = torch.cumsum(torch.tensor(dims), 0).to(torch.int32)
    cat = torch.cat([t.detach() for t in tables], 0).contiguous()
    out, flag = guarded((M, H)), torch.zeros(1, dtype=torch.int32)
    rc = sim().call('gsatb_embedding_sum_fwd', idx, cat, offs, out, flag, M, K, H, None)
    assert rc == OK and intact(out) and int(flag) == 0
    if M:
        assert torch.equal(out, ref.detach())
    R = int(offs[-1])
    from tests.simt.emulate import emulated_lib
    ws_bytes = int(emulated_lib().cdll.gsatb_embedding_sum_bwd_workspace(M, R, H))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8)
    dt = guarded((R, H))
    import ctypes
    rc = sim().call('gsatb_embedding_sum_bwd', gout, idx, offs, dt, M, K, H, ws, ctypes.c_size_t(ws_bytes), None)
    assert rc == OK and intact(dt)
    # ground truth in fp64 (the fp32 index_add of autograd is itself ~1e-5 off on rows that collect hundreds of terms,
    # and its accumulation order depends on the thread count)
    t64 = [t.detach().double().requires_grad_(True) for t in tables]
    if M:
        sum((t64[k][idx[:, k]] * gout.double()).sum() for k in range(K)).backward()
    want = torch.cat([t.grad if t.grad is not None else torch.zeros_like(t) for t in t64], 0)
    assert torch.allclose(dt.double(), want, rtol=1e-5, atol=2e-6 * max(1.0, float(want.abs().max())))
    rc = sim().call('gsatb_embedding_sum_bwd', gout, idx, offs, dt, M, K, H, ws, ctypes.c_size_t(16), None)
    assert rc == -4                                       # GSATB_EWS_TOO_SMALL


def test_embedding_sum_clamps_and_flags_out_of_range_indices():
    dims, H = [3, 5], 8
    idx = torch.tensor([[0, 4], [3, 1], [-1, 7]], dtype=torch.int64)          # 3 >= dims[0], -1, 7 >= dims[1]
    cat = torch.arange(8 * H, dtype=torch.float32).view(8, H).contiguous()
    offs = torch.tensor([0, 3, 8], dtype=torch.int32)
    out, flag = guarded((3, H)), torch.zeros(1, dtype=torch.int32)
    assert sim().call('gsatb_embedding_sum_fwd', idx, cat, offs, out, flag, 3, 2, H, None) == OK
    assert int(flag) == 1 and intact(out)
    assert torch.equal(out[0], cat[0] + cat[3 + 4])
    assert torch.equal(out[1], cat[2] + cat[3 + 1])                         # clamped to the last row of table 0
    assert torch.equal(out[2], cat[0] + cat[3 + 4])                         # clamped to row 0 / the last row of table 1
    bad = torch.tensor([0, 3, 3], dtype=torch.int32)                          # an empty table
    assert sim().call('gsatb_embedding_sum_fwd', idx, cat, bad, out, flag, 3, 2, H, None) == -1


def test_fused_encoders_autograd_wrapper_on_emulator(sim_ops):
    """nn.AtomEncoder / BondEncoder with fused = True against their own library-lookup path (the reference's loop):
    forward bit-exact, the gradient of every table within fp32 reordering."""
    import dp_gsat_b200 as G
    from dp_gsat_b200.data import molhiv_like_batch
    b = molhiv_like_batch(6, seed=3, with_edge_attr=True)
    for enc, idx in ((G.AtomEncoder(20), b.x), (G.BondEncoder(12), b.edge_attr)):
        w = torch.randn(idx.shape[0], enc._tables()[0].shape[1], generator=torch.Generator().manual_seed(1))
        enc.fused = False
        ref = enc(idx)
        (ref * w).sum().backward()
        want = [t.grad.clone() for t in enc._tables()]
        enc.zero_grad()
        enc.fused = True
        out = enc(idx)
        assert torch.equal(out, ref)
        (out * w).sum().backward()
        for t, g0 in zip(enc._tables(), want):
            assert torch.allclose(t.grad, g0, rtol=1e-5, atol=1e-5)


def _sample_graphs(seed=0):
    from dp_gsat_b200.data import molhiv_like_batch
    from dp_gsat_b200.loader import split_batch
    b = molhiv_like_batch(9, seed=seed, with_edge_attr=True)
    b.edge_label = (torch.arange(b.num_edges) % 3 == 0).float()
    b.node_label = torch.arange(b.num_nodes, dtype=torch.float32)
    return split_batch(b)


@pytest.mark.parametrize('ids', [[0], [3, 1, 4, 1, 5], list(range(9)), [8, 7, 6, 5, 4, 3, 2, 1, 0]])
def test_device_collate_on_emulator(sim_ops, ids):
    """PackedDataset.collate(ids) (csrc/collate.cu) is bit-identical to the restated PyG Batch.from_data_list of the
    same graphs: x, edge_index with cumulative node offsets, batch vector, y, edge_attr, edge / node labels; repeated
    ids, reversed order and single-graph batches included."""
    from dp_gsat_b200.loader import PackedDataset
    graphs = _sample_graphs()
    ds = PackedDataset.from_data_list(graphs, device='cpu')
    got = ds.collate(ids)
    want = O.collate_data_list([graphs[i] for i in ids])
    assert got.num_graphs == len(ids)
    for k in ('x', 'edge_index', 'batch', 'y', 'edge_attr', 'edge_label', 'node_label'):
        assert torch.equal(getattr(got, k), want[k]), k
        assert getattr(got, k).dtype == want[k].dtype


def test_device_collate_edge_cases_on_emulator(sim_ops):
    """Graphs without edges, a dataset without optional tensors, float features, the loader's batching."""
    from dp_gsat_b200.loader import PackedDataset, DeviceLoader, Graph
    g = torch.Generator().manual_seed(0)
    graphs = []
    for n, e in ((3, 4), (1, 0), (5, 7), (2, 0), (4, 12)):
        ei = torch.randint(0, n, (2, e), generator=g)
        graphs.append(Graph(torch.rand(n, 5, generator=g), ei, torch.randint(0, 3, (1,), generator=g)))
    ds = PackedDataset.from_data_list(graphs, device='cpu')
    for ids in ([1], [1, 3], [0, 1, 2, 3, 4], [3, 4, 1]):
        got, want = ds.collate(ids), O.collate_data_list([graphs[i] for i in ids])
        for k in ('x', 'edge_index', 'batch', 'y'):
            assert torch.equal(getattr(got, k), want[k]), (ids, k)
        assert got.edge_attr is None and got.edge_label is None
    batches = list(DeviceLoader(ds, ids=[4, 3, 2, 1, 0], batch_size=2))
    assert [b.num_graphs for b in batches] == [2, 2, 1]
    assert torch.equal(batches[2].x, graphs[0].x)
    with pytest.raises(IndexError):
        ds.collate([5])
    with pytest.raises(ValueError):
        PackedDataset.from_data_list([Graph(torch.rand(2, 3), torch.tensor([[0], [2]]), torch.zeros(1))], device='cpu')


def test_cached_loader_builds_each_batch_and_its_index_once(sim_ops):
    """DeviceLoader(cache=True): the second epoch hands out the SAME resident batches (same addresses), so the K0 index of
    every batch is built once for the whole run even when an epoch has more batches than the default index cache."""
    import dp_gsat_b200 as G
    from dp_gsat_b200.loader import PackedDataset, DeviceLoader, split_batch
    from dp_gsat_b200.data import ba2motifs_batch
    from tests.simt.emulate import emulated_lib
    old_cap = G.set_index_cache_capacity(8)
    try:
        ds = PackedDataset.from_data_list(split_batch(ba2motifs_batch(24, seed=1)), device='cpu')
        loader = DeviceLoader(ds, batch_size=2, cache=True)                  # 12 batches > the default capacity of 8
        assert len(loader) == 12
        calls = {'n': 0}
        emu = emulated_lib()
        orig = emu.call

        def counting(name, *a):
            calls['n'] += name == 'gsatb_index_build'
            return orig(name, *a)
        emu.call = counting
        try:
            first = [(b, G.get_graph_index(b.edge_index, b.batch, b.num_graphs)) for b in loader]
            second = [(b, G.get_graph_index(b.edge_index, b.batch, b.num_graphs)) for b in loader]
        finally:
            emu.call = orig
        assert calls['n'] == 12
        for (b1, g1), (b2, g2) in zip(first, second):
            assert b1 is b2 and g1 is g2 and b1.x.data_ptr() == b2.x.data_ptr()
        with pytest.raises(ValueError):
            DeviceLoader(ds, batch_size=2, cache=True, shuffle=True)
        plain = list(DeviceLoader(ds, batch_size=2))
        assert all(torch.equal(a.edge_index, b[0].edge_index) for a, b in zip(plain, first))
    finally:
        G.set_index_cache_capacity(old_cap)
        G.clear_index_cache()


def test_device_step_counter_feeds_the_in_kernel_random_streams(sim_ops):
    """gsatb_set_step_counter: the device-resident counter is added to the sampler's Philox offset INSIDE the kernel, which
    is what lets a CUDA graph of the whole step (frozen kernel arguments) draw fresh noise on every replay."""
    import dp_gsat_b200 as G
    from tests.simt.emulate import emulated_lib
    counter = emulated_lib().step_counter()
    counter.zero_()
    try:
        logits = torch.zeros(4096, 1)
        a = G.concrete_sample(logits, 1, True, seed=5, offset=100)
        assert torch.equal(a, G.concrete_sample(logits, 1, True, seed=5, offset=100))      # same counter: same draw
        counter.add_(1)
        b = G.concrete_sample(logits, 1, True, seed=5, offset=100)
        assert not torch.equal(a, b) and abs(float(b.mean()) - 0.5) < 0.03                 # fresh, still uniform noise
        counter.sub_(1)
        assert torch.equal(a, G.concrete_sample(logits, 1, True, seed=5, offset=100))
        assert torch.equal(G.concrete_sample(logits, 1, False), torch.full_like(logits, 0.5))   # eval: no noise at all
    finally:
        counter.zero_()
