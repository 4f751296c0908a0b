"""Property-based tests (hypothesis) of the kernels on the host SIMT emulator: random small batches -- ragged graphs,
isolated nodes, graphs without nodes or edges, duplicate and self-loop edges, directed and symmetric edge sets, widths
that are not a power of two -- against the CPU oracle.  Same bars as the GPU suite: index arrays bit-exact, fp32 values
rtol 1e-5.  (Emulator scope and limits: DESIGN.md section 2a.)"""
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import gsat_oracle as O

SETTINGS = dict(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture,
                                                                        HealthCheck.too_slow])


@pytest.fixture
def G(monkeypatch):
    from tests.simt import emulate
    emulate.patch_product(monkeypatch.setattr)
    import dp_gsat_b200 as g
    yield g
    g.clear_index_cache()


@st.composite
def batches(draw, symmetric=None, allow_empty_graphs=True):
    """(edge_index [2, E], batch [N], num_graphs): nodes grouped by graph, edges grouped by graph and inside one graph."""
    n_graphs = draw(st.integers(1, 5))
    sizes = [draw(st.integers(0 if allow_empty_graphs else 1, 9)) for _ in range(n_graphs)]
    if sum(sizes) == 0:
        sizes[0] = 1
    sym = draw(st.booleans()) if symmetric is None else symmetric
    src, dst, batch, off = [], [], [], 0
    for g, n in enumerate(sizes):
        batch += [g] * n
        if n:
            m = draw(st.integers(0, 3 * n))
            for _ in range(m):
                a, b = draw(st.integers(0, n - 1)), draw(st.integers(0, n - 1))
                src.append(off + a)
                dst.append(off + b)
                if sym and a != b:
                    src.append(off + b)
                    dst.append(off + a)
        off += n
    ei = torch.tensor([src, dst], dtype=torch.int64).reshape(2, -1)
    return ei, torch.tensor(batch, dtype=torch.int64), n_graphs


def close(a, b, rtol=1e-5, atol=2e-6):
    scale = max(1.0, float(b.abs().max())) if b.numel() else 1.0
    return torch.allclose(a.double(), b.double(), rtol=rtol, atol=atol * scale)


@settings(**SETTINGS)
@given(batches())
def test_index_build_matches_the_specification(G, case):
    ei, batch, ng = case
    ref = O.build_index_oracle(ei, batch, ng)
    gi = G.GraphIndex(ei, batch, ng)
    for k in ('src', 'dst', 'rowptr_dst', 'eid_by_dst', 'src_by_dst', 'rowptr_src', 'eid_by_src', 'dst_by_src',
              'node_ptr', 'edge_ptr', 'edge_graph'):
        assert torch.equal(getattr(gi, k), ref[k]), k
    assert gi.symmetric == ref['symmetric'] == O.is_undirected(ei)
    assert gi.has_duplicates == ref['has_dup']
    assert gi.graph_contiguous == ref['graph_contiguous']
    if ref['symmetric'] and not ref['has_dup']:
        assert torch.equal(gi.rev, ref['rev'])
        assert torch.equal(gi.rev[gi.rev.long()].long(), torch.arange(ei.shape[1]))       # an involution


@settings(**SETTINGS)
@given(batches(), st.sampled_from([4, 8, 20, 36, 132]), st.booleans(), st.integers(0, 2 ** 31 - 1))
def test_gin_aggregate_and_pool(G, case, H, with_att, seed):
    ei, batch, ng = case
    N, E = batch.numel(), ei.shape[1]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, H, generator=g)
    att = torch.rand(E, 1, generator=g) if with_att else None
    w = torch.randn(N, H, generator=g)
    xr = x.clone().requires_grad_(True)
    ar = att.clone().requires_grad_(True) if with_att else None
    ref = O.GINConv(torch.nn.Identity())(xr, ei, edge_atten=ar)
    pooled_ref = O.global_add_pool(ref, batch, ng)
    mean_ref = O.global_mean_pool(ref, batch, ng)
    ((ref * w).sum() + pooled_ref.square().sum() + mean_ref.sum()).backward()
    gi = G.get_graph_index(ei, batch, ng)
    xd = x.clone().requires_grad_(True)
    ad = att.clone().requires_grad_(True) if with_att else None
    out = G.ops.gin_aggregate(xd, ad, gi, 0.0)
    pooled, mean = G.ops.global_add_pool(out, gi), G.ops.global_mean_pool(out, gi)
    ((out * w).sum() + pooled.square().sum() + mean.sum()).backward()
    assert close(out, ref.detach()) and close(pooled, pooled_ref.detach()) and close(mean, mean_ref.detach())
    assert close(xd.grad, xr.grad, rtol=2e-5, atol=4e-6)
    if with_att and E:
        assert close(ad.grad, ar.grad, rtol=2e-5, atol=4e-6)


@settings(**SETTINGS)
@given(batches(), st.sampled_from([4, 12, 64, 100]), st.integers(0, 2 ** 31 - 1))
def test_instance_norm_over_ragged_segments(G, case, C, seed):
    _, batch, ng = case
    N = batch.numel()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, generator=g) * 3 + 1
    w = torch.randn(N, C, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = O.InstanceNorm(C)(xr, batch, ng)
    (ref * w).sum().backward()
    gi = G.get_graph_index(torch.zeros((2, 0), dtype=torch.int64), batch, ng)
    xd = x.clone().requires_grad_(True)
    out = G.InstanceNorm(C)(xd, batch, seg=(gi.node_ptr, gi.G))
    (out * w).sum().backward()
    assert close(out, ref.detach(), rtol=2e-5, atol=2e-5)
    assert close(xd.grad, xr.grad, rtol=1e-4, atol=1e-4)


@settings(**SETTINGS)
@given(batches(symmetric=True), st.booleans(), st.booleans(), st.integers(0, 2 ** 31 - 1))
def test_sampler_average_info_and_lift(G, case, training, info_on_edge_att, seed):
    ei, batch, ng = case
    # coalesce: the reverse map (reorder_like) is defined for duplicate-free edge lists only
    N = batch.numel()
    key = torch.unique(ei[0] * N + ei[1])
    ei = torch.stack([key // N, key % N])
    ei = ei[:, torch.argsort(batch[ei[0]], stable=True)]
    E = ei.shape[1]
    if E == 0:
        return
    g = torch.Generator().manual_seed(seed)
    logit = torch.randn(E, 1, generator=g) * 2
    u = torch.rand(E, 1, generator=g).clamp(1e-10, 1 - 1e-10)
    w = torch.randn(E, 1, generator=g)
    lr = logit.clone().requires_grad_(True)
    att = O.concrete_sample(lr, 1, training, u)
    ea = O.undirected_average(att, ei)
    il = O.info_loss(ea if info_on_edge_att else att, 0.7)
    ((ea * w).sum() + 3.0 * il).backward()
    gi = G.get_graph_index(ei, batch, ng)
    assert gi.symmetric
    ld = logit.clone().requires_grad_(True)
    att_d, ea_d, il_d = G.ops.sample_avg_info(ld, training=training, rev=gi.rev, average=True, r=0.7, noise_u=u,
                                               info_on_edge_att=info_on_edge_att)
    ((ea_d * w).sum() + 3.0 * il_d).backward()
    assert close(att_d, att.detach()) and close(ea_d, ea.detach()) and close(il_d, il.detach())
    assert close(ld.grad, lr.grad, rtol=2e-5, atol=4e-6)
    node_att = torch.rand(N, 1, generator=g)
    nr, nd = node_att.clone().requires_grad_(True), node_att.clone().requires_grad_(True)
    lift_ref = O.lift_node_att_to_edge_att(nr, ei)
    (lift_ref * w).sum().backward()
    lift = G.ops.lift_node_att(nd, gi)
    (lift * w).sum().backward()
    assert close(lift, lift_ref.detach()) and close(nd.grad, nr.grad, rtol=2e-5, atol=4e-6)


@settings(**SETTINGS)
@given(batches(allow_empty_graphs=False), st.booleans())
def test_line_graph_dual_matches_the_reference_loops(G, case, halve):
    ei, batch, ng = case
    if halve:                       # the halved relabelling pairs consecutive rows: needs both directions back to back
        und = ei[:, ei[0] < ei[1]]
        ei = torch.stack([und, und.flip(0)], dim=2).reshape(2, -1)
        ei = ei[:, torch.argsort(batch[ei[0]].repeat_interleave(1), stable=True)] if ei.shape[1] else ei
        if ei.shape[1] % 2:
            return
    exp_ei, exp_b = O.line_graph_dual(ei, batch, halve=halve)
    got_ei, got_b = G.line_graph_dual(ei, batch, halve=halve)
    assert torch.equal(got_ei, exp_ei) and torch.equal(got_b, exp_b)


@settings(**SETTINGS)
@given(batches(allow_empty_graphs=False), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_precision_at_k(G, case, k, seed):
    ei, batch, ng = case
    ei = ei[:, torch.argsort(batch[ei[0]], stable=True)] if ei.shape[1] else ei
    E = ei.shape[1]
    g = torch.Generator().manual_seed(seed)
    att = (torch.randint(0, 6, (E,), generator=g).float() / 5).view(-1, 1)            # many exact ties
    labels = (torch.rand(E, generator=g) > 0.5).float()
    counts = torch.bincount(batch[ei[0]], minlength=ng) if E else torch.zeros(ng, dtype=torch.long)
    if E == 0 or int(counts.min()) == 0:
        return                       # the reference loop divides by zero for a graph without edges
    exp = O.get_precision_at_k(att, labels, k, batch, ei)
    got = G.get_precision_at_k(att, labels, k, batch, ei, ng)
    assert torch.allclose(got.double(), torch.tensor(exp, dtype=torch.float64), rtol=0, atol=1e-6)
