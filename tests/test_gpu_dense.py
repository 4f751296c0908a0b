"""GPU suite: the dense layers (dp_gsat_b200/dense.py, csrc/dense.cu) -- nn.Linear / BatchNorm1d of the reference
(src/models/gin.py:22-25,42,55-62; src/models/pna.py:20-50; src/utils/get_model.py:57-68) on this library's tcgen05
GEMM kernels, forward and backward, in both precision modes.

  precision 'fp32' (split-bf16 x3 strict mode)   rtol 1e-5 against the fp32 arithmetic of the oracle's torch CPU layers
                                                  (and at least as close to fp64 as 4x the fp32 CPU result)
  precision 'bf16'                                3e-2 of max (documented bf16 bound)
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def G():
    import dp_gsat_b200 as g
    return g


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


SHAPES = [(300, 64, 64), (1000, 128, 128), (2571, 640, 80), (513, 10, 64), (50, 80, 1), (4099, 512, 128), (777, 600, 1200),
          (129, 3, 3)]


@pytest.mark.parametrize('rows,K,OUT', SHAPES)
def test_strict_linear_matches_fp32(G, rows, K, OUT):
    from dp_gsat_b200 import dense
    g = torch.Generator().manual_seed(rows + K)
    x = torch.randn(rows, K, generator=g)
    w = torch.randn(OUT, K, generator=g) / K ** 0.5
    b = torch.randn(OUT, generator=g)
    dy = torch.randn(rows, OUT, generator=g)
    ref = [t.clone().requires_grad_(True) for t in (x, w, b)]
    ref64 = [t.double().requires_grad_(True) for t in (x, w, b)]
    y32 = torch.nn.functional.linear(*ref)
    y32.backward(dy)
    y64 = torch.nn.functional.linear(*ref64)
    y64.backward(dy.double())
    got = [t.clone().cuda().requires_grad_(True) for t in (x, w, b)]
    y = dense.linear(got[0], got[1], got[2], 'fp32')
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    for name, a, r32, r64 in [('y', y, y32, y64), ('dx', got[0].grad, ref[0].grad, ref64[0].grad),
                              ('dW', got[1].grad, ref[1].grad, ref64[1].grad), ('db', got[2].grad, ref[2].grad, ref64[2].grad)]:
        scale = float(r32.abs().max())
        assert torch.allclose(a.cpu(), r32, rtol=1e-5, atol=1e-5 * scale), \
            f'{name}: max abs err {float((a.cpu() - r32).abs().max()):.3e} at scale {scale:.3e}'
        e_got, e_32 = _rel(a, r64), _rel(r32, r64)
        assert e_got <= 4 * e_32 + 1e-7, f'{name}: rel L2 vs fp64 {e_got:.3e} (fp32 CPU: {e_32:.3e})'


@pytest.mark.parametrize('rows,K,OUT', [(1000, 128, 128), (2571, 640, 80), (513, 10, 64)])
def test_bf16_linear_within_documented_bound(G, rows, K, OUT):
    from dp_gsat_b200 import dense
    g = torch.Generator().manual_seed(rows)
    x, w, b = torch.randn(rows, K, generator=g), torch.randn(OUT, K, generator=g) / K ** 0.5, torch.randn(OUT, generator=g)
    dy = torch.randn(rows, OUT, generator=g)
    ref = [t.clone().requires_grad_(True) for t in (x, w, b)]
    torch.nn.functional.linear(*ref).backward(dy)
    got = [t.clone().cuda().requires_grad_(True) for t in (x, w, b)]
    y = dense.linear(got[0], got[1], got[2], 'bf16')
    y.backward(dy.cuda())
    # same operand rounding, fp32 accumulate: tight
    same = x.bfloat16().float() @ w.bfloat16().float().t() + b
    assert torch.allclose(y.cpu(), same, rtol=1e-4, atol=1e-4)
    for a, r in zip([got[0].grad, got[1].grad, got[2].grad], [ref[0].grad, ref[1].grad, ref[2].grad]):
        assert float((a.cpu() - r).abs().max()) <= 3e-2 * float(r.abs().max())


@pytest.mark.parametrize('rows,K,OUT', [(1000, 960, 80), (513, 640, 80)])
def test_bf16x2_linear(G, rows, K, OUT):
    """'bf16x2' (two bf16 parts per operand, three K-segments; PNA post_nn in precision 'bf16'): 1e-4 of max."""
    from dp_gsat_b200 import dense
    g = torch.Generator().manual_seed(K)
    x, w, b = torch.randn(rows, K, generator=g), torch.randn(OUT, K, generator=g) / K ** 0.5, torch.randn(OUT, generator=g)
    dy = torch.randn(rows, OUT, generator=g)
    ref = [t.clone().requires_grad_(True) for t in (x, w, b)]
    y_r = torch.nn.functional.linear(*ref)
    y_r.backward(dy)
    got = [t.clone().cuda().requires_grad_(True) for t in (x, w, b)]
    y = dense.linear(got[0], got[1], got[2], 'bf16x2')
    y.backward(dy.cuda())
    for a, r in zip([y, got[0].grad, got[1].grad, got[2].grad], [y_r, ref[0].grad, ref[1].grad, ref[2].grad]):
        assert float((a.detach().cpu() - r.detach()).abs().max()) <= 1e-4 * float(r.detach().abs().max())


@pytest.mark.parametrize('rows,C,relu', [(1000, 80, False), (5000, 128, True), (37, 64, False), (2, 16, True)])
def test_batch_norm_matches_torch(G, rows, C, relu):
    from dp_gsat_b200 import dense
    g = torch.Generator().manual_seed(C)
    x = torch.randn(rows, C, generator=g) * 3 + 5
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    dy = torch.randn(rows, C, generator=g)
    for training in (True, False):
        rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
        ref = [t.clone().requires_grad_(True) for t in (x, gamma, beta)]
        rm_r, rv_r = rm.clone(), rv.clone()
        y_r = torch.nn.functional.batch_norm(ref[0], rm_r, rv_r, ref[1], ref[2], training, 0.1, 1e-5)
        if relu:
            y_r = torch.relu(y_r)
        y_r.backward(dy)
        got = [t.clone().cuda().requires_grad_(True) for t in (x, gamma, beta)]
        rm_g, rv_g = rm.clone().cuda(), rv.clone().cuda()
        y = dense.batch_norm(got[0], got[1], got[2], rm_g, rv_g, training, 0.1, 1e-5, relu)
        y.backward(dy.cuda())
        torch.cuda.synchronize()
        # rows = 2: the two-row batch variance is as ill-conditioned as a statistic gets -- absolute bound there
        rt = 1e-5 if rows > 2 else 1e-3
        assert torch.allclose(y.cpu(), y_r, rtol=rt, atol=1e-5)
        for a, r in zip([t.grad for t in got], [t.grad for t in ref]):
            assert torch.allclose(a.cpu(), r, rtol=10 * rt, atol=1e-5 * max(1.0, float(r.abs().max()))), \
                float((a.cpu() - r).abs().max())
        assert torch.allclose(rm_g.cpu(), rm_r, rtol=1e-5, atol=1e-6)
        assert torch.allclose(rv_g.cpu(), rv_r, rtol=1e-5, atol=1e-6)


def test_no_library_gemm_in_a_strict_step(G):
    """A whole strict-mode GSAT-GIN training step launches none of the library GEMM / batch-norm kernels: every kernel
    name the profiler sees is either one of this library's or an ATen elementwise / reduction / copy kernel."""
    from torch.profiler import profile, ProfilerActivity
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(32, seed=0).to('cuda')
    cfg = {'model_name': 'GIN', 'hidden_size': 64, 'n_layers': 2, 'dropout_p': 0.3, 'use_edge_attr': False}
    torch.manual_seed(0)
    clf = G.get_model(10, 0, 2, False, cfg, 'cuda')
    ext = G.ExtractorMLP(64, {'learn_edge_att': True, 'extractor_dropout_p': 0.5}).cuda()
    gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5)
    gsat.train()
    for precision in ('fp32', 'bf16'):
        clf.precision = ext.precision = precision
        gsat.forward_pass(b, 0, True)[1].backward()      # warm-up (index build, lazy module loads)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            gsat.forward_pass(b, 0, True)[1].backward()
            torch.cuda.synchronize()
        names = {e.key for e in prof.key_averages()}
        ours = ('k_tc_gemm', 'k_ext_fused', 'k_tc_dw')
        banned = [n for n in names if not any(o in n for o in ours) and
                  any(t in n.lower() for t in ('gemm', 'nvjet', 'cutlass', 'cublas', 'batch_norm', 'splitkreduce', 'gemv'))]
        assert not banned, (precision, banned)
        assert any('k_tc_gemm' in n or 'k_ext_fused' in n for n in names), names
