"""GPU probe for the fused extractor kernels (gsatb_ext_fused_fwd / _bwd): correctness against the same-rounding torch
restatement (tests/helpers/ext_ref.py) on small ragged batches, then timing at the cfg4 shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200._lib import lib, ptr, stream
from dp_gsat_b200 import tc
from tests.helpers.ext_ref import extractor_forward

L = lib()
dev = 'cuda'


def fwd(emb, gi, w1, w2, w3, b3, edge, m1=None, m2=None, pdrop=0.0, training=0, seed=3, want_x=True):
    H, C1 = emb.shape[1], w1.shape[0]
    ms = int(L.cdll.gsatb_ext_tile_slots(H, int(edge)))
    plan = gi.ext_plan('edge' if edge else 'node', ms)
    rows, T = plan['rows'], plan['T']
    logit = torch.empty(rows, device=dev)
    xh2t = torch.zeros(T, (H + 127) // 128 * 128, 128, dtype=torch.bfloat16, device=dev) if want_x else None      # tile-major
    seeds = torch.zeros(2, dtype=torch.int32, device=dev)
    w1p, w2p = tc.prep_weight(w1), tc.prep_weight(w2)
    L.call('gsatb_ext_fused_fwd', ptr(emb), ptr(gi.src) if edge else None, ptr(gi.dst) if edge else None,
           ptr(gi.node_ptr) if edge else None, ptr(gi.rowptr_src) if edge else None, ptr(gi.rowptr_dst) if edge else None,
           ptr(plan['seg_ptr']), ptr(plan['tile_seg']), ptr(plan['out2']), max(gi.G, 1), ms, ptr(w1p), ptr(w2p), ptr(w3),
           ptr(b3), ptr(m1), ptr(m2), ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training), ptr(logit), ptr(xh2t),
           T * 128, None, None, ptr(seeds), rows, H, C1, ctypes.c_float(1e-5), stream())
    return logit, xh2t


def check(H, n_graphs, edge=True, masks=False, gen='molhiv'):
    from dp_gsat_b200.data import molhiv_like_batch, ba2motifs_batch
    b = (molhiv_like_batch(n_graphs, seed=H) if gen == 'molhiv' else ba2motifs_batch(n_graphs, seed=H)).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    g = torch.Generator().manual_seed(H)
    N = b.num_nodes
    emb = (torch.randn(N, H, generator=g) + 0.5).to(dev)
    Kin, C1 = (2 * H, 4 * H) if edge else (H, 2 * H)
    w1 = (torch.randn(C1, Kin, generator=g) / Kin ** 0.5).to(dev)
    w2 = (torch.randn(H, C1, generator=g) / C1 ** 0.5).to(dev)
    w3 = (torch.randn(H, generator=g) / H ** 0.5).to(dev)
    b3 = torch.randn(1, generator=g).to(dev)
    rows = gi.E if edge else gi.N
    ms_ = int(L.cdll.gsatb_ext_tile_slots(H, int(edge)))
    if gi.ext_plan('edge' if edge else 'node', ms_)['oversize']:
        print(f'H={H} graphs={n_graphs}: skipped (a graph exceeds {ms_} slots)')
        return True
    pd = 0.5
    m1 = (torch.rand(rows, C1, generator=g) > pd).to(torch.uint8).to(dev) if masks else None
    m2 = (torch.rand(rows, H, generator=g) > pd).to(torch.uint8).to(dev) if masks else None
    logit, _ = fwd(emb, gi, w1, w2, w3, b3, edge, m1, m2, pd if masks else 0.0, 1 if masks else 0)
    torch.cuda.synchronize()
    seg = (gi.edge_ptr if edge else gi.node_ptr).long()
    ref = extractor_forward(emb, gi.src.long() if edge else None, gi.dst.long() if edge else None, seg, w1, w2, w3, b3,
                            m1, m2, pd if masks else 0.0, rounding='bf16').view(-1)
    ref64 = extractor_forward(emb.double(), gi.src.long() if edge else None, gi.dst.long() if edge else None, seg,
                              w1.double(), w2.double(), w3.double(), b3.double(), m1, m2, pd if masks else 0.0).view(-1)
    e = ((logit - ref).abs().max() / ref.abs().max()).item()
    e64 = ((logit.double() - ref64).norm() / ref64.norm()).item()
    ok = e < 5e-3
    print(f'H={H} graphs={n_graphs} rows={rows} edge={edge} masks={masks}: max rel err vs same-rounding ref {e:.2e}, '
          f'rel L2 vs fp64 reference {e64:.2e} {"OK" if ok else "FAIL"}', flush=True)
    return ok


def timing(n_graphs=196000, H=128):
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(n_graphs, seed=0).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    emb = torch.randn(b.num_nodes, H, device=dev)
    w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
    w2 = torch.randn(H, 4 * H, device=dev) / 22
    w3 = torch.randn(H, device=dev) / 11
    b3 = torch.zeros(1, device=dev)
    plan = gi.ext_plan('edge', int(L.cdll.gsatb_ext_tile_slots(H, 1)))
    print(f'E={gi.E} tiles={plan["T"]} slots/tile={gi.E / plan["T"]:.1f}', flush=True)
    for tr in (0, 1):
        for _ in range(3):
            fwd(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, tr)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fwd(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, tr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = gi.E * (24.0 * H * H + 2 * H)
        print(f'fused extractor fwd (training={tr}) E={gi.E} H={H}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s', flush=True)


if __name__ == '__main__' and not (len(sys.argv) > 1 and sys.argv[1] in ('roles', 'bwd', 'roles_bwd')):
    ok = True
    ok &= check(64, 40)
    ok &= check(64, 40, masks=True)
    ok &= check(128, 300, gen='ba')
    ok &= check(128, 300, masks=True, gen='ba')
    ok &= check(16, 60)
    ok &= check(80, 30, masks=True)
    ok &= check(64, 50, edge=False)
    ok &= check(128, 50, edge=False, masks=True)
    print('ALL OK' if ok else 'SOME FAILED', flush=True)
    if ok and len(sys.argv) > 1 and sys.argv[1] == 'time':
        timing()
    sys.exit(0 if ok else 1)


def roles(n_graphs=40000, H=128):
    """Per-role cycle counters of gsatb_ext_fused_fwd (development aid): where each role of the CTA spends its time."""
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(n_graphs, seed=0).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    emb = torch.randn(b.num_nodes, H, device=dev)
    w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
    w2 = torch.randn(H, 4 * H, device=dev) / 22
    w3 = torch.randn(H, device=dev) / 11
    b3 = torch.zeros(1, device=dev)
    plan = gi.ext_plan('edge', int(L.cdll.gsatb_ext_tile_slots(H, 1)))
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    fwd(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, 1)
    L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    fwd(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, 1)
    torch.cuda.synchronize()
    L.cdll.gsatb_tc_set_profile_buffer(None)
    d = buf.view(148, 16).double()
    tiles = plan['T'] / 148.0
    names = ['mma_total', 'mma_wait_x', 'mma_wait_acc1_empty', 'mma_wait_w', 'mma_wait_h1', 'mma_wait_acc2_empty',
             'epiA_wait_acc1', 'epiA_work1', 'epiA_wait_acc2', 'epiA_work2', 'epiB_wait_acc1', 'epiB_work1',
             'epiB_wait_acc2', 'epiB_work2', 'pro_wait_x_empty', 'pro_work']
    print(f'roles: tiles/CTA {tiles:.1f}; cycles per tile (mean over CTAs)')
    for i, n in enumerate(names):
        print(f'  {n:22s} {d[:, i].mean().item() / tiles:10.0f}')


if len(sys.argv) > 1 and sys.argv[1] == 'roles':
    roles()


def bwd(emb, gi, w1, w2, w3, dlogit, xh2t, rstd2, seeds_xs, edge, m1=None, m2=None, pdrop=0.0, training=0):
    seeds, xs = seeds_xs
    H, C1 = emb.shape[1], w1.shape[0]
    Kin = 2 * H if edge else H
    ms = int(L.cdll.gsatb_ext_tile_slots(H, int(edge)))
    plan = gi.ext_plan('edge' if edge else 'node', ms)
    rows, T = plan['rows'], plan['T']
    ld = T * 128
    bf = dict(dtype=torch.bfloat16, device=dev)
    HP, C1P = (H + 127) // 128 * 128, (C1 + 127) // 128 * 128
    dz2t, dz1t, h1t = torch.empty(T, HP, 128, **bf), torch.empty(T, C1P, 128, **bf), torch.empty(T, C1P, 128, **bf)
    df12 = torch.empty(rows, Kin, device=dev)
    dw3p = torch.zeros(min(max(gi.G, 1), 148) * 2, H, device=dev)
    w1p, w2t, w1t = tc.prep_weight(w1), tc.prep_weight(w2, transpose=True), tc.prep_weight(w1, transpose=True)
    L.call('gsatb_ext_fused_bwd', ptr(plan['seg_ptr']), ptr(plan['tile_seg']), ptr(plan['out2']), max(gi.G, 1), ms, int(edge),
           ptr(w1p), ptr(w2t), ptr(w1t), ptr(w3), ptr(dlogit), ptr(xh2t), ptr(rstd2), ptr(xs), ptr(m1), ptr(m2), ptr(seeds),
           ctypes.c_float(pdrop), int(training), ptr(dz2t), ptr(dz1t), ptr(h1t), ptr(df12), 0, ptr(dw3p), ld, rows, H, C1,
           ctypes.c_float(1e-5), stream())
    return dz2t, dz1t, h1t, xs, df12, dw3p


_XS = {}


def _xs_buffer(rows, ldx):
    """zero-filled once, re-used: rows [128 t + max_slots, 128 t + 128) are never written by the kernels"""
    key = (rows, ldx)
    if key not in _XS:
        _XS.clear()
        _XS[key] = torch.zeros(rows, ldx, dtype=torch.bfloat16, device=dev)
    return _XS[key]


def fwd_full(emb, gi, w1, w2, w3, b3, edge, m1, m2, pdrop, training, seed=3):
    H, C1 = emb.shape[1], w1.shape[0]
    ms = int(L.cdll.gsatb_ext_tile_slots(H, int(edge)))
    plan = gi.ext_plan('edge' if edge else 'node', ms)
    rows, T = plan['rows'], plan['T']
    logit = torch.empty(rows, device=dev)
    xh2t = torch.zeros(T, (H + 127) // 128 * 128, 128, dtype=torch.bfloat16, device=dev)
    rstd2 = torch.empty(max(gi.G, 1), H, device=dev)
    seeds = torch.zeros(2, dtype=torch.int32, device=dev)
    Kin = 2 * H if edge else H
    xs = _xs_buffer(T * 128, (Kin + 63) // 64 * 64)
    w1p, w2p = tc.prep_weight(w1), tc.prep_weight(w2)
    L.call('gsatb_ext_fused_fwd', ptr(emb), ptr(gi.src) if edge else None, ptr(gi.dst) if edge else None,
           ptr(gi.node_ptr) if edge else None, ptr(gi.rowptr_src) if edge else None, ptr(gi.rowptr_dst) if edge else None,
           ptr(plan['seg_ptr']), ptr(plan['tile_seg']), ptr(plan['out2']), max(gi.G, 1), ms, ptr(w1p), ptr(w2p), ptr(w3),
           ptr(b3), ptr(m1), ptr(m2), ctypes.c_uint64(seed), ctypes.c_float(pdrop), int(training), ptr(logit), ptr(xh2t),
           T * 128, ptr(rstd2), ptr(xs), ptr(seeds), rows, H, C1, ctypes.c_float(1e-5), stream())
    return logit, xh2t, rstd2, (seeds, xs)


def check_bwd(H, n_graphs, edge=True, masks=False, gen='ba'):
    from dp_gsat_b200.data import molhiv_like_batch, ba2motifs_batch
    from tests.helpers.ext_ref import extractor_backward_emulated
    b = (molhiv_like_batch(n_graphs, seed=H) if gen == 'molhiv' else ba2motifs_batch(n_graphs, seed=H)).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    g = torch.Generator().manual_seed(H)
    emb = (torch.randn(b.num_nodes, H, generator=g) + 0.5).to(dev)
    Kin, C1 = (2 * H, 4 * H) if edge else (H, 2 * H)
    w1 = (torch.randn(C1, Kin, generator=g) / Kin ** 0.5).to(dev)
    w2 = (torch.randn(H, C1, generator=g) / C1 ** 0.5).to(dev)
    w3 = (torch.randn(H, generator=g) / H ** 0.5).to(dev)
    b3 = torch.randn(1, generator=g).to(dev)
    rows = gi.E if edge else gi.N
    ms = int(L.cdll.gsatb_ext_tile_slots(H, int(edge)))
    if gi.ext_plan('edge' if edge else 'node', ms)['oversize']:
        print(f'H={H} graphs={n_graphs}: skipped (a graph exceeds {ms} slots)')
        return True
    pd = 0.5
    m1 = (torch.rand(rows, C1, generator=g) > pd).to(torch.uint8).to(dev) if masks else None
    m2 = (torch.rand(rows, H, generator=g) > pd).to(torch.uint8).to(dev) if masks else None
    dlogit = torch.randn(rows, generator=g).to(dev)
    logit, xh2t, rstd2, seeds = fwd_full(emb, gi, w1, w2, w3, b3, edge, m1, m2, pd if masks else 0.0, 1 if masks else 0)
    dz2t, dz1t, h1t, xs, df12, dw3p = bwd(emb, gi, w1, w2, w3, dlogit, xh2t, rstd2, seeds, edge, m1, m2,
                                          pd if masks else 0.0, 1 if masks else 0)
    torch.cuda.synchronize()
    seg = (gi.edge_ptr if edge else gi.node_ptr).long()
    src, dst = (gi.src.long(), gi.dst.long()) if edge else (None, None)
    e_df, e_dW1, e_dW2, e_dw3 = extractor_backward_emulated(emb, src, dst, seg, w1, w2, w3, dlogit, m1, m2, pd if masks else 0.0)
    rel = lambda a, c: ((a.double() - c.double()).norm() / c.double().norm().clamp_min(1e-30)).item()
    cm = lambda t, C: t.float().permute(1, 0, 2).reshape(t.shape[1], -1)[:C]      # tile-major -> [C, slots]
    dW2 = cm(dz2t, H) @ cm(h1t, C1).t()
    dW1 = cm(dz1t, C1) @ xs.float()[:, :Kin]
    errs = [rel(df12, e_df), rel(dW1, e_dW1), rel(dW2, e_dW2), rel(dw3p.sum(0), e_dw3)]
    ok = all(e < 5e-3 for e in errs)
    print(f'bwd H={H} graphs={n_graphs} rows={rows} edge={edge} masks={masks}: vs same-rounding emulation df12 {errs[0]:.2e} '
          f'dW1 {errs[1]:.2e} dW2 {errs[2]:.2e} dw3 {errs[3]:.2e} {"OK" if ok else "FAIL"}', flush=True)
    return ok


def timing_bwd(n_graphs=196000, H=128):
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(n_graphs, seed=0).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    emb = torch.randn(b.num_nodes, H, device=dev)
    w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
    w2 = torch.randn(H, 4 * H, device=dev) / 22
    w3 = torch.randn(H, device=dev) / 11
    b3 = torch.zeros(1, device=dev)
    dlogit = torch.randn(gi.E, device=dev)
    logit, xh2t, rstd2, seeds = fwd_full(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, 1)
    for _ in range(2):
        out = bwd(emb, gi, w1, w2, w3, dlogit, xh2t, rstd2, seeds, True, None, None, 0.5, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = bwd(emb, gi, w1, w2, w3, dlogit, xh2t, rstd2, seeds, True, None, None, 0.5, 1)
    e1.record()
    torch.cuda.synchronize()
    print(f'fused extractor bwd E={gi.E} H={H}: {e0.elapsed_time(e1) / 3:.3f} ms', flush=True)


if len(sys.argv) > 1 and sys.argv[1] == 'bwd':
    ok = True
    ok &= check_bwd(64, 40)
    ok &= check_bwd(64, 40, masks=True)
    ok &= check_bwd(128, 300)
    ok &= check_bwd(128, 300, masks=True)
    ok &= check_bwd(16, 60, gen='molhiv')
    ok &= check_bwd(80, 30, masks=True, gen='molhiv')
    ok &= check_bwd(64, 50, edge=False)
    ok &= check_bwd(128, 50, edge=False, masks=True)
    print('BWD ALL OK' if ok else 'BWD SOME FAILED', flush=True)
    if ok:
        timing_bwd()


def roles_bwd(n_graphs=40000, H=128):
    from dp_gsat_b200.data import ba2motifs_batch
    b = ba2motifs_batch(n_graphs, seed=0).to(dev)
    gi = G.get_graph_index(b.edge_index, b.batch)
    emb = torch.randn(b.num_nodes, H, device=dev)
    w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
    w2 = torch.randn(H, 4 * H, device=dev) / 22
    w3 = torch.randn(H, device=dev) / 11
    b3 = torch.zeros(1, device=dev)
    dlogit = torch.randn(gi.E, device=dev)
    logit, xh2t, rstd2, seeds = fwd_full(emb, gi, w1, w2, w3, b3, True, None, None, 0.5, 1)
    bwd(emb, gi, w1, w2, w3, dlogit, xh2t, rstd2, seeds, True, None, None, 0.5, 1)
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    bwd(emb, gi, w1, w2, w3, dlogit, xh2t, rstd2, seeds, True, None, None, 0.5, 1)
    torch.cuda.synchronize()
    L.cdll.gsatb_tc_set_profile_buffer(None)
    d = buf.view(148, 16).double()
    tiles = gi.ext_plan('edge', int(L.cdll.gsatb_ext_tile_slots(H, 1)))['T'] / 148.0
    names = ['mma_total', 'mma_wait_inputs', 'mma_wait_acc_empty', 'mma_wait_w', 'mma_wait_dz1', 'mma_wait_dx_empty',
             'epi_wait_head', 'epi_head', 'epi_wait_cb', 'epi_cb(all 4)', 'epi_wait_dx', 'epi_dxout']
    print(f'bwd roles: tiles/CTA {tiles:.1f}; cycles per tile (mean over CTAs)')
    for i, nme in enumerate(names):
        print(f'  {nme:22s} {d[:, i].mean().item() / tiles:10.0f}')


if len(sys.argv) > 1 and sys.argv[1] == 'roles_bwd':
    roles_bwd()
