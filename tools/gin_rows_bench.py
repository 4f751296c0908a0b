"""Row-owner node-MLP kernels (csrc/gin_rows.cu) against the channel-owner skeleton kernels they replace, at the cfg4
layer shape (4.9 M rows, H = 128) and at H = 64: CUDA-event times, fraction of the measured HBM copy peak for the
algorithmic bytes (every distinct input read once, every output written once)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dp_gsat_b200 import tc
from dp_gsat_b200._lib import lib, ptr, stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4900000
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6537.6) \
    if os.path.exists('MEASURED_PEAKS.json') else 6537.6
L = lib()
dev = 'cuda'


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def line(name, ms, gbytes):
    print(f'  {name:46s} {ms:7.3f} ms  {gbytes / ms:7.0f} GB/s  {gbytes / ms / PEAK:5.2f} of {PEAK:.0f}', flush=True)


for H in (128, 64):
    torch.manual_seed(0)
    print(f'N = {N}, H = {H}  (NST={os.environ.get("GSATB_ROWS_NST", "auto")} NSB={os.environ.get("GSATB_ROWS_NSB", "1")})', flush=True)
    NH = N * H / 1e6      # MB per byte-per-element
    x16 = torch.randn(N, H, device=dev).bfloat16()
    w1, w2 = torch.randn(H, H, device=dev) / H ** 0.5, torch.randn(H, H, device=dev) / H ** 0.5
    b1, b2 = torch.randn(H, device=dev), torch.randn(H, device=dev)
    w1p, w2p, w1t = tc.prep_weight(w1), tc.prep_weight(w2), tc.prep_weight(w1, transpose=True)
    scale, shift = torch.rand(H, device=dev) + 0.5, torch.randn(H, device=dev) * 0.2
    z1, _ = tc.linear_bf16(x16, w1p, b1, H, want_stats=True)
    a1 = torch.empty_like(z1)
    pm = torch.empty((N, H // 32), dtype=torch.int32, device=dev)
    h = torch.empty((N, H), device=dev)
    line('old lin1 (skeleton, bf16 out + stats)', timed(lambda: tc.linear_bf16(x16, w1p, b1, H, want_stats=True)), 4 * NH)
    line('new lin1 (rows)', timed(lambda: tc._rows_lin1(x16, w1p, b1, H, True)), 4 * NH)
    t_bn = timed(lambda: L.call('gsatb_bn_relu_bf16', ptr(z1), ptr(scale), ptr(shift), ptr(a1), N, H, stream()))
    t_l2 = timed(lambda: tc.linear_bf16(a1, w2p, b2, H, out_bf16=False, relu_out=True, pdrop=0.3, drop_seed=3, posmask=pm))
    line('old bn_relu + lin2 (relu, dropout, sign bits)', t_bn + t_l2, 8 * NH + N * H / 8 / 1e6)
    line('new lin2 (rows, BN+ReLU in the operand path)',
         timed(lambda: L.call('gsatb_gin_rows_lin2', ptr(z1), ptr(scale), ptr(shift), ptr(w2p), ptr(b2), ptr(a1), ptr(h), ptr(pm), None,
                              ctypes.c_uint64(3), ctypes.c_float(0.3), N, H, stream())), 8 * NH + N * H / 8 / 1e6)
    g16 = torch.randn(N, H, device=dev).bfloat16()
    cA, cB, cC = torch.randn(H, device=dev), torch.randn(H, device=dev) * 0.1, torch.randn(H, device=dev) * 0.01
    dz, dx = torch.empty_like(g16), torch.empty((N, H), device=dev)
    line('old bwd1 (skeleton, producer layout)',
         timed(lambda: L.call('gsatb_tc_gin_bwd1', ptr(g16), ptr(z1), ptr(cA), ptr(cB), ptr(cC), ptr(w1t), ptr(dz), ptr(dx), N, H, H, stream())), 10 * NH)
    line('new bwd1 (rows)',
         timed(lambda: L.call('gsatb_gin_rows_bwd1', ptr(g16), ptr(z1), ptr(cA), ptr(cB), ptr(cC), ptr(w1t), ptr(dz), ptr(dx), N, H, stream())), 10 * NH)
    dh = torch.randn(N, H, device=dev)
    mean, rstd = torch.randn(H, device=dev) * 0.3, torch.rand(H, device=dev) + 0.5
    w2t = tc.prep_weight(w2, transpose=True)
    d2, gg, stats = torch.empty_like(g16), torch.empty_like(g16), torch.empty(2 * H, device=dev)
    part_o = torch.empty(int(L.cdll.gsatb_tc_stat_partials_elems(H)), device=dev)
    part_n = torch.empty(int(L.cdll.gsatb_gin_rows_stat_partials_elems(H)), device=dev)
    bwd2_old = lambda: L.call('gsatb_tc_gin_bwd2', ptr(dh), None, ptr(pm), ctypes.c_float(1.43), ptr(w2t), ptr(z1), ptr(scale), ptr(shift),
                              ptr(mean), ptr(rstd), ptr(d2), ptr(gg), None, ptr(part_o), ptr(stats), N, H, H, stream())
    bwd2_new = lambda: L.call('gsatb_gin_rows_bwd2', ptr(dh), ptr(pm), ctypes.c_float(1.43), ptr(w2t), ptr(z1), ptr(scale), ptr(shift),
                              ptr(mean), ptr(rstd), ptr(d2), ptr(gg), ptr(part_n), ptr(stats), N, H, stream())
    line('old bwd2 (skeleton, producer layout)', timed(bwd2_old), 10 * NH + N * H / 8 / 1e6)
    line('new bwd2 (rows)', timed(bwd2_new), 10 * NH + N * H / 8 / 1e6)
    # role counters of the new kernels (MMA thread: total / wait input / wait accumulator; epilogue warp 4: wait / work;
    # transform warps: wait / work), cycles per tile of the CTA
    tiles = N / 128 / 148
    for name, fn in (('lin1', lambda: tc._rows_lin1(x16, w1p, b1, H, True)),
                     ('lin2', lambda: L.call('gsatb_gin_rows_lin2', ptr(z1), ptr(scale), ptr(shift), ptr(w2p), ptr(b2), ptr(a1), ptr(h), ptr(pm), None,
                                             ctypes.c_uint64(3), ctypes.c_float(0.3), N, H, stream())),
                     ('bwd2', bwd2_new),
                     ('bwd1', lambda: L.call('gsatb_gin_rows_bwd1', ptr(g16), ptr(z1), ptr(cA), ptr(cB), ptr(cC), ptr(w1t), ptr(dz), ptr(dx), N, H, stream()))):
        dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
        L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(dbg.data_ptr()))
        fn(); torch.cuda.synchronize()
        L.cdll.gsatb_tc_set_profile_buffer(None)
        d = dbg.double().mean(0).cpu() / tiles
        print(f'    roles {name}: mma total {d[0]:.0f}, wait in {d[1]:.0f}, wait acc {d[2]:.0f}; epi(w4, own tiles only) wait {d[4]:.0f} work {d[5]:.0f}; '
              f'xf wait {d[6]:.0f} work {d[7]:.0f}; gate/stat warp wait {d[8]:.0f} work {d[9]:.0f}  cycles per CTA tile', flush=True)
    del x16, z1, a1, h, g16, dz, dx, pm, dh, d2, gg
    torch.cuda.empty_cache()
