"""Per-role cycle counters of the tcgen05 skeleton for the fused extractor kernels (development aid)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import tc
from dp_gsat_b200._lib import lib, ptr, stream
from dp_gsat_b200.data import ba2motifs_batch

dev = 'cuda'
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
b = ba2motifs_batch(ng, seed=0).to(dev)
gi = G.get_graph_index(b.edge_index, b.batch, b.num_graphs)
torch.manual_seed(0)
emb = torch.relu(torch.randn(gi.N, H, device=dev))
w1 = torch.randn(4 * H, 2 * H, device=dev) / 16
w2 = torch.randn(H, 4 * H, device=dev) / 22
w3 = torch.randn(1, H, device=dev) / 11
b3 = torch.zeros(1, device=dev)
L = lib()
names = ['mma_total', 'mma_wait_Bfull', 'mma_wait_accfree', 'mma_wait_W', 'epi_wait', 'epi_work', 'pro_wait', 'pro_fill', 'tiles']
plan = gi.tile_plan('edge')
tile_row, tile_seg, T = plan
C1 = 4 * H
w1p, w2p = tc.prep_weight(w1), tc.prep_weight(w2)
xhat1 = torch.empty((gi.E, C1), dtype=torch.bfloat16, device=dev)
rstd1 = torch.empty((gi.G, C1), device=dev)
xhat2 = torch.empty((gi.E, H), dtype=torch.bfloat16, device=dev)
rstd2 = torch.empty((gi.G, H), device=dev)
logit = torch.empty((gi.E, 1), device=dev)
w3f = w3.reshape(-1).contiguous()
f12buf = (torch.randn((gi.E, 2 * H), device=dev) / 4).bfloat16()
h1buf = torch.zeros((gi.E, C1), dtype=torch.bfloat16, device=dev)


def k1():
    L.call('gsatb_tc_ext_fwd1', ptr(f12buf), ptr(w1p), ptr(tile_row), ptr(tile_seg), ptr(gi.edge_ptr), T,
           ptr(xhat1), ptr(rstd1), gi.E, 2 * H, C1, ctypes.c_float(1e-5), stream())


def k2():
    L.call('gsatb_tc_ext_fwd2', ptr(h1buf), ptr(w2p), ptr(w3f), ptr(b3), None, ctypes.c_uint64(1), ctypes.c_float(0.5), 1,
           ptr(tile_row), ptr(tile_seg), ptr(gi.edge_ptr), T, ptr(xhat2), ptr(rstd2), ptr(logit), gi.E, C1, H, ctypes.c_float(1e-5), stream())


for name, fn in (('ext_fwd1', k1), ('ext_fwd2', k2)):
    for _ in range(2):
        fn()
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.cdll.gsatb_tc_set_profile_buffer(None)
    d = dbg.double().mean(0).cpu()
    tiles = max(d[8].item(), 1)
    print(f'{name}: {e0.elapsed_time(e1):.3f} ms, E={gi.E}, tiles/CTA {tiles:.0f}; cycles per tile: ' +
          ', '.join(f'{n}={d[i].item() / tiles:.0f}' for i, n in enumerate(names[:8])))


def k2_nodrop():
    L.call('gsatb_tc_ext_fwd2', ptr(h1buf), ptr(w2p), ptr(w3f), ptr(b3), None, ctypes.c_uint64(1), ctypes.c_float(0.5), 0,
           ptr(tile_row), ptr(tile_seg), ptr(gi.edge_ptr), T, ptr(xhat2), ptr(rstd2), ptr(logit), gi.E, C1, H, ctypes.c_float(1e-5), stream())


xl = torch.randn(gi.E, 128, device=dev)
wl = tc.prep_weight(torch.randn(128, 128, device=dev) / 11)
ol = torch.empty(gi.E, 128, device=dev)


def klin():
    L.call('gsatb_tc_linear_fwd', ptr(xl), 0, 128, None, None, ptr(wl), None, ptr(ol), 128, 0, None, None, None, ctypes.c_uint64(0), ctypes.c_float(0.0), gi.E, 128, 128, stream())


for name, fn in (('ext_fwd2 (eval, no dropout)', k2_nodrop), ('linear 128x128 on E rows', klin)):
    for _ in range(2):
        fn()
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.cdll.gsatb_tc_set_profile_buffer(None)
    d = dbg.double().mean(0).cpu()
    tiles = max(d[8].item(), 1)
    print(f'{name}: {e0.elapsed_time(e1):.3f} ms, tiles/CTA {tiles:.0f}; cycles per tile: ' +
          ', '.join(f'{n}={d[i].item() / tiles:.0f}' for i, n in enumerate(names[:8])))


dz2 = (torch.randn(gi.E, H, device=dev) / 10).bfloat16()
dz1 = torch.empty(gi.E, C1, dtype=torch.bfloat16, device=dev)
w2t = tc.prep_weight(w2, transpose=True)
w1t = tc.prep_weight(w1, transpose=True)
df = torch.empty(gi.E, 2 * H, device=dev)


def kbwd1():
    L.call('gsatb_tc_ext_bwd1', ptr(dz2), ptr(w2t), ptr(xhat1), ptr(rstd1), None, ctypes.c_uint64(1), ctypes.c_float(0.5), 1,
           ptr(tile_row), ptr(tile_seg), ptr(gi.edge_ptr), T, ptr(dz1), gi.E, H, C1, stream())


def kbf16in():
    L.call('gsatb_tc_linear_bf16in_fwd', ptr(dz1), C1, ptr(w1t), ptr(df), 2 * H, gi.E, C1, 2 * H, stream())


for name, fn in (('ext_bwd1', kbwd1), ('linear_bf16in 512->256', kbf16in)):
    for _ in range(2):
        fn()
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.cdll.gsatb_tc_set_profile_buffer(None)
    d = dbg.double().mean(0).cpu()
    tiles = max(d[8].item(), 1)
    print(f'{name}: {e0.elapsed_time(e1):.3f} ms, B-buffers/CTA {tiles:.0f}; cycles per B buffer: ' +
          ', '.join(f'{n}={d[i].item() / tiles:.0f}' for i, n in enumerate(names[:8])))


x16 = (torch.randn(gi.N, 128, device=dev)).bfloat16()
o32 = torch.empty(gi.N, 128, device=dev)
o16 = torch.empty(gi.N, 128, device=dev, dtype=torch.bfloat16)
pm = torch.empty(gi.N, 4, device=dev, dtype=torch.int32)
bias = torch.zeros(128, device=dev)


def klin16():
    L.call('gsatb_tc_linear_bf16_fwd', ptr(x16), 128, ptr(wl), ptr(bias), ptr(o16), 1, 128, 0, None, None, None, ctypes.c_uint64(0),
           ctypes.c_float(0.0), None, gi.N, 128, 128, stream())


def klin32():
    L.call('gsatb_tc_linear_bf16_fwd', ptr(x16), 128, ptr(wl), ptr(bias), ptr(o32), 0, 128, 1, None, None, None, ctypes.c_uint64(3),
           ctypes.c_float(0.3), ptr(pm), gi.N, 128, 128, stream())


def klin32_plain():
    L.call('gsatb_tc_linear_bf16_fwd', ptr(x16), 128, ptr(wl), ptr(bias), ptr(o32), 0, 128, 0, None, None, None, ctypes.c_uint64(3),
           ctypes.c_float(0.0), None, gi.N, 128, 128, stream())


for name, fn in (('linear_bf16 -> bf16 (Linear1)', klin16), ('linear_bf16 -> fp32 relu+dropout+posmask (Linear2)', klin32),
                 ('linear_bf16 -> fp32 plain', klin32_plain)):
    for _ in range(2):
        fn()
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.cdll.gsatb_tc_set_profile_buffer(None)
    d = dbg.double().mean(0).cpu()
    print(f'{name}: {e0.elapsed_time(e1):.3f} ms on N={gi.N} rows; cycles per CTA: ' +
          ', '.join(f'{n}={d[i].item():.0f}' for i, n in enumerate(names[:6])))
