"""Launch one row-owner node-MLP kernel at the cfg4 layer shape (for ncu): python tools/gin_rows_one.py bwd2|bwd1|lin1|lin2 [N] [H]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dp_gsat_b200 import tc
from dp_gsat_b200._lib import lib, ptr, stream

which = sys.argv[1] if len(sys.argv) > 1 else 'bwd2'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4900000
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
L, dev = lib(), 'cuda'
torch.manual_seed(0)
x16 = torch.randn(N, H, device=dev).bfloat16()
w = torch.randn(H, H, device=dev) / H ** 0.5
wp, wt = tc.prep_weight(w), tc.prep_weight(w, transpose=True)
b = torch.randn(H, device=dev)
scale, shift, mean, rstd = torch.rand(H, device=dev) + 0.5, torch.randn(H, device=dev) * 0.2, torch.randn(H, device=dev) * 0.3, torch.rand(H, device=dev) + 0.5
o16a, o16b = torch.empty_like(x16), torch.empty_like(x16)
o32 = torch.empty(N, H, device=dev)
pm = torch.randint(-2 ** 31, 2 ** 31 - 1, (N, H // 32), dtype=torch.int64, device=dev).to(torch.int32)
stats = torch.empty(2 * H, device=dev)
part = torch.empty(int(L.cdll.gsatb_gin_rows_stat_partials_elems(H)), device=dev)
for _ in range(2):
    if which == 'bwd2':
        L.call('gsatb_gin_rows_bwd2', ptr(o32), ptr(pm), ctypes.c_float(1.43), ptr(wt), ptr(x16), ptr(scale), ptr(shift), ptr(mean),
               ptr(rstd), ptr(o16a), ptr(o16b), ptr(part), ptr(stats), N, H, stream())
    elif which == 'bwd1':
        L.call('gsatb_gin_rows_bwd1', ptr(x16), ptr(o16a), ptr(scale), ptr(shift), ptr(mean), ptr(wt), ptr(o16b), ptr(o32), N, H, stream())
    elif which == 'lin1':
        tc._rows_lin1(x16, wp, b, H, True)
    else:
        L.call('gsatb_gin_rows_lin2', ptr(x16), ptr(scale), ptr(shift), ptr(wp), ptr(b), ptr(o16a), ptr(o32), ptr(pm), None,
               ctypes.c_uint64(3), ctypes.c_float(0.3), N, H, stream())
torch.cuda.synchronize()
print('done', which)
