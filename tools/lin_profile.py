import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dp_gsat_b200 import tc
from dp_gsat_b200._lib import lib, ptr, stream
dev = 'cuda'
L = lib()
rows = 1_020_000
xl = torch.randn(rows, 128, device=dev)
wl = tc.prep_weight(torch.randn(128, 128, device=dev) / 11)
ol = torch.empty(rows, 128, device=dev)
for _ in range(3):
    L.call('gsatb_tc_linear_fwd', ptr(xl), 0, 128, None, None, ptr(wl), None, ptr(ol), 128, 0, None, None, None, ctypes.c_uint64(0), ctypes.c_float(0.0), rows, 128, 128, stream())
torch.cuda.synchronize()
print('ok')
