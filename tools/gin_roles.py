"""Per-role cycle counters (tcgen05 skeleton, gsatb_tc_set_profile_buffer) of the GIN node-MLP kernels at the cfg4 shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dp_gsat_b200 as G
from dp_gsat_b200 import tc
from dp_gsat_b200._lib import lib, ptr, stream

dev = 'cuda'
N, H = int(sys.argv[1]) if len(sys.argv) > 1 else 4900000, 128
L = lib()
torch.manual_seed(0)
names = ['mma_total', 'mma_wait_Bfull', 'mma_wait_accfree', 'mma_wait_W', 'epi_wait', 'epi_work', 'pro_wait', 'pro_fill', 'tiles']
x16 = torch.randn(N, H, device=dev).bfloat16()
w = torch.randn(H, H, device=dev) / 11
wp = tc.prep_weight(w)
bias = torch.zeros(H, device=dev)


def prof(name, fn):
    for _ in range(2):
        fn()
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    L.cdll.gsatb_tc_set_profile_buffer(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.cdll.gsatb_tc_set_profile_buffer(None)
    d = dbg.double().mean(0).cpu()
    tiles = N / 128 / 148
    print(f'{name}: {e0.elapsed_time(e1):.3f} ms; cycles per tile: ' + ', '.join(f'{n}={d[i].item() / tiles:.0f}' for i, n in enumerate(names[:8])), flush=True)


prof('linear_bf16 -> bf16 + stats', lambda: tc.linear_bf16(x16, wp, bias, H, want_stats=True))
prof('linear_bf16 -> fp32 relu dropout posmask', lambda: tc.linear_bf16(x16, wp, bias, H, out_bf16=False, relu_out=True, pdrop=0.3, drop_seed=1,
                                                                     posmask=torch.empty((N, 4), dtype=torch.int32, device=dev)))
prof('linear_bf16 -> fp32 plain', lambda: tc.linear_bf16(x16, wp, bias, H, out_bf16=False))
# backward kernels through the module path
seq = G.GIN.MLP(H, H).to(dev)
xin = torch.randn(N, H, device=dev, requires_grad=True)
out = tc.gin_mlp_relu(xin, seq, True, 0.3, 1)
g = torch.randn(N, H, device=dev)
prof('gin_mlp backward (bwd2 + bwd1 + dW x2; counters = last skeleton kernel)', lambda: torch.autograd.grad(out, [xin] + list(seq.parameters()), g, retain_graph=True))
