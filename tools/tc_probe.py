"""GPU probe for the tcgen05 GEMM skeleton: gsatb_tc_linear_fwd vs torch (bf16-rounded operands, fp32 accumulate)."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dp_gsat_b200._lib import lib, ptr, stream

L = lib()
dev = 'cuda'


def prep(w, transpose=False):
    OUT, K = w.shape
    rows, cols = (K, OUT) if transpose else (OUT, K)
    wp = torch.empty(((rows + 127) // 128 * 128, (cols + 63) // 64 * 64), dtype=torch.bfloat16, device=dev)
    L.call('gsatb_tc_prep_weight', ptr(w), OUT, K, int(transpose), ptr(wp), stream())
    return wp


def run(rows, K, OUT, stats=False, scale=False, relu=False):
    g = torch.Generator(device='cpu').manual_seed(rows + K + OUT)
    x = torch.randn(rows, K, generator=g).to(dev)
    w = (torch.randn(OUT, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(OUT, generator=g).to(dev)
    sc = (torch.rand(K, generator=g) + 0.5).to(dev) if scale else None
    shf = torch.randn(K, generator=g).to(dev) if scale else None
    out = torch.empty(rows, OUT, device=dev)
    wp = prep(w)
    part = torch.empty(int(L.cdll.gsatb_tc_stat_partials_elems(OUT)), device=dev) if stats else None
    st = torch.empty(2 * OUT, dtype=torch.float64, device=dev) if stats else None
    L.call('gsatb_tc_linear_fwd', ptr(x), 0, K, ptr(sc), ptr(shf), ptr(wp), ptr(b), ptr(out), OUT, int(relu), ptr(part),
           ptr(st), None, ctypes.c_uint64(0), ctypes.c_float(0.0), rows, K, OUT, stream())
    torch.cuda.synchronize()
    xin = torch.relu(x * sc + shf) if scale else x
    ref = xin.bfloat16().float() @ w.bfloat16().float().t() + b
    err = (out - (torch.relu(ref) if relu else ref)).abs().max().item()
    msg = f'rows={rows} K={K} OUT={OUT} stats={stats} scale={scale} relu={relu}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})'
    if stats:
        e1 = (st[:OUT] - ref.double().sum(0)).abs().max().item() / max(1.0, ref.double().sum(0).abs().max().item())
        e2 = (st[OUT:] - ref.double().square().sum(0)).abs().max().item() / ref.double().square().sum(0).abs().max().item()
        msg += f' stat rel err {e1:.2e} {e2:.2e}'
    print(msg, flush=True)
    return err


if __name__ == '__main__':
    run(128, 64, 64)
    run(1000, 128, 128, stats=True)
    run(5000, 128, 128, scale=True, relu=True)
    run(3333, 256, 512)
    run(2500, 512, 128)
    run(2500, 512, 256)
    run(777, 80, 80)
    run(999, 300, 300)
    # timing at cfg4 node-MLP size
    rows, K, OUT = 4_900_000, 128, 128
    x = torch.randn(rows, K, device=dev); w = torch.randn(OUT, K, device=dev) / 11; b = torch.zeros(OUT, device=dev)
    out = torch.empty(rows, OUT, device=dev); wp = prep(w)
    for _ in range(3):
        L.call('gsatb_tc_linear_fwd', ptr(x), 0, K, None, None, ptr(wp), ptr(b), ptr(out), OUT, 0, None, None, None, ctypes.c_uint64(0), ctypes.c_float(0.0), rows, K, OUT, stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.call('gsatb_tc_linear_fwd', ptr(x), 0, K, None, None, ptr(wp), ptr(b), ptr(out), OUT, 0, None, None, None, ctypes.c_uint64(0), ctypes.c_float(0.0), rows, K, OUT, stream())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'node linear 4.9M x 128 x 128: {ms:.3f} ms, {rows*(K+OUT)*4/ms/1e6:.0f} GB/s algorithmic, {2*rows*K*OUT/ms/1e9:.1f} TFLOP/s')
    t0 = time.time()
    for _ in range(10):
        ref = torch.nn.functional.linear(x, w, b)
    torch.cuda.synchronize(); print('torch fp32 linear', (time.time() - t0) * 100, 'ms')
