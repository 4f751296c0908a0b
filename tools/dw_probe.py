"""GPU probe for gsatb_tc_dw (tcgen05 split-K weight gradient, MN-major / K-major TMA operands) vs torch on the same
bf16-rounded operands; then timing at the cfg4 shapes of the step."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dp_gsat_b200._lib import lib, ptr, stream

L = lib()
dev = 'cuda'


def lay(X, cm):
    rows, C = X.shape
    if cm:
        ld = (rows + 7) // 8 * 8
        T = torch.zeros(C, ld, dtype=torch.bfloat16, device=dev)
        T[:, :rows] = X.t()
        return T, ld
    ld = (C + 7) // 8 * 8
    T = torch.zeros(rows, ld, dtype=torch.bfloat16, device=dev)
    T[:, :C] = X
    return T, ld


def dw(At, a_cm, lda, Bt, b_cm, ldb, rows, M, N, bias=True):
    dW = torch.empty(M, N, device=dev)
    db = torch.empty(M, device=dev) if bias else None
    nb = int(L.cdll.gsatb_tc_dw_workspace(rows, M, N))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    L.call('gsatb_tc_dw', ptr(At), a_cm, lda, ptr(Bt), b_cm, ldb, rows, M, N, ptr(dW), N, ptr(db), 0, ptr(ws),
           ctypes.c_size_t(nb), stream())
    return dW, db


def check(rows, M, N, a_cm, b_cm):
    g = torch.Generator().manual_seed(rows + M + N)
    A = torch.randn(rows, M, generator=g).bfloat16().to(dev)
    B = torch.randn(rows, N, generator=g).bfloat16().to(dev)
    At, lda = lay(A, a_cm)
    Bt, ldb = lay(B, b_cm)
    dW, db = dw(At, a_cm, lda, Bt, b_cm, ldb, rows, M, N)
    torch.cuda.synchronize()
    ref = A.double().t() @ B.double()
    e = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
    eb = ((db.double() - A.double().sum(0)).abs().max() / A.double().sum(0).abs().max()).item()
    ok = e < 1e-4 and eb < 1e-4
    print(f'rows={rows} M={M} N={N} a_cm={a_cm} b_cm={b_cm}: rel err {e:.2e} bias {eb:.2e} {"OK" if ok else "FAIL"}', flush=True)
    return ok


def timeit(rows, M, N, a_cm, b_cm, tag):
    A = torch.randn(M, rows, device=dev).bfloat16() if a_cm else torch.randn(rows, M, device=dev).bfloat16()
    B = torch.randn(N, rows, device=dev).bfloat16() if b_cm else torch.randn(rows, N, device=dev).bfloat16()
    lda, ldb = (rows if a_cm else M), (rows if b_cm else N)
    for _ in range(3):
        dw(A, a_cm, lda, B, b_cm, ldb, rows, M, N)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dw(A, a_cm, lda, B, b_cm, ldb, rows, M, N)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = rows * (M + N) * 2 / 1e9
    print(f'{tag}: rows={rows} M={M} N={N} cm=({a_cm},{b_cm}): {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s of operand bytes, '
          f'{2.0 * rows * M * N / ms / 1e9:.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    ok = True
    for a_cm in (0, 1):
        for b_cm in (0, 1):
            ok &= check(300, 128, 128, a_cm, b_cm)
            ok &= check(5000, 80, 320, a_cm, b_cm)
            ok &= check(100000, 512, 256, a_cm, b_cm)
    ok &= check(77, 64, 64, 0, 0)
    print('ALL OK' if ok else 'SOME FAILED', flush=True)
    if ok and len(sys.argv) > 1 and sys.argv[1] == 'time':
        timeit(4_900_000, 128, 128, 0, 0, 'GIN MLP dW (row-major operands)')
        timeit(9_996_000, 128, 512, 1, 1, 'extractor dW2 (channel-major)')
        timeit(9_996_000, 512, 256, 1, 0, 'extractor dW1 (dz1^T channel-major, f12 row-major)')
        timeit(9_996_000, 512, 256, 0, 0, 'extractor dW1 (row-major both)')
    sys.exit(0 if ok else 1)
