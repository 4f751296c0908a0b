"""ctypes binding of tests/simt/build/libgsat_sim.so (the host-SIMT build of selected kernel sources): same C ABI, same
header-derived argument types as dp_gsat_b200/_lib.py, but the pointers are HOST pointers of CPU tensors."""
from __future__ import annotations

import ctypes
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class SimLib:
    def __init__(self):
        build = _load('_gsatb_sim_build', os.path.join(HERE, 'build_sim.py'))
        binding = _load('_gsatb_binding', os.path.join(ROOT, 'dp_gsat_b200', '_lib.py'))     # parse_header only
        self.cdll = ctypes.CDLL(build.build())
        self.protos = binding.parse_header()

    def has(self, name):
        return hasattr(self.cdll, name)

    def call(self, name, *args):
        fn = getattr(self.cdll, name)
        fn.restype, fn.argtypes = self.protos[name]
        conv = []
        for a in args:
            if isinstance(a, torch.Tensor):
                assert not a.is_cuda and a.is_contiguous()
                conv.append(ctypes.c_void_p(a.data_ptr()))
            else:
                conv.append(a)
        return fn(*conv)


_SIM = None


def sim() -> SimLib:
    global _SIM
    if _SIM is None:
        _SIM = SimLib()
    return _SIM


def guarded(shape, dtype=torch.float32, fill=None):
    """An output tensor carved out of a larger buffer with canary words on both sides; check with `intact`."""
    n = 1
    for s in shape:
        n *= s
    pad = 64
    canary = {torch.float32: 1.2345e30, torch.int64: -0x5A5A5A5A5A5A, torch.int32: -0x5A5A5A5, torch.uint8: 0xA5}[dtype]
    buf = torch.full((n + 2 * pad,), canary, dtype=dtype)
    view = buf[pad:pad + n].view(*shape)
    if fill is not None:
        view.fill_(fill)
    view._guard = (buf, pad, n, canary)
    return view


def intact(view) -> bool:
    buf, pad, n, canary = view._guard
    ref = torch.full((pad,), canary, dtype=buf.dtype)
    return bool(torch.equal(buf[:pad], ref) and torch.equal(buf[pad + n:], ref))
