"""ctypes binding of libgsat_b200.so (C ABI declared in include/gsat_b200.h).

There is NO fallback: if the shared library is missing or the device is not sm_100-class the import / first call
raises.  Argument types are derived from the header itself so the binding cannot drift from the declared ABI.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgsat_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'gsat_b200.h')

_CTYPES = {
    'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64, 'size_t': ctypes.c_size_t,
    'float': ctypes.c_float, 'double': ctypes.c_double, 'gsatb_stream_t': ctypes.c_void_p, 'void': None,
}


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[object, List[object]]]:
    """Return {symbol: (restype, [argtypes])} for every ``gsatb_*`` prototype in the header."""
    txt = open(path).read()
    txt = re.sub(r'/\*.*?\*/', ' ', txt, flags=re.S)
    txt = re.sub(r'//[^\n]*', ' ', txt)
    txt = re.sub(r'^\s*#.*$', ' ', txt, flags=re.M)
    protos: Dict[str, Tuple[object, List[object]]] = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(gsatb_\w+)\s*\(([^;{}]*?)\)\s*;', txt, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if 'typedef' in ret:
            continue

        def conv(decl: str):
            decl = decl.replace('const', ' ').strip()
            if '*' in decl:
                return ctypes.c_char_p if decl.split('*')[0].strip() == 'char' and name == 'gsatb_strerror' else ctypes.c_void_p
            base = decl.split()[0]
            return _CTYPES[base]
        if '*' in ret:
            restype = ctypes.c_char_p
        else:
            restype = _CTYPES[ret.replace('const', '').split()[0]]
        argtypes = [] if args in ('', 'void') else [conv(a) for a in args.split(',')]
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} not found: build it with `python -m dp_gsat_b200.build` (or __graft_entry__.build()). '
                'dp_gsat_b200 has no CPU / eager fallback.')
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._device_checked = False
        self._step_counter = None
        self.launches = 0          # kernels launched through this library (bench.py reports the timed-region delta)
        self.timer = None          # optional {entry name: [(start_event, end_event, tag), ...]} filled by call()
        self.timer_all = False     # time every entry point, not only the names already in `timer`
        self.timer_tag = None      # optional fn(name, args) -> tag stored with each timed call (bench.py: shapes)

    def step_counter(self, device=None):
        """The process-wide device step counter (int64 [1]) registered with gsatb_set_step_counter; created on first
        use and never freed, so the pointer the library holds cannot dangle."""
        if self._step_counter is None:
            self._step_counter = torch.zeros(1, dtype=torch.int64, device=device or 'cuda')
            self.cdll.gsatb_set_step_counter(ctypes.c_void_p(self._step_counter.data_ptr()))
        return self._step_counter

    def strerror(self, code: int) -> str:
        return self.cdll.gsatb_strerror(code).decode()

    def check_device(self):
        if self._device_checked:
            return
        if not torch.cuda.is_available():
            raise RuntimeError('dp_gsat_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        rc = self.cdll.gsatb_check_device()
        if rc != 0:
            raise RuntimeError(f'dp_gsat_b200: {self.strerror(rc)}')
        self._device_checked = True

    # kernels launched per C-ABI call (index_build: see csrc/index_build.cu -- 5 fixed + 2 orders x (2 sorts x
    # 3 kernels x passes + 3) + 1; counted with passes=3)
    KERNELS_PER_CALL = {'gsatb_index_build': 5 + 2 * (2 * 3 * 3 + 3) + 1, 'gsatb_sample_avg_info_fwd': 2,
                        'gsatb_le_aggregate_bwd': 2,       # by source (da, d att, d w) + by destination (db)
                        'gsatb_embedding_sum_bwd': 2}      # per-CTA slabs + fixed-order reduction

    def call(self, name: str, *args):
        """Call an int-returning entry point; raise on a non-zero code."""
        self.check_device()
        self.launches += self.KERNELS_PER_CALL.get(name, 1)
        t = self.timer
        if t is not None and (name in t or self.timer_all):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()                      # on torch's current stream == the stream the kernel is launched on
            rc = getattr(self.cdll, name)(*args)
            e1.record()
            t.setdefault(name, []).append((e0, e1, self.timer_tag(name, args) if self.timer_tag else None))
        else:
            rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            msg = f'{name} failed: {self.strerror(rc)} (code {rc})'
            if rc in (-1, -2, -3, -6):
                raise ValueError(msg)
            raise RuntimeError(msg)


_LIB = None


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def require_cuda(t) -> None:
    """Every op of this package takes CUDA tensors: there is no CPU path to fall back to."""
    if not t.is_cuda:
        raise RuntimeError('CUDA tensor expected (dp_gsat_b200 has no CPU path)')


def device_guard(device):
    """Context manager that makes ``device`` current for raw C-ABI calls issued inside it."""
    return torch.cuda.device(device)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
