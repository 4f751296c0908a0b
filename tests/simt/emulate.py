"""Run the product's Python stack (dp_gsat_b200.*) against the host-SIMT build of the kernel sources.

TEST INFRASTRUCTURE ONLY.  Two users:

* the ``emulated`` fixture of the CPU suite (tests/test_simt_step.py): the package's autograd Functions, GraphIndex, GSAT
  step, loaders ... run unchanged on CPU tensors, every C-ABI call lands in tests/simt/build/libgsat_sim.so -- the same
  .cu sources compiled by g++ against tests/simt/simt.h.  Nothing in the product imports this; the product itself still
  refuses CPU tensors (``_lib.require_cuda``) and has no CPU path.
* ``python -m pytest -p tests.simt.emulate -m gpu tests/test_gpu_parity.py ...``: a DRY RUN of the GPU suites on a machine
  without a GPU.  As a pytest plugin this module additionally redirects ``.cuda()`` / ``device='cuda'`` in torch to the
  CPU, so the GPU test bodies themselves execute (against the emulated kernels, the tcgen05 ones included).  Tests
  that need CUDA graphs or BASELINE-size inputs are skipped.  A dry run is not a GPU result and is never reported as one.
"""
from __future__ import annotations

import contextlib
import ctypes
import importlib

import torch

from .simlib import sim

_PRODUCT_MODULES = ('_lib', 'ops', 'index', 'gsat', 'loader', 'dual', 'metrics', 'nn', 'pna', 'dense')


class EmulatedLib:
    """Same contract as dp_gsat_b200._lib._Lib, backed by the emulator build."""

    KERNELS_PER_CALL = {}

    def __init__(self):
        s = sim()
        self.cdll, self.protos = s.cdll, s.protos
        self.missing = []
        for name, (restype, argtypes) in self.protos.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:          # an entry point whose source is not part of the emulator build
                self.missing.append(name)
                continue
            fn.restype, fn.argtypes = restype, argtypes
        self.launches = 0
        self.timer, self.timer_all, self.timer_tag = None, False, None
        self._step_counter = None

    def strerror(self, code: int) -> str:
        return self.cdll.gsatb_strerror(code).decode()

    def check_device(self):
        pass

    def step_counter(self, device=None):
        if self._step_counter is None:
            self._step_counter = torch.zeros(1, dtype=torch.int64)
            self.cdll.gsatb_set_step_counter(ctypes.c_void_p(self._step_counter.data_ptr()))
        return self._step_counter

    def call(self, name: str, *args):
        if name in self.missing:
            raise NotImplementedError(f'{name} is not part of the emulator build')
        self.launches += 1
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            msg = f'{name} failed: {self.strerror(rc)} (code {rc})'
            raise (ValueError if rc in (-1, -2, -3, -6) else RuntimeError)(msg)


_EMU = None


def emulated_lib() -> EmulatedLib:
    global _EMU
    if _EMU is None:
        _EMU = EmulatedLib()
    return _EMU


def patch_product(setattr_fn) -> EmulatedLib:
    """Point every product module at the emulated library.  ``setattr_fn(obj, name, value)`` does the patching
    (monkeypatch.setattr in the fixture: undone after the test)."""
    emu = emulated_lib()
    for m in _PRODUCT_MODULES:
        mod = importlib.import_module(f'dp_gsat_b200.{m}')
        if hasattr(mod, 'lib'):
            setattr_fn(mod, 'lib', lambda: emu)
        if hasattr(mod, 'stream'):
            setattr_fn(mod, 'stream', lambda: None)
        if hasattr(mod, 'require_cuda'):
            setattr_fn(mod, 'require_cuda', lambda t: None)
        if hasattr(mod, 'device_guard'):
            setattr_fn(mod, 'device_guard', lambda dev: contextlib.nullcontext())
    tc = importlib.import_module('dp_gsat_b200.tc')
    setattr_fn(tc, 'lib', lambda: emu)
    setattr_fn(tc, 'stream', lambda: None)
    index = importlib.import_module('dp_gsat_b200.index')
    index.clear_index_cache()
    return emu


# ---------------------------------------------------------------------------------------------------------------
# pytest plugin: dry run of the GPU suites (-p tests.simt.emulate)
# ---------------------------------------------------------------------------------------------------------------
_SKIP_IN_DRY_RUN = ('test_cuda_graph_training_step', 'test_large_batch_properties', 'one_big_graph',
                    '300000')


def _to_cpu_device(d):
    if isinstance(d, str) and d.startswith('cuda'):
        return 'cpu'
    if isinstance(d, torch.device) and d.type == 'cuda':
        return torch.device('cpu')
    return d


def redirect_torch_to_cpu(setattr_fn):
    """``.cuda()``, ``.to('cuda')`` and ``device='cuda'`` become no-ops / CPU placements, so that a GPU test body runs on
    host tensors.  ``setattr_fn`` as in patch_product."""
    setattr_fn(torch.cuda, 'is_available', lambda: True)
    setattr_fn(torch.cuda, 'synchronize', lambda *a, **k: None)
    setattr_fn(torch.cuda, 'is_current_stream_capturing', lambda: False)      # asked by torch.optim when cuda 'is available'
    setattr_fn(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    setattr_fn(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
    t_to, m_to = torch.Tensor.to, torch.nn.Module.to

    def tensor_to(self, *a, **k):
        a = tuple(_to_cpu_device(x) for x in a)
        if 'device' in k:
            k['device'] = _to_cpu_device(k['device'])
        return t_to(self, *a, **k)

    def module_to(self, *a, **k):
        a = tuple(_to_cpu_device(x) for x in a)
        if 'device' in k:
            k['device'] = _to_cpu_device(k['device'])
        return m_to(self, *a, **k)
    setattr_fn(torch.Tensor, 'to', tensor_to)
    setattr_fn(torch.nn.Module, 'to', module_to)
    for name in ('zeros', 'ones', 'empty', 'full', 'arange', 'tensor', 'as_tensor', 'rand', 'randn', 'randint', 'eye',
                 'randperm', 'linspace', 'zeros_like', 'empty_like', 'ones_like', 'full_like', 'rand_like', 'randn_like'):
        orig = getattr(torch, name)

        def wrapped(*a, __orig=orig, **k):
            if 'device' in k:
                k['device'] = _to_cpu_device(k['device'])
            return __orig(*a, **k)
        setattr_fn(torch, name, wrapped)


def pytest_configure(config):
    redirect_torch_to_cpu(setattr)
    patch_product(setattr)


def pytest_collection_modifyitems(config, items):
    import pytest
    for it in items:
        if any(s in it.nodeid for s in _SKIP_IN_DRY_RUN):
            it.add_marker(pytest.mark.skip(reason='not part of the emulator dry run (CUDA graph / full size)'))
