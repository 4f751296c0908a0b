#!/usr/bin/env python
"""Dry run of bench.py's control flow WITHOUT a GPU: `python tools/bench_dry_run.py [--ranks 2] [--graphs 96]`.

bench.py's own code (the b200 arm: warm-up, preload, timed region, e2e loop with the prefetching copy stream, max over
ranks, the JSON line) runs unchanged on N gloo ranks; the kernels run on the host SIMT emulator (tests/simt) and the CUDA
runtime objects it touches (events, streams, pinned memory, NCCL) are replaced by inert stand-ins.  What this checks
is that every rank issues the SAME sequence of collectives and reaches the end (an earlier version of the pre-load loop
let ranks run different step counts and dead-locked the 2-GPU run) and that the line has every contract key.  The
numbers it prints are meaningless (CPU emulation) and are never reported."""
import argparse
import importlib.util
import json
import os
import socket
import sys
import time

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


class _Stream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def wait_event(self, ev):
        pass

    def wait_stream(self, s):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _patch_cuda_runtime():
    from tests.simt import emulate
    emulate.redirect_torch_to_cpu(setattr)
    emulate.patch_product(setattr)
    torch.cuda.set_device = lambda d: None
    torch.cuda.Event = _Event
    torch.cuda.Stream = _Stream
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.stream = lambda s: s
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.record_stream = lambda self, s: None
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init('gloo', **{k: v for k, v in kw.items() if k != 'device_id'})


def _worker(rank, world, port, a, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    _patch_cuda_runtime()
    spec = importlib.util.spec_from_file_location('_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sys.argv = ['bench.py', '--gpus', str(world), '--steps', str(a.steps), '--warmup', str(a.warmup), '--graphs',
                str(a.graphs), '--hidden', str(a.hidden), '--precision', a.precision, '--cuda-graph', 'off',
                '--e2e-steps', '2'] + ([] if a.with_cpu_baseline else ['--no-cpu-baseline'])
    args = bench.parse()
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.preload.__defaults__ = (0.05,)            # keep the emulated pre-load short
        bench.run_b200(args)
    if rank == 0:
        open(out, 'w').write(buf.getvalue())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ranks', type=int, default=2)
    ap.add_argument('--graphs', type=int, default=96)
    ap.add_argument('--hidden', type=int, default=64)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--with-cpu-baseline', action='store_true', help='also run the cpu_baseline leg (rank 0, N = 1)')
    a = ap.parse_args()
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = f'/tmp/bench_dry_run_{os.getpid()}.json'
    if a.ranks == 1:
        _worker(0, 1, port, a, out)
    else:
        mp.spawn(_worker, args=(a.ranks, port, a, out), nprocs=a.ranks, join=True)
    line = json.loads(open(out).read().strip().splitlines()[-1])
    need = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'cpu_baseline', 'e2e', 'gpu_launches', 'clocks'}
    missing = need - set(line)
    assert not missing, missing
    assert line['n_gpus'] == a.ranks and line['e2e']['h2d_bytes_per_step'] > 0
    if a.with_cpu_baseline and a.ranks == 1:
        cb = line['cpu_baseline']
        assert cb['value'] > 0 and cb['reference_thread_setting']['value'] > 0, cb
    print(f"bench.py control flow ok on {a.ranks} rank(s): keys complete, n_gpus {line['n_gpus']}, "
          f"gpu_launches {line['gpu_launches']}, e2e steps {line['e2e']['steps']}  (timings are emulation: not reported)")


if __name__ == '__main__':
    main()
