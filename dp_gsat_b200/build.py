"""Build libgsat_b200.so in-tree with nvcc for sm_100a (no torch / pybind linkage: plain C ABI)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgsat_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--use_fast_math=false',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '-shared', '-cudart', 'static']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'gsat_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    procs = []
    flags = [f for f in FLAGS if f not in ('-shared', '--use_fast_math=false')]
    for s in sources():
        o = os.path.join(HERE, 'build', os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        cmd = [NVCC] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {s}')
    cmd = [NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-cudart', 'static', '-o', LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
