"""Build tests/simt/build/libgsat_sim.so: the kernel sources listed in SIM_SOURCES compiled by g++ against the host
SIMT emulator (tests/simt/simt.h, -DGSATB_HOST_SIM).  TEST INFRASTRUCTURE ONLY: the product library is built by
dp_gsat_b200/build.py with nvcc; this one exists so that the `-m "not gpu"` suite can run the kernels' indexing /
mask / reduction logic on the CPU through the same C ABI entry points (host pointers instead of device pointers)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'dp_gsat_b200', 'csrc')
OUT_DIR = os.path.join(HERE, 'build')
LIB = os.path.join(OUT_DIR, 'libgsat_sim.so')
# kernel sources that launch through GSATB_LAUNCH and use no inline PTX outside common.cuh's guarded helpers
SIM_SOURCES = ['leconv.cu', 'encoders.cu', 'collate.cu']


def sources():
    return [os.path.join(HERE, 'selftest.cpp')] + [os.path.join(CSRC, f) for f in SIM_SOURCES
                                                    if os.path.exists(os.path.join(CSRC, f))]


def build(force: bool = False) -> str:
    deps = sources() + [os.path.join(HERE, 'simt.h'), os.path.join(CSRC, 'common.cuh'),
                        os.path.join(ROOT, 'include', 'gsat_b200.h')]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    for s in sources():
        o = os.path.join(OUT_DIR, os.path.splitext(os.path.basename(s))[0] + '.o')
        cmd = ['g++', '-x', 'c++', '-std=c++17', '-O1', '-g', '-fPIC', '-DGSATB_HOST_SIM', '-I', HERE,
               '-Wno-attributes', '-Wno-unused', '-c', s, '-o', o]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'g++ (host SIMT build) failed on {s}:\n{r.stdout}')
        objs.append(o)
    r = subprocess.run(['g++', '-shared', '-o', LIB] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv))
