"""Build tests/simt/build/libgsat_sim.so: the kernel sources listed in SIM_SOURCES compiled by g++ against the host
SIMT emulator (tests/simt/simt.h, -DGSATB_HOST_SIM).  TEST INFRASTRUCTURE ONLY: the product library is built by
dp_gsat_b200/build.py with nvcc; this one exists so that the `-m "not gpu"` suite can run the kernels' indexing /
mask / reduction logic on the CPU through the same C ABI entry points (host pointers instead of device pointers)."""
from __future__ import annotations

import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'dp_gsat_b200', 'csrc')
OUT_DIR = os.environ.get('GSATB_SIM_BUILD_DIR', os.path.join(HERE, 'build'))
LIB = os.path.join(OUT_DIR, 'libgsat_sim.so')
# every kernel source without tcgen05 / TMA / mbarrier PTX: the fp32 path of the step, K0, the line-graph builder, metrics
SIM_SOURCES = ['aggregate.cu', 'gine.cu', 'leconv.cu', 'pna.cu', 'index_build.cu', 'line_graph.cu', 'sampler.cu',
               'segnorm.cu', 'small_ops.cu', 'metrics.cu', 'encoders.cu', 'collate.cu', 'api.cu',
               'tc_ops.cu', 'tc_gin.cu', 'tc_extractor.cu', 'tc_extractor_bwd.cu', 'tc_dw.cu', 'ext_fused_fwd.cu', 'ext_fused_bwd.cu', 'dense.cu', 'gin_rows.cu']

def _host_has_fma() -> bool:
    try:
        return ' fma ' in open('/proc/cpuinfo').read()
    except OSError:
        return False


# nvcc contracts a * b + c into one FMA (-fmad=true, the default); let g++ do the same where the host can, so that
# cancellation-prone expressions (E[m^2] - E[m]^2 in the PNA std aggregator, ...) round as they do on the device
FMA_FLAGS = ['-mfma', '-ffp-contract=fast'] if _host_has_fma() else []

_LAUNCH = re.compile(r'([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<')


def _match(txt: str, i: int, open_ch: str, close_ch: str) -> int:
    """index just past the bracket that closes txt[i] (== open_ch)"""
    depth = 0
    for j in range(i, len(txt)):
        if txt[j] == open_ch:
            depth += 1
        elif txt[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError('unbalanced launch expression')


def _split_top(txt: str):
    parts, depth, cur = [], 0, ''
    for ch in txt:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur)
            cur = ''
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def rewrite_launches(txt: str) -> str:
    """`kernel<T...><<<grid, block, smem, stream>>>(args)` -> `simt::launch(dim3(grid), dim3(block), [&]() { kernel<T...>(args); })`
    (the product sources stay byte-identical; only this generated copy is compiled by g++).  Line continuations inside
    macros are preserved because the argument text is copied verbatim."""
    out, pos = '', 0
    while True:
        m = _LAUNCH.search(txt, pos)
        if not m:
            return out + txt[pos:]
        cfg_end = txt.index('>>>', m.end())
        cfg = _split_top(txt[m.end():cfg_end].replace('\\\n', ' '))
        smem = cfg[2] if len(cfg) >= 3 and cfg[2] else '0'
        a0 = cfg_end + 3
        while txt[a0] in ' \t\\\n':
            a0 += 1
        assert txt[a0] == '(', txt[m.start():a0 + 20]
        a1 = _match(txt, a0, '(', ')')
        out += txt[pos:m.start()] + f'simt::launch(dim3({cfg[0]}), dim3({cfg[1]}), [&]() {{ {m.group(1)}{txt[a0:a1]}; }}, {smem})'
        pos = a1


def sources():
    return [os.path.join(HERE, 'selftest.cpp')] + [os.path.join(CSRC, f) for f in SIM_SOURCES
                                                    if os.path.exists(os.path.join(CSRC, f))]


def build(force: bool = False) -> str:
    deps = sources() + [os.path.join(HERE, h) for h in os.listdir(HERE) if h.endswith('.h')] + \
        [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith('.cuh')] + [
                        os.path.join(ROOT, 'include', 'gsat_b200.h'), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    gen = os.path.join(OUT_DIR, 'gen')
    os.makedirs(gen, exist_ok=True)
    for h in sorted(os.listdir(CSRC)):                      # headers may launch kernels too (tc_gemm.cuh)
        if h.endswith('.cuh'):
            with open(os.path.join(gen, h), 'w') as f:
                f.write(f'#line 1 "{os.path.join(CSRC, h)}"\n' + rewrite_launches(open(os.path.join(CSRC, h)).read()))
    procs = []
    for s in sources():
        base = os.path.splitext(os.path.basename(s))[0]
        g = os.path.join(gen, base + '.cpp')
        with open(g, 'w') as f:
            f.write(f'#line 1 "{s}"\n' + rewrite_launches(open(s).read()))
        o = os.path.join(OUT_DIR, base + '.o')
        cmd = ['g++', '-x', 'c++', '-std=c++17', '-O2', '-g', '-fPIC', '-DGSATB_HOST_SIM', '-U_FORTIFY_SOURCE'] + FMA_FLAGS + ['-I', HERE, '-I', gen, '-I', CSRC,
               '-Wno-attributes', '-Wno-unused', '-Wno-unknown-pragmas', '-c', g, '-o', o]
        procs.append((s, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, o, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'g++ (host SIMT build) failed on {s}:\n{out[-6000:]}')
        objs.append(o)
    r = subprocess.run(['g++', '-shared', '-o', LIB] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv))
