"""SASS evidence that the tensor-core kernels use tcgen05 / TMA / TMEM: counts of the Blackwell mnemonics per kernel of
the built library (cuobjdump -sass).  Writes profiles/r2_sass_summary.txt.

  UTCHMMA   tcgen05.mma kind::f16        UTMALDG   TMA tensor load       UTMASTG   TMA tensor store
  LDTM      tcgen05.ld (TMEM -> regs)    UTCBAR    tcgen05.commit        SYNCS     mbarrier ops"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'dp_gsat_b200', 'libgsat_b200.so')
MNEMONICS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA', 'LDG', 'STG', 'LDS', 'STS']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(['c++filt', n], stdout=subprocess.PIPE, text=True).stdout.strip()
    rows, name, counts = [], None, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if name:
                rows.append((name, counts))
            name, counts = m.group(1), dict.fromkeys(MNEMONICS, 0)
            continue
        if name:
            m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if m:
                op = m.group(1).split('.')[0]
                if op in counts:
                    counts[op] += 1
    if name:
        rows.append((name, counts))
    lines = [f'cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static)', '',
             f'{"kernel":72s} ' + ' '.join(f'{m:>8s}' for m in MNEMONICS)]
    tot = dict.fromkeys(MNEMONICS, 0)
    for name, c in sorted(rows, key=lambda r: -(r[1]['UTCHMMA'] * 1000 + r[1]['UTMALDG'])):
        dn = demangle(name)
        dn = dn.replace('(anonymous namespace)::', '').replace('void ', '', 1)
        dn = re.sub(r'\((?!.*>).*$', '', dn) if '>' in dn else re.sub(r'\(.*$', '', dn)
        dn = dn[:72]
        lines.append(f'{dn:72s} ' + ' '.join(f'{c[m]:8d}' for m in MNEMONICS))
        for m in MNEMONICS:
            tot[m] += c[m]
    lines.append(f'{"TOTAL":72s} ' + ' '.join(f'{tot[m]:8d}' for m in MNEMONICS))
    txt = '\n'.join(lines) + '\n'
    path = os.path.join(ROOT, 'profiles', 'r2_sass_summary.txt')
    open(path, 'w').write(txt)
    sys.stdout.write(txt)


if __name__ == '__main__':
    main()
