"""Summarise an `ncu --page source --csv` dump of a warp-specialised kernel: instructions executed and stall-reason
samples per code region (regions split at the USETMAXREG instructions = the role branches)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) > col['stall_wait']]
marks = [i for i, r in enumerate(data) if 'USETMAXREG' in r[col['Source']]] + [len(data)]
bounds = [0] + marks
tot_inst = sum(int(r[col['Instructions Executed']]) for r in data)
tot_smp = sum(int(r[col['# Samples']]) for r in data)
print(f'total warp-instructions {tot_inst}, samples {tot_smp}')
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = data[a:b]
    if not seg:
        continue
    inst = sum(int(r[col['Instructions Executed']]) for r in seg)
    smp = sum(int(r[col['# Samples']]) for r in seg)
    st = {h: sum(int(r[col[h]]) for r in seg) for h in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print(f'sass[{a}:{b}] inst {100.0 * inst / tot_inst:5.1f}%  samples {100.0 * smp / tot_smp:5.1f}%  ' +
          ' '.join(f'{k[6:]}={100.0 * v / max(smp, 1):.0f}%' for k, v in top))
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    seg = data[a:b]
    hot = sorted(range(len(seg)), key=lambda i: -int(seg[i][col['# Samples']]))[:25]
    for i in sorted(hot):
        r = seg[i]
        st = sorted(((h[6:], int(r[col[h]])) for h in stalls), key=lambda kv: -kv[1])[:2]
        print(a + i, r[col['Instructions Executed']], r[col['# Samples']], st, r[col['Source']][:70])
