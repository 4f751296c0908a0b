#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into the text committed under profiles/: per captured launch the
duration, DRAM bytes read / written, DRAM throughput, tensor-pipe activity, issue-slot utilisation, warps, registers and
the dominant stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [profiles/out.txt] [traffic.json]"""
import csv
import json
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'duration'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active % (elapsed)'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % of max'),
        ('launch__registers_per_thread', 'registers / thread'), ('launch__grid_size', 'grid'),
        ('launch__block_size', 'block'), ('smsp__inst_executed.sum', 'warp instructions'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts')]
STALLS = ['long_scoreboard', 'barrier', 'sleeping', 'short_scoreboard', 'wait', 'mio_throttle', 'math_pipe_throttle',
          'not_selected', 'no_instruction', 'branch_resolving', 'lg_throttle', 'dispatch_stall']


def to_bytes(v, unit):
    f = float(v.replace(',', ''))
    return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main(rep, out=None, traffic_json=None):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f'# ncu --set full --clock-control none, report {rep} ({len(rows) - 2} captured launches)', '']
    traffic = {}
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = d['Kernel Name']
        lines.append(f'== {name[:150]}')
        for k, label in WANT:
            if k in d:
                lines.append(f'   {label:38s} {d[k]} {u[k]}')
        if 'dram__bytes_read.sum' in d:
            tot = to_bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + \
                to_bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
            lines.append(f'   {"DRAM read + write":38s} {tot / 1e9:.3f} GB')
            short = name.split('(')[0].split('::')[-1]
            traffic.setdefault(short, []).append(tot)
        st = []
        for s in STALLS:
            k = f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio'
            if k in d:
                st.append((float(d[k].replace(',', '')), s))
        st.sort(reverse=True)
        lines.append('   stall reasons (warps per issue)        ' + ', '.join(f'{s} {v:.2f}' for v, s in st[:5]))
        lines.append('')
    text = '\n'.join(lines) + '\n'
    if out:
        open(out, 'w').write(text)
    sys.stdout.write(text)
    if traffic_json:
        json.dump(traffic, open(traffic_json, 'w'), indent=1)


if __name__ == '__main__':
    main(*sys.argv[1:4])
