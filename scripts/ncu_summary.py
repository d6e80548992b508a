#!/usr/bin/env python
"""Condenses ncu output into the small tracked files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/x_launches.csv profiles/rN_launches.md
    python scripts/ncu_summary.py full gpurun_out/x.ncu-rep profiles/rN_kernels_full.json
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def short(name):
    m = re.search(r'(\w+)(<[^>]*>)?\(', name)
    base = m.group(1) + (m.group(2) or '') if m else name[:60]
    return base.replace('sc2::', '').replace('(anonymous namespace)::', '')


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write('| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n')
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.3f | %.4f | %.1f %% |\n' % (k, n, t, t / n, 100 * t / total))
        f.write('\n(ncu `--metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised: compare shares.)\n')
    print(open(dst).read())


METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__inst_executed.sum', 'smsp__cycles_elapsed.max']


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ki = h.index('Kernel Name')
    res = []
    for r in rows[2:]:
        e = {'kernel': short(r[ki] + '(')}
        full_name = r[ki]
        m = re.search(r'<([^>]*)>', full_name)
        if m:
            e['template_args'] = m.group(1)
        for mname in METRICS:
            if mname in h:
                i = h.index(mname)
                try:
                    e[mname] = float(r[i].replace(',', ''))
                except ValueError:
                    e[mname] = r[i]
                e[mname + '.unit'] = units[i]
        res.append(e)
    json.dump(res, open(dst, 'w'), indent=1)
    for e in res:
        print(e['kernel'], e.get('template_args', ''), e.get('gpu__time_duration.sum'), e.get('gpu__time_duration.sum.unit'))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
