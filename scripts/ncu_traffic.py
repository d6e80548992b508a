#!/usr/bin/env python
"""ncu raw-page CSV of ONE pass of the bottleneck's transforms (scripts/prof_transforms.py, batch 256) -> profiles/rN_traffic.json
keyed by bench.py's kernel tags, plus a markdown table of the per-kernel metrics.

    ncu --set full --clock-control none -k regex:"tc_|ga_|nchw" -s 18 -c 9 -o /tmp/x python scripts/prof_transforms.py
    ncu -i /tmp/x.ncu-rep --page raw --csv > gpurun_out/x_raw.csv
    python scripts/ncu_traffic.py gpurun_out/x_raw.csv profiles/rN_traffic.json profiles/rN_transforms_ncu.md
"""
import csv
import json
import sys

# launch order of one pass (g_a: first + GDN1, mid + GDN1, last + quantise; g_s: layout, K9, IGDN1(512), K10, IGDN1(256), K11)
TAGS = ['ga_first[3->96,k5,s2]+gdn1', 'ga_halo[96->48,k5,s2]+gdn1', 'tc_split[48->24,k2,s1,m2]', 'nchw_to_nhwc_f16', 'tc_conv[64->512,k2,m4]',
        'tc_conv[512->512,k1,m5]', 'tc_conv[512->256,k2,m4]', 'tc_conv[256->256,k1,m5]', 'tc_conv[256->256,k2,m1]']
EXPECT = ['ga_first', 'ga_halo', 'tc_split_conv', 'nchw', 'tc_conv', 'tc_conv', 'tc_conv', 'tc_conv', 'tc_conv']
COLS = [('gpu__time_duration.sum', 'ms', 1e-3), ('dram__bytes_read.sum', 'DRAM read GB', 1.0), ('dram__bytes_write.sum', 'DRAM write GB', 1.0),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %', 1.0),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %', 1.0),
        ('lts__t_sectors_srcunit_tex.sum', 'L2->SM GB', 32e-9), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %', 1.0),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %', 1.0),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts', 1.0), ('launch__registers_per_thread', 'regs', 1.0)]


def to_bytes(v, unit):
    return float(v) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)


def main(src, dst_json, dst_md):
    rows = list(csv.reader(open(src)))
    h, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(h)}
    body = rows[2:]
    if len(body) != len(TAGS):
        raise SystemExit('expected %d launches (one pass), got %d' % (len(TAGS), len(body)))
    out, lines = {}, ['| kernel (bench tag) | ' + ' | '.join(c[1] for c in COLS) + ' |', '|---|' + '---:|' * len(COLS)]
    for tag, want, r in zip(TAGS, EXPECT, body):
        name = r[ix['Kernel Name']]
        if want not in name:
            raise SystemExit('launch order changed: %s is not a %s kernel' % (name[:60], want))
        rd = to_bytes(r[ix['dram__bytes_read.sum']], units[ix['dram__bytes_read.sum']])
        wr = to_bytes(r[ix['dram__bytes_write.sum']], units[ix['dram__bytes_write.sum']])
        t = float(r[ix['gpu__time_duration.sum']]) * {'us': 1e-3, 'ms': 1.0, 'ns': 1e-6}[units[ix['gpu__time_duration.sum']]]
        out[tag] = {'dram_bytes_per_launch': rd + wr, 'dram_read': rd, 'dram_write': wr, 'ncu_ms': t, 'batch': 256,
                    'tensor_pipe_active_pct': float(r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
                    'dram_throughput_pct': float(r[ix['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]),
                    'l2_to_sm_bytes': float(r[ix['lts__t_sectors_srcunit_tex.sum']]) * 32.0}
        cells = []
        for m, _, scale in COLS:
            v, u = r[ix[m]], units[ix[m]]
            if m.startswith('dram__bytes'):
                cells.append('%.3f' % (to_bytes(v, u) / 1e9))
            elif m == 'gpu__time_duration.sum':
                cells.append('%.3f' % t)
            elif m.startswith('lts__t_sectors'):
                cells.append('%.2f' % (float(v) * scale))
            else:
                cells.append('%.1f' % float(v) if '.' in v else v)
        lines.append('| `%s` | ' % tag + ' | '.join(cells) + ' |')
    json.dump(out, open(dst_json, 'w'), indent=1)
    open(dst_md, 'w').write('\n'.join(lines) + '\n\n(`ncu --set full --clock-control none`, one pass of scripts/prof_transforms.py at batch 256, cold L2, '
                            'serialised: compare with the CUDA-event times of bench.py by SHARE.)\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main(*sys.argv[1:4])
