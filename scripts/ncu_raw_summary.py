#!/usr/bin/env python
"""Condenses the raw-page CSV of the ncu --set full capture of scripts/profile_step.py (passes 2-4, 15 kernels each) into
profiles/<tag>_kernels_ncu_full.json and profiles/<tag>_traffic.json.   python scripts/ncu_raw_summary.py gpurun_out/x_full_raw.csv profiles/x"""
import csv
import json
import re
import sys

ORDER = ['tc_first[3->96,k5,s2]', 'tc_split[96->96,k1,s1,m1]', 'tc_split[96->48,k5,s2,m0]', 'tc_split[48->48,k1,s1,m1]',
         'tc_split[48->24,k2,s1,m2]', 'rans_encode', 'rans_offsets', 'rans_pack', 'rans_decode', 'nchw_to_nhwc_f16',
         'tc_conv[64->512,k2,m0]', 'tc_conv[512->512,k1,m2]', 'tc_conv[512->256,k2,m0]', 'tc_conv[256->256,k1,m2]', 'tc_conv[256->256,k2,m1]']
PASSES = ['warp-per-stream coder (pass 2)', 'lane-per-stream coder (pass 3, warm-up)', 'lane-per-stream coder (pass 4)']
METR = ['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]


def val(r, name):
    try:
        return float(r[h.index(name)].replace(',', ''))
    except ValueError:
        return None


def unit(name):
    return units[h.index(name)]


BYTES = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
MS = {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}
out = []
for i, r in enumerate(rows[2:]):
    e = {'pass': PASSES[i // 15], 'tag': ORDER[i % 15], 'kernel': re.sub(r'\(.*', '', r[h.index('Kernel Name')])[-70:],
         'ncu_ms': val(r, 'gpu__time_duration.sum') * MS[unit('gpu__time_duration.sum')],
         'dram_read': val(r, 'dram__bytes_read.sum') * BYTES[unit('dram__bytes_read.sum')],
         'dram_write': val(r, 'dram__bytes_write.sum') * BYTES[unit('dram__bytes_write.sum')]}
    for m in METR:
        if m in h:
            e[m] = val(r, m)
    out.append(e)
keep = [e for e in out if 'pass 3' not in e['pass']]
json.dump(keep, open(sys.argv[2] + '_kernels_ncu_full.json', 'w'), indent=1)
traffic = {}
for e in keep:
    lanes = 'lane' in e['pass']
    if e['tag'] in ('rans_encode', 'rans_decode'):
        key = e['tag'] if lanes else e['tag'] + '[warp]'
    elif lanes:
        continue
    else:
        key = e['tag']
    traffic[key] = {'dram_bytes_per_launch': e['dram_read'] + e['dram_write'], 'dram_read': e['dram_read'], 'dram_write': e['dram_write'],
                    'ncu_ms': e['ncu_ms'], 'tensor_pipe_active_pct': e.get(METR[0]), 'dram_throughput_pct': e.get(METR[1]), 'batch': 256}
json.dump(traffic, open(sys.argv[2] + '_traffic.json', 'w'), indent=1)
for e in keep:
    if 'lane' in e['pass'] and not e['tag'].startswith('rans_e') and not e['tag'].startswith('rans_d'):
        continue
    print('%-28s %8.3f ms  dram %7.1f MB  tensor %5.1f%%  regs %3d  grid %5d x %3d  bank conflicts %.1fM' % (
        e['tag'] + (' [lanes]' if 'lane' in e['pass'] else ''), e['ncu_ms'], (e['dram_read'] + e['dram_write']) / 1e6, e.get(METR[0]) or 0,
        e.get(METR[2]) or 0, e.get(METR[3]) or 0, e.get(METR[4]) or 0, (e.get(METR[8]) or 0) / 1e6))
