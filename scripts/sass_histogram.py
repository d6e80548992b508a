#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libsc2b200.so: the evidence that the kernels are Blackwell-native (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit; HMMA would be the legacy mma.sync path).
    python scripts/sass_histogram.py > profiles/r4_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'sc2-benchmark_b200', 'lib', 'libsc2b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'UBLKCP', 'SYNCS', 'HMMA', 'FFMA', 'LDGSTS']
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
hist, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void ', '').replace('sc2::', '')
        hist[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and name:
        op = m.group(1)
        hist[name]['total'] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[name][k] += 1
print('# SASS opcode histogram per kernel (cuobjdump -sass sc2-benchmark_b200/lib/libsc2b200.so)\n')
print('| kernel | instr | ' + ' | '.join(KEYS) + ' |')
print('|---|---:|' + '---:|' * len(KEYS))
tot = collections.Counter()
for name, c in hist.items():
    tot.update(c)
    print('| `%s` | %d | %s |' % (name[:90], c['total'], ' | '.join(str(c[k]) if c[k] else '' for k in KEYS)))
print('| **all kernels** | %d | %s |' % (tot['total'], ' | '.join(str(tot[k]) for k in KEYS)))
