#!/usr/bin/env python
"""Dynamic instruction mix (warp-level executed counts by opcode) of a profiled kernel."""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     stdout=subprocess.PIPE, universal_newlines=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
mix = Counter()
for row in rows:
    if 'Address' in row and 'Source' in row:
        hdr = row
        continue
    if not hdr or len(row) != len(hdr):
        continue
    d = dict(zip(hdr, row))
    text = d['Source']
    parts = text.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    mix[op] += float(d['Instructions Executed'] or 0)
total = sum(mix.values())
print('total warp-instructions', total)
for op, n in mix.most_common(30):
    print('{:10s} {:12.0f} {:5.1f}%'.format(op, n, 100 * n / total))
