#!/usr/bin/env python
"""Dump SASS with per-instruction samples / stall reasons for an address window.
    python tools/ncu_sass.py report.ncu-rep [min_samples]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     stdout=subprocess.PIPE, universal_newlines=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for row in rows:
    if 'Address' in row and 'Source' in row:
        hdr = row
        continue
    if not hdr or len(row) != len(hdr):
        continue
    d = dict(zip(hdr, row))
    smp = float(d['# Samples'] or 0)
    stalls = sorted(((float(d[c] or 0), c.replace('stall_', '')) for c in hdr
                     if c.startswith('stall_') and 'Not Issued' not in c), reverse=True)[:3]
    mark = '*' if smp >= thr else ' '
    print('{} {} {:6.0f} {:9s} {:60s} {}'.format(mark, d['Address'][-5:], smp, d['Instructions Executed'],
          d['Source'][:60], ' '.join('{}:{:.0f}'.format(n, v) for v, n in stalls if v > 0)))
