#!/usr/bin/env python
"""Warp-stall samples of each profiled kernel of an .ncu-rep by SASS region (regions split
at barriers / branches): which phase of a fused kernel the time goes to.
    python tools/ncu_phases.py report.ncu-rep [min percent = 1.2]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     stdout=subprocess.PIPE, universal_newlines=True).stdout
kernels, hdr = [], None
for row in csv.reader(io.StringIO(out)):
    if 'Address' in row and 'Source' in row:
        hdr = row
        kernels.append([])
        continue
    if hdr and len(row) == len(hdr):
        kernels[-1].append(dict(zip(hdr, row)))
for num, insts in enumerate(kernels):
    total = sum(float(i['# Samples'] or 0) for i in insts) or 1.0
    print('== kernel', num, 'instructions', len(insts), 'samples', total)
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    region, regions = [], []
    for ins in insts:
        region.append(ins)
        text = ins['Source']
        if text.startswith('BAR') or 'BAR.' in text or text.startswith('WARPSYNC') or \
                text.startswith('BRA') or (text.startswith('@') and 'BRA' in text):
            regions.append(region)
            region = []
    if region:
        regions.append(region)
    merged = []
    for reg in regions:
        smp = sum(float(i['# Samples'] or 0) for i in reg)
        if merged and (len(reg) < 12 and smp < 0.01 * total):
            merged[-1] += reg
        else:
            merged.append(list(reg))
    for reg in merged:
        smp = sum(float(i['# Samples'] or 0) for i in reg)
        if smp < floor / 100.0 * total:
            continue
        ex = sum(float(i['Instructions Executed'] or 0) for i in reg)
        stalls = {c: sum(float(i[c] or 0) for i in reg) for c in stall_cols}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
        ops = {}
        for i in reg:
            parts = i['Source'].split()
            op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        topops = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
        print('{:>6s}-{:>6s} n={:4d} exec={:10.0f} samples={:6.0f} ({:4.1f}%) {} | {}'.format(
            reg[0]['Address'][-5:], reg[-1]['Address'][-5:], len(reg), ex, smp,
            100 * smp / total,
            ' '.join('{}:{:.0f}'.format(k.replace('stall_', ''), v) for k, v in top),
            ' '.join('{}:{}'.format(k, v) for k, v in topops)))
