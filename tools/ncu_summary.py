#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics and the hottest source
lines by sampled stall reason.   python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         stdout=subprocess.PIPE, universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for num, row in enumerate(data):
        print('== launch', num, row[hdr.index('Kernel Name')][:60])
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print('  {:78s} {:>16s} {}'.format(key, row[i], units[i]))
        stalls = [(float(row[i] or 0), h) for i, h in enumerate(hdr)
                  if h.startswith('smsp__pcsamp_warps_issue_stalled_') and
                  not h.endswith('_not_issued')]
        total = sum(v for v, _ in stalls) or 1
        print('  stall samples:', ', '.join('{} {:.1f}%'.format(
            h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / total)
            for v, h in sorted(stalls, reverse=True)[:9]))


def source(rep, top=40):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                          '--print-source', 'sass'],
                         stdout=subprocess.PIPE, universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    per_line = defaultdict(float)
    insts = []
    for row in rows:
        if 'Source' in row and '# Samples' in ' '.join(row) or (hdr is None and 'Address' in row):
            hdr = row
            continue
        if hdr is None or len(row) != len(hdr):
            continue
        insts.append(dict(zip(hdr, row)))
    return hdr, insts


if __name__ == '__main__':
    raw(sys.argv[1])
