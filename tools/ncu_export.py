#!/usr/bin/env python
"""Export the figures bench.py and DESIGN.md quote from an .ncu-rep into
profiles/<name>.json and profiles/<name>.txt (read here, no GPU needed).

    python tools/ncu_export.py gpurun_out/prof.ncu-rep profiles/r2_fused_ncu_summary [evals per launch]
"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                     universal_newlines=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
row = data[0]


def get(name, scale=1.0):
    if name not in hdr:
        return None
    idx = hdr.index(name)
    val = float(row[idx]) * scale
    unit = units[idx]
    mult = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0}.get(unit)
    return val * mult if mult else val


summary = {
    'report': rep,
    'kernel': row[hdr.index('Kernel Name')],
    'grid': row[hdr.index('Grid Size')], 'block': row[hdr.index('Block Size')],
    'duration_us': get('gpu__time_duration.sum'),
    'sm_cycles': get('sm__cycles_elapsed.max'),
    'warp_instructions': get('smsp__inst_executed.sum'),
    'issue_active_pct': get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
    'pipe_fma_pct': get('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
    'pipe_alu_pct': get('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
    'pipe_xu_pct': get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
    'pipe_lsu_pct': get('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
    'pipe_tensor_pct': get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
    'smem_wavefronts': get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
    'smem_wavefronts_pct': get(
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
    'smem_bank_conflicts_ld': get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum'),
    'smem_bank_conflicts_st': get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum'),
    'dram_bytes_read': get('dram__bytes_read.sum'),
    'dram_bytes_write': get('dram__bytes_write.sum'),
    'l2_hit_rate_pct': get('lts__t_sector_hit_rate.pct'),
    'registers_per_thread': get('launch__registers_per_thread'),
    'dyn_smem_bytes': get('launch__shared_mem_per_block_dynamic'),
    'warps_active_pct': get('sm__warps_active.avg.pct_of_peak_sustained_active'),
}
summary['dram_bytes_per_launch'] = (summary['dram_bytes_read'] or 0) + \
    (summary['dram_bytes_write'] or 0)
stalls = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(row[i] or 0)
          for i, h in enumerate(hdr)
          if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')}
total = sum(stalls.values()) or 1.0
summary['stall_pct'] = {k: round(100 * v / total, 1)
                        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]}
# Executed FP32 operations from the dynamic instruction mix (source page): per warp-level
# instruction 32 lanes x (FADD/FMUL 1, FFMA 2, FADD2/FMUL2 2, FFMA2 4) FLOP. Transcendentals
# (MUFU) and float64 are not counted; lanes switched off by predicates are (an upper bound
# that is tight here: the kernels' warps run full).
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     stdout=subprocess.PIPE, universal_newlines=True).stdout
shdr, mix = None, {}
for srow in csv.reader(io.StringIO(src)):
    if 'Address' in srow and 'Source' in srow:
        if shdr is not None:
            break                      # first kernel of the report only
        shdr = srow
        continue
    if not shdr or len(srow) != len(shdr):
        continue
    d = dict(zip(shdr, srow))
    parts = d['Source'].split()
    op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
    mix[op] = mix.get(op, 0.0) + float(d['Instructions Executed'] or 0)
weights = {'FADD': 1, 'FMUL': 1, 'FFMA': 2, 'FADD2': 2, 'FMUL2': 2, 'FFMA2': 4}
summary['op_mix_warp_instructions'] = {k: v for k, v in sorted(
    mix.items(), key=lambda kv: -kv[1])[:24]}
summary['executed_fp32_flop_per_launch'] = 32.0 * sum(
    mix.get(op, 0.0) * w for op, w in weights.items())
if len(sys.argv) > 3:
    summary['evals_per_launch'] = int(sys.argv[3])
with open(out + '.json', 'w') as fobj:
    json.dump(summary, fobj, indent=1)
with open(out + '.txt', 'w') as fobj:
    fobj.write('ncu --set full --clock-control none (one launch); source: {}\n'.format(rep))
    for key, val in summary.items():
        fobj.write('{:28s} {}\n'.format(key, val))
print(json.dumps(summary, indent=1))
