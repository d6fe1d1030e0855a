"""Digest of an `ncu --set full` report: one JSON record per launch (duration, tensor-pipe / DRAM / L1TEX / L2 utilisation, DRAM
bytes read + written, registers, grid).  usage: python tools/ncu_summary.py report.ncu-rep > summary.json   (needs ncu, no GPU)"""
import csv
import io
import json
import subprocess
import sys

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def col(*subs):
    for i, h in enumerate(hdr):        # exact name first, then substring
        if h == subs[0]:
            return i
    for i, h in enumerate(hdr):
        if all(s in h for s in subs):
            return i
    return None


C = {'kernel': col('Kernel Name'), 'dur': col('gpu__time_duration.sum'),
     'tensor': col('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
     'rd': hdr.index('dram__bytes_read.sum') if 'dram__bytes_read.sum' in hdr else None,
     'wr': hdr.index('dram__bytes_write.sum') if 'dram__bytes_write.sum' in hdr else None,
     'dram': col('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), 'l1': col('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),
     'lts': col('lts__throughput.avg.pct_of_peak_sustained_elapsed'), 'regs': col('launch__registers_per_thread'),
     'grid': col('launch__grid_size'), 'block': col('launch__block_size'), 'wave': col('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')}


def scale(i, v):
    u = units[i].lower()
    f = float(v.replace(',', ''))
    if u.startswith('gbyte') or u == 'gb':
        return f * 1e9
    if u.startswith('mbyte') or u == 'mb':
        return f * 1e6
    if u.startswith('kbyte') or u == 'kb':
        return f * 1e3
    if u in ('ms', 'msecond'):
        return f * 1e3      # -> us
    if u in ('ns', 'nsecond'):
        return f / 1e3
    if u in ('s', 'second'):
        return f * 1e6
    return f


out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    g = lambda k: (scale(C[k], r[C[k]]) if C[k] is not None and r[C[k]] not in ('', 'n/a') else None)
    name = r[C['kernel']].replace('void ', '').replace('<unnamed>::', '').replace('(int)', '').replace('(bool)', '')
    name = name.split('(CUtensorMap')[0]
    out.append({'kernel': name, 'dur_us': g('dur'), 'tensor_pct': g('tensor'), 'dram_rd_MB': (g('rd') or 0) / 1e6, 'dram_wr_MB': (g('wr') or 0) / 1e6,
                'dram_pct': g('dram'), 'l1tex_pct': g('l1'), 'lts_pct': g('lts'), 'regs': g('regs'), 'smem_wavefronts': g('wave'),
                'grid': r[C['grid']] if C['grid'] is not None else None, 'block': r[C['block']] if C['block'] is not None else None})
json.dump(out, sys.stdout, indent=0)
