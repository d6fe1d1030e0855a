#!/usr/bin/env python
"""Condenses what tools/gpu_round2_entry.sh left under gpurun_out/ into one screen: probe verdicts, pytest outcomes per stage,
headline bench lines per switch setting, and the per-kernel timings of every candidate beside the shipped kernel.
usage: python tools/summarize_round2.py [TAG]   (default TAG r2a)"""
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')


def read(name):
    p = os.path.join(OUT, name)
    return open(p, errors='replace').read() if os.path.exists(p) else None


def pytest_outcome(text):
    if text is None:
        return 'missing'
    tail = [ln for ln in text.strip().splitlines() if re.search(r'\b(passed|failed|error|skipped|no tests ran)\b', ln)]
    return tail[-1].strip('= ') if tail else 'no summary line (timeout / crash?): ' + text.strip().splitlines()[-1][:120] if text.strip() else 'empty'


def bench_line(text):
    if not text:
        return None
    for ln in reversed(text.strip().splitlines()):
        if ln.startswith('{'):
            try:
                return json.loads(ln)
            except ValueError:
                pass
    return None


def main(tag):
    print(f'== operand probe ({tag}_operand_probe.txt)')
    probe = read(f'{tag}_operand_probe.txt')
    for ln in (probe or 'missing').splitlines():
        if ln.startswith('E') or 'PASS' in ln or 'FAIL' in ln or 'error' in ln.lower():
            print('  ' + ln.rstrip())
    print(f'== L2 -> shared memory ceiling ({tag}_l2_to_sm.txt)')
    for ln in (read(f'{tag}_l2_to_sm.txt') or 'missing').splitlines():
        print('  ' + ln.rstrip())
    print('== pytest stages')
    for stage, label in (('pytest', 'default -m gpu suite'), ('v3_pytest', 'UAD_TC_V3=1'), ('unverified_pytest', 'UAD_UNVERIFIED=1'),
                         ('swz_pytest', 'UAD_TC_V2=21'), ('wgrad2_pytest', 'UAD_WGRAD_V2=1'), ('ss_pytest', 'UAD_TC_SS=7'),
                         ('ss_rawhi_pytest', 'UAD_TC_SS=15'), ('halo_pytest', 'UAD_TC_HALO=1')):
        print(f'  {label:24s} {pytest_outcome(read(f"{tag}_{stage}.log"))}')
    print('== bench lines (slices/s, ms/step, e2e, dominant kernel frac)')
    for path in sorted(glob.glob(os.path.join(OUT, f'{tag}_bench*.json'))):
        b = bench_line(open(path).read())
        name = os.path.basename(path)[len(tag) + 1:-5]
        if b is None:
            print(f'  {name:16s} no JSON line')
            continue
        r = b.get('roofline') or {}
        print(f'  {name:16s} {b.get("value", 0):9.0f}  {b.get("ms_per_step", 0):6.3f} ms  e2e {(b.get("e2e") or {}).get("value", 0):9.0f}  '
              f'{r.get("kernel", "?")} frac {r.get("frac", 0):.3f}')
    print('== per-kernel timings, first column of every time_tc file (ms; the kernel as it would ship under that switch)')
    tables = {}
    for path in sorted(glob.glob(os.path.join(OUT, f'{tag}_time_tc_*.txt'))):
        col = {}
        for ln in open(path, errors='replace'):
            m = re.match(r'(.{45}) 0:([0-9.]+)ms', ln)
            if m:
                col[m.group(1).strip()] = float(m.group(2))
        tables[os.path.basename(path)[len(tag) + 9:-4]] = col
    names = []
    for col in tables.values():
        names += [k for k in col if k not in names]
    if tables:
        print('  ' + ' ' * 42 + ''.join(f'{k:>10s}' for k in tables))
        for n in names:
            print(f'  {n[:42]:42s}' + ''.join(f'{tables[k].get(n, float("nan")):10.3f}' for k in tables))
    print('== per-layer tables (ms: all probed kernels | conv fwd + dgrad | filter gradients | everything else)')
    for path in sorted(glob.glob(os.path.join(OUT, f'{tag}_layers*.json'))):
        try:
            t = json.load(open(path))['table']
        except (ValueError, KeyError):
            continue
        w = sum(r['ms'] for r in t if 'wgrad' in r['op'])
        conv = sum(r['ms'] for r in t if re.search(r'(fwd|dgrad)$', r['op']) and 'conv' in r['op'].split(':')[1])
        tot = sum(r['ms'] for r in t)
        print(f'  {os.path.basename(path)[len(tag) + 1:-5]:18s} total {tot:6.3f}  conv fwd+dgrad {conv:6.3f}  wgrad {w:6.3f}  other {tot - conv - w:6.3f}')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r2a')
