#!/usr/bin/env python
"""Stall-reason totals and the hottest SASS instructions of the first kernel in an .ncu-rep
(ncu --page source --csv).  usage: ncu_stalls.py rep [top_n]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ci['# Samples']].isdigit()]
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
tot = collections.Counter()
for r in data:
    for k in stalls:
        tot[k] += int(r[ci[k]] or 0)
T = sum(tot.values())
print('samples', T)
print(' '.join(f"{k[6:]}:{v / T * 100:.1f}%" for k, v in tot.most_common() if v))
wf = sum(int(r[ci['L1 Wavefronts Shared']] or 0) for r in data)
print('shared wavefronts', wf, 'ideal', sum(int(r[ci['L1 Wavefronts Shared Ideal']] or 0) for r in data))
byop = collections.Counter(); wfop = collections.Counter(); exop = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ci['Source']]); op = m.group(2) if m else '?'
    byop[op.split('.')[0]] += int(r[ci['# Samples']]); wfop[op] += int(r[ci['L1 Wavefronts Shared']] or 0)
    exop[op.split('.')[0]] += int(r[ci['Instructions Executed']])
print('samples by opcode:', ' '.join(f"{o}:{n / T * 100:.1f}%" for o, n in byop.most_common(16)))
E = sum(exop.values())
print('executed by opcode:', ' '.join(f"{o}:{n / E * 100:.1f}%" for o, n in exop.most_common(24)))
print('shared wavefronts by opcode:', ' '.join(f"{o}:{n}" for o, n in wfop.most_common(8) if n))
print('--- hottest instructions (index, samples, executed, top stall, source)')
order = sorted(range(len(data)), key=lambda i: -int(data[i][ci['# Samples']]))[:topn]
for i in sorted(order):
    r = data[i]
    top = max(stalls, key=lambda k: int(r[ci[k]] or 0))
    print(i, r[ci['# Samples']], r[ci['Instructions Executed']], top[6:], r[ci['Source']].strip()[:90])
