#!/usr/bin/env python
"""Executed-instruction mix of the first kernel in an .ncu-rep, and the non-FP/non-memory ("overhead")
instructions grouped by execution count.  usage: ncu_overhead.py rep [min_count]"""
import csv, subprocess, collections, re, sys
rep = sys.argv[1]; minc = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci={k:i for i,k in enumerate(hdr)}
data=[]
for r in rows[2:]:
    if r and r[0]=='Kernel Name': break
    if len(r)==len(hdr) and r[ci['Instructions Executed']].isdigit(): data.append(r)
fp=('FFMA','FADD','FMUL','FADD2','FMUL2','FFMA2'); mem=('LDS','STS','LDGSTS','STG')
tot=collections.Counter(); ops=collections.Counter(); lines=[]
for i,r in enumerate(data):
    c=int(r[ci['Instructions Executed']]); src=r[ci['Source']].strip()
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src); op=m.group(2).split('.')[0] if m else '?'
    kind='fp' if op in fp else 'mem' if op in mem else 'other'
    tot[kind]+=c; ops[op]+=c
    if kind=='other' and c>=minc: lines.append((i,c,r[ci['# Samples']],src[:100]))
T=sum(tot.values()); print('total', T, {k:f'{v/T*100:.1f}%' for k,v in tot.items()})
print(' '.join(f'{o}:{n/T*100:.1f}%' for o,n in ops.most_common(24)))
for l in lines: print(*l)
