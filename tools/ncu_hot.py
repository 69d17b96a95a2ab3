#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` of a .ncu-rep."""
import csv, subprocess, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu','-i',path,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ci = {k:i for i,k in enumerate(hdr)}
stall_cols = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r[ci['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {k:0 for k in stall_cols}
for r in data:
    for k in stall_cols: agg[k] += int(r[ci[k]])
print({k:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1]) if v})
idx = sorted(range(len(data)), key=lambda i:-int(data[i][ci['# Samples']]))[:topn]
for i in sorted(idx):
    r = data[i]
    st = {k[6:]:int(r[ci[k]]) for k in stall_cols if int(r[ci[k]])}
    top = sorted(st.items(), key=lambda kv:-kv[1])[:3]
    print(f"{i:5d} {r[ci['# Samples']]:>6s} {r[ci['Instructions Executed']]:>9s}  {r[ci['Source']].strip()[:70]:70s} {top}")
