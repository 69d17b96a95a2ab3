#!/usr/bin/env python
"""Turn the ncu artefacts in gpurun_out/ into the committed summaries under profiles/."""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)
# 1. launch list
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 8]
hdr = rows[0]; ni, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = (r[ni].split("(")[0][:70], r[gi], r[bi]); agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
with open(os.path.join(P, f"{tag}_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 80 python bench.py --steps 5 --warmup 3 --cpu-sample 2 --stress-sectors 0\n")
    f.write("# per-launch device time (cold cache, serialised: compare shares, not absolutes)\n")
    tot = sum(sum(v) for v in agg.values())
    f.write(f"{'kernel':72s} {'grid':>14s} {'block':>12s} {'n':>4s} {'mean_us':>10s} {'share':>7s}\n")
    for (k, g, b), v in agg.items():
        f.write(f"{k:72s} {g:>14s} {b:>12s} {len(v):4d} {sum(v)/len(v)/1e3:10.2f} {sum(v)/tot:7.3f}\n")
    first = [float(r[vi].replace(",", "")) / 1e3 for r in rows[1:9]]
    f.write(f"# launches 1-8 = 3 warm-up + 5 timed steps (one 143-sector launch of {rows[1][ni].split('(')[0]} each, the whole step): "
            f"mean {sum(first)/8:.1f} us;\n# launches 9-24 = the two side-figure forms (chain_forms in the bench line), then the e2e leg "
            f"(decode_wire_kernel + chain kernel per 16-sector piece) and the wire-resident leg.\n")
# 2. full capture summary + hot SASS
rep = os.path.join(G, "prof_chain.ncu-rep")
s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_chain_ncu_summary.txt"), "w").write("# ncu --set full --clock-control none --import-source on -k regex:chain_ -s 4 -c 1 python bench.py --steps 5 --warmup 3 (one 143-sector launch)\n" + s)
h = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), rep, "40"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_chain_hot_sass.txt"), "w").write("# stall reasons, instruction mix and the SASS instructions with most warp-stall samples (ncu --page source)\n" + h)
# 3. traffic for bench.py
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); hd, un, d = rr[0], rr[1], rr[2:]
def col(name):
    i = hd.index(name); sc = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[un[i]]
    return [float(x[i]) * sc for x in d]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
traffic = sum(a + b for a, b in zip(rd, wr)) / len(rd)
kname = d[0][hd.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0].strip()
summ = {kname + "_dram_bytes_per_launch": traffic, "dram_read_bytes": sum(rd)/len(rd), "dram_write_bytes": sum(wr)/len(wr),
        "launches_profiled": len(rd), "source": f"profiles/{tag}_chain_ncu_summary.txt", "sectors_per_launch": 143}
json.dump(summ, open(os.path.join(P, "latest_summary.json"), "w"), indent=1)
bl = open(os.path.join(G, "bench_plain.log")).read().strip().splitlines()[-1]
open(os.path.join(P, f"{tag}_bench_line.json"), "w").write(bl + "\n")
ref = os.path.join(G, "bench_reference.log")
if os.path.exists(ref):
    open(os.path.join(P, f"{tag}_bench_reference_line.json"), "w").write(open(ref).read().strip().splitlines()[-1] + "\n")
print(json.dumps(summ))
