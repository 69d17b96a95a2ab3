#!/usr/bin/env python
"""Turn the artefacts tools/profile.sh left in gpurun_out/ into the committed summaries under profiles/:
  <tag>_launches.csv / <tag>_launches.md   ncu launch list of the bench command (shares, not absolutes)
  <tag>_stream_ncu_summary.txt             counters of one chain_stream_kernel launch (ncu --set full)
  <tag>_stream_hot_sass.txt                stall reasons, instruction mix, hottest SASS lines
  latest_summary.json                      DRAM bytes per launch, read by bench.py for roofline.traffic
  <tag>_bench_line.json, <tag>_bench_reference_line.json
usage: python tools/make_profiles.py [tag]   (no GPU needed; ncu -i reads the report)"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
CMD = ("python bench.py --steps 2 --warmup 3 --cpu-sample 1 --sustain-seconds 0.001 --stress-sectors 8 --volume-steps 0 "
       "--skip-reference-gpu")
# 1. launch list
raw = open(os.path.join(G, "launches.csv")).read()
rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
h = rows[0]
ni, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
per = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    per.setdefault(r[ni].split("(")[0], []).append({"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[ui], v / 1e3))
tot = sum(sum(v) for v in per.values())
open(os.path.join(P, f"{tag}_launches.csv"), "w").write(raw)
md = [f"# ncu launch list of the bench command ({tag}, 1 x B200)\n",
      f"Command (exited 0 without ncu immediately before): `{CMD}`",
      "under `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv`.",
      "Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n",
      "| kernel | launches | total us | share of all kernel time | mean us |", "|---|---|---|---|---|"]
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    md.append(f"| {k} | {len(v)} | {sum(v):.1f} | {sum(v) / tot:.3f} | {sum(v) / len(v):.1f} |")
open(os.path.join(P, f"{tag}_launches.md"), "w").write("\n".join(md) + "\n")
# 2. full capture
rep = os.path.join(G, "prof_bench_stream.ncu-rep")
run = lambda tool, *a: subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), rep, *a], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_stream_ncu_summary.txt"), "w").write(
    f"# ncu --set full --clock-control none --import-source on -k regex:chain_stream_kernel -s 4 -c 1 {CMD}\n"
    "# (the same command exited 0 without ncu immediately before; one 143-sector launch of the planar streaming kernel)\n" + run("ncu_summary.py"))
open(os.path.join(P, f"{tag}_stream_hot_sass.txt"), "w").write(
    "# stall reasons, instruction mix and the SASS instructions with most warp-stall samples of the launch above\n" + run("ncu_stalls.py", "40"))
rr = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hd, un, d = rr[0], rr[1], rr[2:]
def col(name):
    i = hd.index(name)
    return float(d[0][i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[un[i]]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
json.dump({"chain_stream_kernel_dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
           "algorithmic_bytes_per_launch": 143 * 12587008, "traffic_over_algorithmic": (rd + wr) / (143 * 12587008),
           "source": f"profiles/{tag}_stream_ncu_summary.txt", "sectors_per_launch": 143},
          open(os.path.join(P, "latest_summary.json"), "w"), indent=1)
# 3. bench lines
for src, dst in (("bench.json", "bench_line.json"), ("bench_ref.json", "bench_reference_line.json")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        lines = [l for l in open(p) if l.startswith("{")]
        if lines:
            open(os.path.join(P, f"{tag}_{dst}"), "w").write(lines[-1])
print("profiles written for", tag)
