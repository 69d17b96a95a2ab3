#!/bin/bash
# Run under gpurun: plain bench first (must exit 0), the reference arm, then the ncu launch list and
# one full capture of the chain kernel.  Outputs land in gpurun_out/.
set -o pipefail
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_plain.log 2>&1 || { tail -5 gpurun_out/bench_plain.log; exit 1; }
tail -1 gpurun_out/bench_plain.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --cpu-sample 2 --stress-sectors 0 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_ -s 4 -c 1 -f -o gpurun_out/prof_chain \
    python bench.py --steps 5 --warmup 3 --cpu-sample 2 --stress-sectors 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -6
