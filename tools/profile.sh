#!/bin/bash
# Run under gpurun (1 GPU): the ncu launch list and one full capture of the planar streaming kernel, each only after
# the same command has exited 0 without ncu; then the default bench line and the reference arm.  Outputs land in
# gpurun_out/; `python tools/make_profiles.py r02` (here, no GPU needed) turns them into the summaries under profiles/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --cpu-sample 1 --sustain-seconds 0.001 --stress-sectors 8 --volume-steps 0 --skip-reference-gpu"
$CMD > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chain_stream_kernel -s 4 -c 1 -f -o gpurun_out/prof_bench_stream $CMD > gpurun_out/ncu_full.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference rc=$?"
