#!/bin/bash
# final validation of the committed tree: full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/bench_plain.log 2>&1; tail -1 gpurun_out/bench_plain.log | cut -c1-300
