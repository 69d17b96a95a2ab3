#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/ab6_tests.log
timeout 600 python tools/ab.py --reps 30 \
  "" "WRP_LAG=6 WRP_RING=10" "WRP_LAG=6 WRP_RING=11" "WRP_LAG=7 WRP_RING=11" "WRP_LAG=5 WRP_RING=9" "WRP_LAG=4 WRP_RING=8" "WRP_LAG=6 WRP_RING=10 WRP_DEBUG=16" "WRP_LAG=5 WRP_RING=8" "" "WRP_CHAIN=queue" 2>&1 | tee gpurun_out/ab6_default.log
timeout 300 python tools/ab.py --reps 10 --sectors 1 "" "WRP_CHAIN=queue" 2>&1 | tee -a gpurun_out/ab6_default.log
timeout 300 python tools/ab.py --reps 10 --sectors 8 "" "WRP_CHAIN=queue" 2>&1 | tee -a gpurun_out/ab6_default.log
