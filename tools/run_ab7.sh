#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/ab7_sanity.log 2>&1; rc=$?
tail -1 gpurun_out/ab7_sanity.log
if [ $rc -ne 0 ]; then echo "SANITY FAILED rc=$rc"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "energy or fused_products or deterministic or fewer or batch_edges or volume" 2>&1 | tail -3 | tee gpurun_out/ab7_tests.log
timeout 600 python tools/ab.py --reps 30 \
  "" "WRP_LAG=7 WRP_RING=10" "WRP_LAG=7 WRP_RING=11" "WRP_LAG=6 WRP_RING=9" "WRP_LAG=5 WRP_RING=9" "WRP_LAG=5 WRP_RING=8" "WRP_LAG=8 WRP_RING=11" "WRP_DEBUG=16" "WRP_LAG=7 WRP_RING=10 WRP_DEBUG=16" "" 2>&1 | tee gpurun_out/ab7_default.log
