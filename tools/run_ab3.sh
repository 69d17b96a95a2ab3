#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/ab3_sanity.log 2>&1; rc=$?
tail -2 gpurun_out/ab3_sanity.log
if [ $rc -ne 0 ]; then echo "SANITY FAILED rc=$rc"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "energy or fused_products or deterministic or fewer or batch_edges or volume" 2>&1 | tail -3 | tee gpurun_out/ab3_tests.log
timeout 600 python tools/ab.py --reps 30 \
  "WRP_CHAIN=queue" "" "WRP_DEBUG=16" "WRP_LAG=3 WRP_RING=7 WRP_DEBUG=16" "WRP_LAG=3 WRP_RING=6" "WRP_LAG=5 WRP_RING=10 WRP_DEBUG=16" "WRP_LAG=6 WRP_RING=12" \
  "WRP_LAG=2 WRP_RING=5 WRP_DEBUG=16" "WRP_LAG=4 WRP_RING=9" "WRP_CHAIN=queue" "" 2>&1 | tee gpurun_out/ab3_default.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_unified -s 3 -c 1 -f -o gpurun_out/prof_unified2 \
  python tools/ab.py --child --reps 2 "" > gpurun_out/ncu_unified2.log 2>&1
