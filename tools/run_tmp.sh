#!/bin/bash
mkdir -p gpurun_out
B=$PWD/tools/libwrp_bulk.so
timeout 90 python tools/ab.py --child --reps 3 --sectors 20 "WRP_LIB=$B" > gpurun_out/ab8_sanity.log 2>&1; rc=$?
tail -1 gpurun_out/ab8_sanity.log
if [ $rc -ne 0 ]; then echo "SANITY FAILED rc=$rc"; exit 1; fi
WRP_LIB=$B timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_products or deterministic or fewer or batch_edges" 2>&1 | tail -2
timeout 600 python tools/ab.py --reps 30 "" "WRP_LIB=$B" "" "WRP_LIB=$B" "WRP_LIB=$B WRP_LAG=5 WRP_RING=9" "WRP_LIB=$B WRP_DEBUG=16" "WRP_CHAIN=queue" 2>&1 | tee gpurun_out/ab8_default.log
