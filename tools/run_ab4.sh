#!/bin/bash
mkdir -p gpurun_out
WP=$PWD/tools/libwrp_wp.so
timeout 90 python tools/ab.py --child --reps 3 --sectors 20 "WRP_LIB=$WP" > gpurun_out/ab4_sanity.log 2>&1; rc=$?
tail -2 gpurun_out/ab4_sanity.log
if [ $rc -ne 0 ]; then echo "SANITY FAILED rc=$rc"; WP=; fi
timeout 700 python tools/ab.py --reps 30 \
  "WRP_CHAIN=queue" "WRP_DEBUG=16" "WRP_LAG=5 WRP_RING=10" "WRP_LAG=6 WRP_RING=12 WRP_DEBUG=16" "WRP_LAG=6 WRP_RING=12" "WRP_LAG=7 WRP_RING=14" "WRP_LAG=8 WRP_RING=16" \
  "WRP_LIB=$WP WRP_DEBUG=16" "WRP_LIB=$WP" "WRP_LIB=$WP WRP_LAG=5 WRP_RING=10" "WRP_LIB=$WP WRP_LAG=6 WRP_RING=12" "WRP_LIB=$WP WRP_LAG=3 WRP_RING=7" 2>&1 | tee gpurun_out/ab4_default.log
WRP_LAG=6 WRP_RING=12 timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_unified -s 3 -c 1 -f -o gpurun_out/prof_unified3 \
  python tools/ab.py --child --reps 2 "WRP_LAG=6 WRP_RING=12" > gpurun_out/ncu_unified3.log 2>&1
