#!/bin/bash
# One gpurun call: sanity (with a short timeout: a hang must not eat the budget), parity subset,
# A/B of the chain kernels, one ncu capture.
mkdir -p gpurun_out
WC=$PWD/tools/libwrp_wc.so
timeout 90 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/ab2_sanity.log 2>&1; rc=$?
cat gpurun_out/ab2_sanity.log | tail -3
if [ $rc -ne 0 ]; then echo "SANITY FAILED rc=$rc"; timeout 60 python tools/ab.py --child --reps 3 --sectors 1 "" 2>&1 | tail -3; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "energy or fused_products or deterministic or fewer or extreme or batch_edges or submit or volume or golden or native" 2>&1 | tail -5 | tee gpurun_out/ab2_tests.log
timeout 600 python tools/ab.py --reps 30 \
  "WRP_CHAIN=queue" "" "WRP_LIB=$WC" "WRP_CHAIN=queue" "" "WRP_LIB=$WC" \
  "WRP_LAG=3 WRP_RING=7" "WRP_LAG=2 WRP_RING=6" "WRP_LAG=5 WRP_RING=9" "WRP_LAG=6 WRP_RING=9" "WRP_LAG=1 WRP_RING=4" \
  "WRP_EVICT_FIRST=0" 2>&1 | tee gpurun_out/ab2_default.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_unified -s 3 -c 1 -f -o gpurun_out/prof_unified \
  python tools/ab.py --child --reps 2 "" > gpurun_out/ncu_unified.log 2>&1
ls -la gpurun_out | tail -6
