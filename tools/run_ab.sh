#!/bin/bash
# One gpurun call: sanity with a short timeout (a hang must not eat the budget), a parity subset,
# A/B timing of the chain forms / queue settings in one process tree, one ncu capture.
#   gpurun --timeout 900 -- 'bash tools/run_ab.sh'
mkdir -p gpurun_out
timeout 90 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/ab_sanity.log 2>&1; rc=$?
tail -1 gpurun_out/ab_sanity.log
if [ $rc -ne 0 ]; then echo "SANITY FAILED rc=$rc"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "energy or fused_products or deterministic or fewer or batch_edges or volume" 2>&1 | tail -3 | tee gpurun_out/ab_tests.log
timeout 600 python tools/ab.py --reps 30 \
  "" "WRP_CHAIN=queue" "WRP_DOPPLER=fft" "WRP_LAG=5 WRP_RING=9" "WRP_LAG=5 WRP_RING=10" "WRP_LAG=6 WRP_RING=11" "WRP_DEBUG=16" "" 2>&1 | tee gpurun_out/ab_default.log
timeout 400 python tools/ab.py --reps 10 --sectors 32 --distinct 2 --shape 4096x1024 "WRP_DOPPLER=fft" "" 2>&1 | tee gpurun_out/ab_stress.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_ -s 3 -c 1 -f -o gpurun_out/prof_ab \
  python tools/ab.py --child --reps 2 "" > gpurun_out/ncu_ab.log 2>&1
