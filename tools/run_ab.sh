#!/bin/bash
# One gpurun call: quick parity subset, A/B of the Doppler forms and queue settings, one ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "energy or fused_products or deterministic or 4096_range or dwell or extreme or fewer" 2>&1 | tail -5 | tee gpurun_out/ab_tests.log
LATE=$PWD/tools/libwrp_late.so
timeout 600 python tools/ab.py --reps 30 \
  "WRP_DOPPLER=fft" "WRP_DOPPLER=energy" "WRP_DOPPLER=fft" "WRP_DOPPLER=energy" \
  "WRP_DOPPLER=energy WRP_LIB=$LATE" \
  "WRP_DOPPLER=energy WRP_LAG=3 WRP_RING=7" "WRP_DOPPLER=energy WRP_LAG=2 WRP_RING=6" \
  "WRP_DOPPLER=energy WRP_LAG=5 WRP_RING=9" "WRP_DOPPLER=energy WRP_LAG=3 WRP_RING=8" \
  "WRP_DOPPLER=energy WRP_EVICT_FIRST=0" 2>&1 | tee gpurun_out/ab_default.log
timeout 400 python tools/ab.py --reps 10 --sectors 32 --distinct 2 --shape 4096x1024 \
  "WRP_DOPPLER=fft" "WRP_DOPPLER=energy" "WRP_DOPPLER=energy WRP_EVICT_FIRST=1" 2>&1 | tee gpurun_out/ab_stress.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_persistent -s 3 -c 1 -f -o gpurun_out/prof_energy \
  python tools/ab.py --child --reps 2 "WRP_DOPPLER=energy" > gpurun_out/ncu_energy.log 2>&1
ls -la gpurun_out | tail -8
