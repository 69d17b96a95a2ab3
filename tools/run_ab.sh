#!/bin/bash
# One gpurun call that A/Bs the default library against every experimental build found in tools/:
#   make variants pair && gpurun --timeout 1200 -- 'bash tools/run_ab.sh'
# For each tools/libwrp_<name>.so: a sanity launch under a short timeout (a hang must not eat the
# budget), a parity subset of the GPU tests with WRP_LIB pointing at it, then timing next to the
# default build in the same process tree (tools/ab.py).  Results land in gpurun_out/ab_*.log.
mkdir -p gpurun_out
: > gpurun_out/ab_default.log
CFGS=("" "WRP_CHAIN=queue")
for lib in tools/libwrp_*.so; do
  [ -e "$lib" ] || continue
  name=$(basename "$lib" .so); L=$PWD/$lib
  timeout 90 python tools/ab.py --child --reps 3 --sectors 20 "WRP_LIB=$L" > gpurun_out/ab_sanity_$name.log 2>&1; rc=$?
  tail -1 gpurun_out/ab_sanity_$name.log | cut -c1-200
  if [ $rc -ne 0 ]; then echo "$name: SANITY FAILED rc=$rc (skipped)"; continue; fi
  if WRP_LIB=$L timeout 300 python -m pytest tests/test_gpu_parity.py -x -q \
       -k "fused_products or deterministic or fewer or batch_edges or volume or near_nyquist" > gpurun_out/ab_tests_$name.log 2>&1; then
    echo "$name: parity subset ok"; CFGS+=("WRP_LIB=$L")
  else
    echo "$name: PARITY FAILED (skipped)"; tail -5 gpurun_out/ab_tests_$name.log
  fi
done
timeout 900 python tools/ab.py --reps 30 "${CFGS[@]}" "" 2>&1 | tee -a gpurun_out/ab_default.log
