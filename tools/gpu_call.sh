mkdir -p gpurun_out
timeout 120 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/s10_sanity.log 2>&1; echo "sanity rc=$?"; tail -1 gpurun_out/s10_sanity.log | cut -c1-250
timeout 600 python tools/ab.py --reps 30 "" "debug=64" "" "debug=64" 2>&1 | tee gpurun_out/s10_ab.log
timeout 600 python tools/ab.py --reps 10 --sectors 64 --shape 4096x1024 --distinct 2 "" "debug=64" 2>&1 | tee -a gpurun_out/s10_ab.log
