mkdir -p gpurun_out
timeout 600 python tools/ab.py --reps 10 --sectors 64 --shape 4096x1024 --distinct 2 "" "debug=16384" "debug=32768" "debug=65536" 2>&1 | tee gpurun_out/s8_ab.log
timeout 600 python tools/ab.py --reps 30 "" "debug=16384" "debug=32768" "debug=65536" "" 2>&1 | tee -a gpurun_out/s8_ab.log
