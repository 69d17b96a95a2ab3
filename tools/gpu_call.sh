mkdir -p gpurun_out
timeout 120 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/s14_sanity.log 2>&1; echo "sanity rc=$?"; tail -1 gpurun_out/s14_sanity.log | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/s14_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/s14_tests.log
timeout 300 python tools/ab.py --reps 30 "" "debug=64" "input_fmt=1" "" 2>&1 | tee gpurun_out/s14_ab.log
timeout 300 python tools/ab.py --reps 10 --sectors 100 --shape 1024x1024 --distinct 2 "" "debug=64" 2>&1 | tee -a gpurun_out/s14_ab.log
