mkdir -p gpurun_out
timeout 120 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/s1_sanity.log 2>&1; echo "sanity rc=$?"; tail -2 gpurun_out/s1_sanity.log | cut -c1-400
timeout 120 python tools/ab.py --child --reps 3 --sectors 20 "input_fmt=1" > gpurun_out/s1_sanity_wire.log 2>&1; echo "sanity wire rc=$?"; tail -2 gpurun_out/s1_sanity_wire.log | cut -c1-400
timeout 200 python tools/ab.py --child --reps 3 --sectors 6 --shape 4096x1024 --distinct 2 "" > gpurun_out/s1_sanity_4096.log 2>&1; echo "sanity 4096 rc=$?"; tail -2 gpurun_out/s1_sanity_4096.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/s1_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/s1_tests.log
timeout 600 python tools/ab.py --reps 30 "" "chain_impl=1" "input_fmt=1" "" 2>&1 | tee gpurun_out/s1_ab.log
timeout 600 python tools/ab.py --reps 10 --sectors 64 --shape 4096x1024 --distinct 2 "" "chain_impl=1" 2>&1 | tee -a gpurun_out/s1_ab.log
