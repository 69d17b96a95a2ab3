mkdir -p gpurun_out
timeout 120 python tools/ab.py --child --reps 3 --sectors 20 "" > gpurun_out/s3_sanity.log 2>&1; echo "sanity rc=$?"; tail -1 gpurun_out/s3_sanity.log | cut -c1-300
timeout 200 python tools/ab.py --child --reps 3 --sectors 6 --shape 4096x1024 --distinct 2 "" > gpurun_out/s3_sanity_4096.log 2>&1; echo "sanity 4096 rc=$?"; tail -1 gpurun_out/s3_sanity_4096.log | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s3_tests.log
timeout 300 python tools/ab.py --reps 30 "" "input_fmt=1" "" 2>&1 | tee gpurun_out/s3_ab.log
timeout 300 python tools/ab.py --reps 10 --sectors 64 --shape 4096x1024 --distinct 2 "" 2>&1 | tee -a gpurun_out/s3_ab.log
timeout 300 python tools/ab.py --reps 10 --sectors 100 --shape 1024x1024 --distinct 2 "" 2>&1 | tee -a gpurun_out/s3_ab.log
timeout 300 tools/micro/tile_load 2>&1 | tee gpurun_out/tile_load.txt
