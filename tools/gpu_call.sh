mkdir -p gpurun_out
timeout 200 python tools/ab.py --child --reps 3 --sectors 6 --shape 4096x1024 --distinct 2 "" > gpurun_out/s7_sanity_4096.log 2>&1; echo "sanity 4096 rc=$?"; tail -1 gpurun_out/s7_sanity_4096.log | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/s7_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s7_tests.log
timeout 300 python tools/ab.py --reps 10 --sectors 64 --shape 4096x1024 --distinct 2 "" 2>&1 | tee gpurun_out/s7_ab.log
timeout 300 python tools/ab.py --reps 10 --sectors 64 --shape 4096x512 --distinct 2 "" 2>&1 | tee -a gpurun_out/s7_ab.log
python tools/ab.py --child --reps 2 --sectors 64 --shape 4096x1024 --distinct 2 "" > gpurun_out/s7_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chain_stream -s 3 -c 1 -o gpurun_out/prof_stream_4096 python tools/ab.py --child --reps 2 --sectors 64 --shape 4096x1024 --distinct 2 "" > gpurun_out/s7_ncu.log 2>&1
