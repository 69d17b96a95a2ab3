mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/s2_tests.log
timeout 300 python tools/ab.py --reps 20 "" "input_fmt=1" 2>&1 | tee gpurun_out/s2_ab.log
python tools/ab.py --child --reps 2 --sectors 143 "" > gpurun_out/s2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chain_stream -s 3 -c 1 -o gpurun_out/prof_stream python tools/ab.py --child --reps 2 --sectors 143 "" > gpurun_out/s2_ncu.log 2>&1
python tools/ab.py --child --reps 2 --sectors 143 "input_fmt=1" > gpurun_out/s2_plain_w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chain_stream -s 3 -c 1 -o gpurun_out/prof_stream_wire python tools/ab.py --child --reps 2 --sectors 143 "input_fmt=1" > gpurun_out/s2_ncu_w.log 2>&1
ls -la gpurun_out/*.ncu-rep
