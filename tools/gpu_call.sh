mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "udp or c_abi or volume" > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/s4_tests.log
python tools/sanitize_case.py > gpurun_out/s4_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python tools/sanitize_case.py > gpurun_out/s4_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/s4_sanitize_plain.log; tail -5 gpurun_out/r02_memcheck.log; tail -3 gpurun_out/s4_memcheck_stdout.log
