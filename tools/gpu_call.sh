mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/s15_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s15_tests.log
timeout 600 python tools/sanitize_case.py > gpurun_out/s15_small_cases.log 2>&1; echo "small cases rc=$?"; tail -2 gpurun_out/s15_small_cases.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
CMD="python bench.py --steps 2 --warmup 3 --cpu-sample 1 --sustain-seconds 0.001 --stress-sectors 8 --volume-steps 0 --skip-reference-gpu"
$CMD > gpurun_out/s15_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/s15_ncu1.log 2>&1
timeout 900 python bench.py > gpurun_out/s15_bench.json 2> gpurun_out/s15_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s15_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s15_bench.json'))
print({k:d[k] for k in ('value','chain_hbm_frac')}, 'sustained', d['sustained']['value'], 'stress', d['stress']['hbm_frac'], 'volume', d['volume']['value'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_h2d_ceiling'], 'wire', d['wire_resident']['value'], 'refgpu', d['reference_gpu'].get('value'))
PY
