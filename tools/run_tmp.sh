timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "4096" 2>&1 | tail -3
for cfg in "A=1" "WRP_TILE_COLS_4096=4" "WRP_EVICT_FIRST=1" "WRP_LAG=2 WRP_RING=4" "A=2" "WRP_TILE_COLS_4096=4 B=2"; do echo "$cfg: $(env $cfg timeout 120 python tools/time4096.py 1024 3 32 5 2>&1 | tail -1)"; done
echo "N=512: $(timeout 120 python tools/time4096.py 512 3 64 5 2>&1 | tail -1)"
echo "N=512 T4: $(WRP_TILE_COLS_4096=4 timeout 120 python tools/time4096.py 512 3 64 5 2>&1 | tail -1)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_persistent -s 2 -c 1 -o gpurun_out/prof_4096_t2 python tools/time4096.py 1024 3 32 2 > gpurun_out/ncu_4096.log 2>&1; tail -2 gpurun_out/ncu_4096.log
