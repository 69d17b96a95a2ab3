mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/s16_bench_n$N.json 2> gpurun_out/s16_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/s16_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/s16_bench_n$N.json'))
print('N',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'share',round(d['roofline']['kernel_share_of_step'],4),'gathers',d['run']['gathers_in_timed_region'],'sustained',round(d['sustained']['value']),'stress',round(d['stress']['value']),round(d['stress']['hbm_frac'],3),'volume',round(d['volume']['value']),'e2e',round(d['e2e']['value']),round(d['e2e']['h2d_ceiling_gbs'],1),round(d['e2e']['frac_of_h2d_ceiling'],3))
PY
timeout 600 python tools/volume_cabi.py --gpus $N --elevations 9 2>&1 | tail -1 | tee gpurun_out/s16_volume_cabi_n$N.json
