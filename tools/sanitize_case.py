#!/usr/bin/env python
"""Small cases for compute-sanitizer (memcheck / synccheck / racecheck), one process:
  2-sector and 12-sector batches of the default shape through the streaming kernel (planar and wire:
  planes cut between CTAs at different tiles, last-arrival combines, hh/vv product hand-over), the
  two-kind queue kernel in both Doppler forms (12 sectors wrap its x2 ring), one 4096 x 512 sector,
  the staged cascade and the two-shard volume entry.  Prints the worst |dZdB| against the oracle."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
wrp = importlib.import_module("weather-radar-processing_b200")
M, N = 1024, 512
iq = [wrp.synth.make_sector_int16(M, N, s, 0) for s in range(2)]
refs = [oracle.chain(wrp.synth.to_planar(x).astype(np.complex128)) for x in iq]
worst = 0.0
def check(out, name):
    global worst
    for i in range(len(out)):
        d = float(np.max(np.abs(out[i][1:, 0] - refs[i % 2].zdb[1:])))
        worst = max(worst, d)
        assert d < 0.01, (name, i, d)
    print(f"{name}: ok ({len(out)} sectors)", flush=True)
for S in (2, 12):
    planar = np.stack([wrp.synth.to_planar(iq[i % 2]) for i in range(S)])
    wire = np.stack([wrp.synth.to_wire(iq[i % 2]) for i in range(S)])
    with wrp.RadarChain(0, max_batch=S) as ch:
        check(ch.process_host(planar, S), f"stream planar S={S}")
    with wrp.RadarChain(0, max_batch=S, input_fmt=wrp.FMT_WIRE_I16BE) as ch:
        check(ch.process_host(wire, S), f"stream wire S={S}")
    with wrp.RadarChain(0, max_batch=S, chain_impl=wrp.CHAIN_QUEUE) as ch:
        check(ch.process_host(planar, S), f"queue energy S={S}")
    with wrp.RadarChain(0, max_batch=S, doppler_form=wrp.DOPPLER_FFT) as ch:
        check(ch.process_host(planar, S), f"queue fft S={S}")
x4 = wrp.synth.to_planar(wrp.synth.make_sector_int16(4096, 512, 0, 0), 2)
r4 = oracle.chain(x4.astype(np.complex128))
with wrp.RadarChain(0, n_rows_M=4096, n_channels=2, max_batch=1) as ch:
    o = ch.process_host(x4[None], 1)
    d = float(np.max(np.abs(o[0][1:, 0] - r4.zdb[1:]))); worst = max(worst, d); assert d < 0.01
    print("stream 4096x512: ok", flush=True)
with wrp.RadarChain(0, mode=wrp.MODE_STAGED, max_batch=1) as ch:
    check(ch.process_host(wrp.synth.to_planar(iq[0])[None], 1), "staged")
with wrp.VolumeScan([0, 0], 2, 2, input_fmt=wrp.FMT_WIRE_I16BE, max_batch=1) as vs:
    check(vs.process(np.stack([wrp.synth.to_wire(iq[i % 2]) for i in range(4)])), "volume 2 shards")
print(f"sanitize_case: all ok, worst |dZdB| {worst:.2e} dB")
