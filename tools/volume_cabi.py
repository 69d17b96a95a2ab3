#!/usr/bin/env python
"""BASELINE config 4 through the C ABI alone (wrp_volume_*, no torch.distributed, no NCCL): one process,
one host thread and one handle per device, contiguous (elevation, sector) shards, products gathered on
devices[0] by peer copies, one D2H.   python tools/volume_cabi.py [--gpus N] [--steps 3] [--elevations 9]"""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wrp = importlib.import_module("weather-radar-processing_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--elevations", type=int, default=9)
ap.add_argument("--piece", type=int, default=8)
a = ap.parse_args()
M, N, S, E = 1024, 512, 143, a.elevations
U = S * E
base = [wrp.synth.to_wire(wrp.synth.make_sector_int16(M, N, s, 0)) for s in range(4)]
pin = wrp.PinnedBuffer(U * M * N * 12)
view = pin.array.reshape(U, M * N * 12)
for k in range(U):
    view[k] = base[k % 4].reshape(-1)
with wrp.VolumeScan(list(range(a.gpus)), S, E, input_fmt=wrp.FMT_WIRE_I16BE, max_batch=a.piece) as vs:
    vol = vs.process(pin)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        vol = vs.process(pin, vol)
    dt = (time.perf_counter() - t0) / a.steps
ok = bool(np.isfinite(vol[:, 1:]).all() and all(np.allclose(vol[k, 1:], vol[k % 4, 1:], rtol=0, atol=1e-3) for k in range(0, U, 61)))
print(json.dumps({"workload": f"volume scan {E} x {S} wire sectors, wrp_volume_process (C ABI, threads, products stored into the volume on devices[0] by the kernels)",
                  "n_gpus": a.gpus, "ms_per_volume": dt * 1e3, "sectors_per_s": U / dt,
                  "h2d_gbs_per_gpu": U / dt * M * N * 12 / 1e9 / a.gpus, "volume_ok": ok}))
