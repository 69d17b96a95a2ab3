#!/usr/bin/env python
"""Run-to-run determinism stress: the same resident batch is processed `--reps` times; every output is compared
bit for bit with the first.  Prints the number of differing launches and where the first difference sits.
  python tools/determinism.py [--sectors 40] [--reps 300] [--shape 1024x512] [key=value ...]   (wrp_config fields / WRP_LIB=...)"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--sectors", type=int, default=40)
ap.add_argument("--reps", type=int, default=300)
ap.add_argument("--shape", default="1024x512")
ap.add_argument("cfg", nargs="*")
a = ap.parse_args()
over = {}
for kv in a.cfg:
    k, v = kv.split("=", 1)
    if k.isupper():
        os.environ[k] = v
    else:
        over[k] = int(v)
import numpy as np, torch
wrp = importlib.import_module("weather-radar-processing_b200")
M, N = (int(x) for x in a.shape.split("x"))
S = a.sectors
wire = over.get("input_fmt", 0) == 1
base = wrp.synth.make_batch(M, N, min(S, 3), fmt="wire" if wire else "planar", distinct=min(S, 3))
x = torch.from_numpy(np.concatenate([base] * (-(-S // base.shape[0])))[:S]).cuda()
out = torch.empty((S, M // 2, 2), dtype=torch.float32, device="cuda")
bad, first_bad = 0, None
with wrp.RadarChain(0, n_rows_M=M, n_cols_N=N, max_batch=min(S, 64), **over) as ch:
    ch.process_device(x.data_ptr(), S, out.data_ptr(), 0)
    torch.cuda.synchronize()
    ref = out.clone()
    for r in range(a.reps):
        out.zero_()
        ch.process_device(x.data_ptr(), S, out.data_ptr(), 0)
        torch.cuda.synchronize()
        same = torch.eq(out.view(torch.int32), ref.view(torch.int32))
        if not bool(same.all()):
            bad += 1
            if first_bad is None:
                idx = torch.nonzero(~same)
                secs = sorted(set(idx[:, 0].tolist()))
                first_bad = {"rep": r, "n_diff": int(idx.shape[0]), "sectors": secs[:8], "gates": sorted(set(idx[:, 1].tolist()))[:16],
                             "max_abs": float((out - ref)[~torch.isinf(ref)].abs().max())}
    kernel = ch.chain_kernel
print(json.dumps({"kernel": kernel, "shape": a.shape, "sectors": S, "reps": a.reps, "cfg": a.cfg, "differing_launches": bad, "first": first_bad}))
