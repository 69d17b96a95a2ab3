#!/usr/bin/env python
"""A/B timing of the HBM-resident chain under different wrp_config settings, one child process each.

  python tools/ab.py [--sectors 143] [--reps 20] [--shape 1024x512] "" "chain_impl=1" "doppler_form=1" "WRP_LIB=/path/libwrp_x.so" ...

Every configuration is a space-separated list of key=value pairs: UPPER-CASE keys are environment
variables (WRP_LIB=... selects another build of the library), lower-case keys are wrp_config fields
(chain_impl=1 the two-kind queue, doppler_form=1 the literal transform, input_fmt=1 wire records).  Prints sectors/s (median of --reps launches, CUDA events), the HBM fraction, and the
worst |dZdB|, |dZDR| against the double oracle on the first two sectors."""
import argparse, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(args, cfg):
    overrides = {}
    for kv in cfg.split():
        k, v = kv.split("=", 1)
        if k.isupper():
            os.environ[k] = v
        else:
            overrides[k] = int(v)
    import numpy as np, torch
    import oracle
    wrp = importlib.import_module("weather-radar-processing_b200")
    M, N = (int(x) for x in args.shape.split("x"))
    S = args.sectors
    wire = overrides.get("input_fmt", 0) == 1
    base = wrp.synth.make_batch(M, N, min(S, args.distinct), fmt="planar", distinct=min(S, args.distinct))
    feed = wrp.synth.make_batch(M, N, min(S, args.distinct), fmt="wire", distinct=min(S, args.distinct)) if wire else base
    reps = -(-S // base.shape[0])
    x = torch.from_numpy(np.concatenate([feed] * reps)[:S]).cuda()
    out = torch.empty((S, M // 2, 2), dtype=torch.float32, device="cuda")
    with wrp.RadarChain(0, n_rows_M=M, n_cols_N=N, max_batch=min(S, 64), **overrides) as ch:
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            ch.process_device(x.data_ptr(), S, out.data_ptr(), st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ch.process_device(x.data_ptr(), S, out.data_ptr(), st); b.record()
            b.synchronize(); ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        kernel = ch.chain_kernel
    o = out.cpu().numpy()
    dz = dr = 0.0
    for s in range(min(2, S)):
        ref = oracle.chain(base[s].astype(np.complex128))
        dz = max(dz, float(np.max(np.abs(o[s, 1:, 0] - ref.zdb[1:]))))
        dr = max(dr, float(np.max(np.abs(o[s, :, 1] - ref.zdr))))
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6450.6
    bytes_sector = 3 * M * N * (4 if wire else 8) + (M // 2) * 8
    sps = S / (ms * 1e-3)
    print(json.dumps({"cfg": cfg, "kernel": kernel, "shape": args.shape, "sectors": S, "ms": round(ms, 4), "min_ms": round(ts[0], 4),
                      "sectors_per_s": round(sps), "hbm_frac": round(sps * bytes_sector / (peak * 1e9), 4),
                      "max_dZdB": dz, "max_dZDR": dr}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sectors", type=int, default=143)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--shape", default="1024x512")
    ap.add_argument("--distinct", type=int, default=4)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("cfgs", nargs="*", default=[""])
    args = ap.parse_args()
    if args.child:
        run_one(args, args.cfgs[0])
        return
    for cfg in args.cfgs:  # one child per configuration: WRP_LIB and create-time switches take effect
        r = subprocess.run([sys.executable, __file__, "--child", "--sectors", str(args.sectors), "--reps", str(args.reps),
                            "--shape", args.shape, "--distinct", str(args.distinct), cfg], capture_output=True, text=True)
        sys.stdout.write(r.stdout if r.returncode == 0 else f"FAIL {cfg}: {r.stderr[-600:]}\n")
        dbg = [l for l in r.stderr.splitlines() if l.startswith("[wrp debug]")]
        if dbg:
            sys.stdout.write("    " + dbg[-1] + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
