"""Scratch GPU check of the fused M = 4096 path: parity vs the oracle, then device-resident timing."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
wrp = importlib.import_module("weather-radar-processing_b200")
synth = wrp.synth

M = 4096
for N, C in [(1024, 3), (512, 2), (512, 3), (1024, 1)]:
    secs = [synth.to_planar(synth.make_sector_int16(M, N, s, 0), C) for s in range(2)]
    refs = [oracle.chain(x.astype(np.complex128)) for x in secs]
    n = 5  # more sectors than ring slots (3): range tiles wait for Doppler blocks and vice versa
    batch = np.stack([secs[i % 2] for i in range(n)])
    d_in = torch.from_numpy(batch.view(np.float32).reshape(-1)).cuda()
    d_out = torch.zeros((n, M // 2, 2), device="cuda")
    with wrp.RadarChain(0, n_rows_M=M, n_cols_N=N, n_channels=C, max_batch=1) as ch:
        ch.process_device(d_in.data_ptr(), n, d_out.data_ptr(), 0)
        torch.cuda.synchronize()
        out = d_out.cpu().numpy()
        one = ch.process_host(batch[:1], 1)
    print("host path == device path:", np.array_equal(one[0], out[0]))
    del d_in, d_out
    for i in range(n):
        r = refs[i % 2]
        d1 = np.max(np.abs(out[i, 1:, 0] - r.zdb[1:])); d2 = np.max(np.abs(out[i, :, 1] - r.zdr))
        print(f"M={M} N={N} C={C} sector {i}: max|dZdB|={d1:.2e} max|dZDR|={d2:.2e} gate0={out[i,0,0]}", flush=True)

# timing, device resident
N, C = 1024, 3
S = int(os.environ.get("S4096", "32"))
x = synth.to_planar(synth.make_sector_int16(M, N, 0, 0), C)
d_in = torch.from_numpy(np.ascontiguousarray(x).view(np.float32).reshape(-1)).cuda().repeat(S)
d_out = torch.zeros((S, M // 2, 2), device="cuda")
for mode, name in [(wrp.MODE_FUSED, "fused"), (wrp.MODE_STAGED, "staged")]:
    mb = S if mode == wrp.MODE_FUSED else 2
    with wrp.RadarChain(0, n_rows_M=M, n_cols_N=N, n_channels=C, max_batch=mb, mode=mode) as ch:
        ns = S if mode == wrp.MODE_FUSED else 2
        for _ in range(2):
            ch.process_device(d_in.data_ptr(), ns, d_out.data_ptr(), 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ch.process_device(d_in.data_ptr(), ns, d_out.data_ptr(), 0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = ns * (C * M * N * 8 + M // 2 * 8)
        print(f"{name}: {ns} sectors {ms:.3f} ms -> {ns / ms * 1e3:.0f} sectors/s, {by / ms / 1e6:.0f} GB/s algorithmic "
              f"({by / ms / 1e6 / 6450.6 * 100:.1f} % of HBM)", flush=True)
