"""Device-resident timing of the fused chain at the stress shape (M = 4096): tools/time4096.py [N] [C] [S] [reps]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
wrp = importlib.import_module("weather-radar-processing_b200")
synth = wrp.synth
M = int(os.environ.get("M4096", "4096"))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = int(sys.argv[3]) if len(sys.argv) > 3 else 32
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
x = synth.to_planar(synth.make_sector_int16(M, N, 0, 0), C)
d_in = torch.from_numpy(np.ascontiguousarray(x).view(np.float32).reshape(-1)).cuda().repeat(S)
d_out = torch.zeros((S, M // 2, 2), device="cuda")
with wrp.RadarChain(0, n_rows_M=M, n_cols_N=N, n_channels=C, max_batch=1) as ch:
    for _ in range(2):
        ch.process_device(d_in.data_ptr(), S, d_out.data_ptr(), 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ch.process_device(d_in.data_ptr(), S, d_out.data_ptr(), 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = S * (C * M * N * 8 + M // 2 * 8)
    print(f"M={M} N={N} C={C} S={S}: {ms:.3f} ms -> {S / ms * 1e3:.0f} sectors/s, {by / ms / 1e6:.0f} GB/s algorithmic "
          f"({by / ms / 1e6 / 6450.6 * 100:.1f} % of HBM)", flush=True)
