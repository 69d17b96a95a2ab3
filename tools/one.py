import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
wrp = importlib.import_module("weather-radar-processing_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
C = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x = wrp.synth.make_batch(1024, N, S, fmt="planar", distinct=2)[:, :C].copy()
try:
    with wrp.RadarChain(0, n_cols_N=N, n_channels=C) as ch:
        out = ch.process_host(x, S)
    print("ok", S, N, C, out[0, 1:3].tolist())
except Exception as e:
    print("FAIL", S, N, C, e)
