"""Scratch GPU check: fused + staged path vs the oracle on synthetic sectors."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
wrp = importlib.import_module("weather-radar-processing_b200")
synth = wrp.synth

M, N = 1024, 512
S = 3
iq16 = [synth.make_sector_int16(M, N, s, 0) for s in range(S)]
planar = np.stack([synth.to_planar(x, 3) for x in iq16])
wire = np.stack([synth.to_wire(x) for x in iq16])
ref = [oracle.chain(p.astype(np.complex128), dumps=True) for p in planar]

def cmp_products(out, tag):
    for s in range(S):
        zdb, zdr = out[s, :, 0], out[s, :, 1]
        d1 = np.max(np.abs(zdb[1:] - ref[s].zdb[1:])); d2 = np.max(np.abs(zdr - ref[s].zdr))
        print(f"{tag} sector {s}: max|dZdB|={d1:.2e} dB max|dZDR|={d2:.2e} dB gate0={zdb[0]}")

t = time.time()
with wrp.RadarChain(0) as ch:
    print("info chunk", ch.info.chunk_sectors, "sm", ch.info.sm_count, "l2", ch.info.l2_bytes)
    out = ch.process_host(planar, S)
    cmp_products(out, "fused/planar")
    print("launches", ch.launch_count)
with wrp.RadarChain(0, input_fmt=wrp.FMT_WIRE_I16BE) as ch:
    out = ch.process_host(wire, S)
    cmp_products(out, "fused/wire")
with wrp.RadarChain(0, mode=wrp.MODE_STAGED, max_batch=2) as ch:
    out = ch.process_host(planar[:2], 2)
    for s in range(2):
        print("staged", s, np.max(np.abs(out[s,1:,0]-ref[s].zdb[1:])), np.max(np.abs(out[s,:,1]-ref[s].zdr)))
    names = {"01hamm":"s01_hamm","02fft1":"s02_fft1","03fft2":"s03_fft2","04abs":"s04_abs","05fft3":"s05_fft3","06mult":"s06_mult","07conv":"s07_conv","08pow":"s08_pow","power":"power"}
    for st, on in names.items():
        for c in range(3):
            g = ch.dump_stage(st, 1, c).astype(np.complex128 if st in ("01hamm","02fft1","03fft2","05fft3","06mult","07conv") else np.float64)
            r = ref[1].stages[on][c]
            print(f"  stage {st} ch{c}: relL2={np.linalg.norm(g-r)/np.linalg.norm(r):.2e}")
print("total", time.time() - t)
