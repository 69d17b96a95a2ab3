#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference checkout (run in the build container).

TEST INFRASTRUCTURE ONLY.  Two sources, both the reference's own:

1. ``fixtures.npz`` — the shipped stage dumps out/04abs.cpu.out, out/08pow.cpu.out,
   in/09zdb.altb, in/10zdr.altb, out/99result.cpu.out (hh channel, 512 gates), reduced to
   a row subset for the two 3 MB matrices plus the full row sums so the commit stays small.
2. ``ref_run_sector0.npz`` — a literal run of the UNMODIFIED reference sources on synthetic
   sector (sector 0, elevation 0): read.cc (double, hh+vv) observed stage by stage through
   oracle/shim/fftw3.h, and read_single.cc (float, hh+vv+vh, wire ingest) through its product
   datagrams.  Stage matrices are stored on a row/column subset; products in full.

Usage:  python oracle/gen_golden.py [/root/reference]
"""
from __future__ import annotations

import hashlib
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

synth = importlib.import_module("weather-radar-processing_b200.synth")
dumpio = importlib.import_module("weather-radar-processing_b200.dumpio")

M, N = 1024, 512
ROWS_FULL = np.array([0, 1, 2, 3, 255, 256, 511, 512, 513, 1023])
ROWS_HALF = np.array([0, 1, 2, 3, 100, 255, 256, 511])
COLS = np.array([0, 1, 255, 256, 510, 511])


def fixtures(ref: str, out_dir: str) -> None:
    a04 = dumpio.read_dump(os.path.join(ref, "out/04abs.cpu.out"))
    a08 = dumpio.read_dump(os.path.join(ref, "out/08pow.cpu.out"))
    a08_in = dumpio.read_dump(os.path.join(ref, "in/08pow.altb"))
    res = np.genfromtxt(os.path.join(ref, "out/99result.cpu.out"))
    zdb_in = np.genfromtxt(os.path.join(ref, "in/09zdb.altb"))
    zdr_in = np.genfromtxt(os.path.join(ref, "in/10zdr.altb"))
    rows = np.arange(0, 512, 16)
    np.savez_compressed(
        os.path.join(out_dir, "fixtures.npz"),
        rows=rows, s04_rows=a04[rows], s08_rows=a08[rows], s08_in_rows=a08_in[rows],
        s04_rowsum=a04.sum(axis=1), s08_rowsum=a08.sum(axis=1),
        result_99=res, zdb_09=zdb_in, zdr_10=zdr_in,
        source="out/04abs.cpu.out out/08pow.cpu.out in/08pow.altb out/99result.cpu.out in/09zdb.altb in/10zdr.altb",
    )
    print("fixtures.npz written", a04.shape, a08.shape)


def ref_run(ref: str, out_dir: str) -> None:
    oracle.build_ref(ref)
    iq16 = synth.make_sector_int16(M, N, 0, 0)
    wire = synth.to_wire(iq16)
    sha = hashlib.sha256(wire.tobytes()).hexdigest()
    tmp = tempfile.mkdtemp()
    spy = os.path.join(tmp, "spy")
    os.makedirs(spy)
    # read.cc: double, hh + vv, stdin text (read.cc:105-123)
    r = subprocess.run([os.path.join(oracle.REF_DIR, "read_ref")], input=synth.to_text(iq16, 2).encode(),
                       capture_output=True, env=dict(os.environ, WRP_SPY_DIR=spy))
    assert r.returncode == 0, r.stderr
    f_m = np.fromfile(os.path.join(spy, f"exec_f64_n{M}_fwd.bin"), np.complex128).reshape(-1, 2, M)
    f_n = np.fromfile(os.path.join(spy, f"exec_f64_n{N}_fwd.bin"), np.complex128).reshape(-1, 2, N)
    b_n = np.fromfile(os.path.join(spy, f"exec_f64_n{N}_bwd.bin"), np.complex128).reshape(-1, 2, N)
    lg = np.fromfile(os.path.join(spy, "log10_args.bin")).reshape(M // 2, 3)
    # call order: range executes are (column j, channel) ; read.cc:155-181
    rng = f_m.reshape(N, 2, 2, M)                       # [j][ch][in/out][i]
    s01 = np.transpose(rng[:, :, 0, :], (1, 2, 0))      # [ch][i][j]
    s02 = np.transpose(rng[:, :, 1, :], (1, 2, 0))
    fft_ma = f_n[0, 1]                                  # read.cc:97 first N-point execute
    dop = f_n[1:1 + 2 * M].reshape(M, 2, 2, N)          # [i][ch][in/out][j] read.cc:190-254
    s03 = np.conj(np.roll(np.transpose(dop[:, :, 1, :], (1, 0, 2)), N // 2, axis=2))
    s03[:, :, N - 1] = 0
    s03[:, :, N - 2] = 0
    pd = f_n[1 + 2 * M:].reshape(2, M // 2, 2, N)       # [ch][i][in/out][j] read.cc:281-325
    s04 = pd[:, :, 0, :].real
    s05 = pd[:, :, 1, :]
    bw = b_n.reshape(2, M // 2, 2, N)
    s06, s07 = bw[:, :, 0, :], bw[:, :, 1, :]
    s08 = s07.real / N
    with np.errstate(divide="ignore"):
        zdb = 10 * np.log10(lg[:, 0])
        zdr = 10 * (np.log10(lg[:, 1]) - np.log10(lg[:, 2]))
    # read_single.cc: float, three channels, wire in, product datagrams out
    inp = os.path.join(tmp, "wire.bin")
    wire.tofile(inp)
    outb = os.path.join(tmp, "udp")
    r = subprocess.run([os.path.join(oracle.REF_DIR, "read_single_ref")], stdout=subprocess.DEVNULL,
                       env=dict(os.environ, WRP_FAKE_UDP_IN=inp, WRP_FAKE_UDP_OUT=outb))
    assert r.returncode == 0
    zdb_pkt = np.fromfile(outb + ".19002", np.uint8)
    zdr_pkt = np.fromfile(outb + ".19003", np.uint8)
    np.savez_compressed(
        os.path.join(out_dir, "ref_run_sector0.npz"),
        M=M, N=N, sector=0, elevation=0, wire_sha256=sha,
        rows_full=ROWS_FULL, rows_half=ROWS_HALF, cols=COLS,
        s01_rows=s01[:, ROWS_FULL], s02_rows=s02[:, ROWS_FULL], s02_cols=s02[:, :, COLS],
        s03_rows=s03[:, ROWS_FULL], s03_cols=s03[:, :, COLS],
        s04_rows=s04[:, ROWS_HALF], s05_rows=s05[:, ROWS_HALF], s06_rows=s06[:, ROWS_HALF],
        s07_rows=s07[:, ROWS_HALF], s08_rows=s08[:, ROWS_HALF],
        s04_rowsum=s04.sum(axis=2), fft_ma=fft_ma,
        power_hh=lg[:, 1], power_vv=lg[:, 2], z_arg=lg[:, 0], zdb=zdb, zdr=zdr,
        rs_zdb_packet=zdb_pkt, rs_zdr_packet=zdr_pkt,
        source="read.cc + read_single.cc compiled unmodified against oracle/shim (see oracle/Makefile)",
    )
    print("ref_run_sector0.npz written; wire sha256", sha)


if __name__ == "__main__":
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    fixtures(ref_root, out)
    ref_run(ref_root, out)
