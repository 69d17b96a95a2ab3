"""Live run of the UNMODIFIED reference CPU programs (oracle/_ref, built from /root/reference by
oracle/Makefile) against the oracle on a sector the golden vectors do not cover.  Skipped where
the reference binaries do not exist (they are built only in the container that has the checkout)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_l2

M, N = 1024, 512


def _ref_bin(oracle, name):
    p = os.path.join(oracle.REF_DIR, name)
    if not os.path.exists(p):
        if os.path.exists("/root/reference/read.cc"):
            oracle.build_ref("/root/reference")
        if not os.path.exists(p):
            pytest.skip("oracle/_ref not built (reference checkout absent)")
    return p


def test_read_cc_live(wrp, oracle, tmp_path):
    exe = _ref_bin(oracle, "read_ref")
    iq16 = wrp.synth.make_sector_int16(M, N, 5, 2)
    spy = tmp_path / "spy"
    spy.mkdir()
    r = subprocess.run([exe], input=wrp.synth.to_text(iq16, 2).encode(), capture_output=True,
                       env=dict(os.environ, WRP_SPY_DIR=str(spy)))
    assert r.returncode == 0 and b"processing:" in r.stdout
    f_m = np.fromfile(spy / f"exec_f64_n{M}_fwd.bin", np.complex128).reshape(N, 2, 2, M)
    f_n = np.fromfile(spy / f"exec_f64_n{N}_fwd.bin", np.complex128).reshape(-1, 2, N)
    lg = np.fromfile(spy / "log10_args.bin").reshape(M // 2, 3)
    o = oracle.chain(wrp.synth.to_planar(iq16, 2).astype(np.complex128), dumps=True)
    assert rel_l2(o.stages["s02_fft1"], np.transpose(f_m[:, :, 1, :], (1, 2, 0))) < 1e-13
    dop = f_n[1:1 + 2 * M].reshape(M, 2, 2, N)[:, :, 1, :]
    s03 = np.conj(np.roll(np.transpose(dop, (1, 0, 2)), N // 2, axis=2))
    s03[:, :, N - 2:] = 0
    assert rel_l2(o.stages["s03_fft2"], s03) < 1e-13
    assert np.allclose(o.stages["power"][0], lg[:, 1], rtol=1e-12)
    assert np.allclose(o.stages["power"][1], lg[:, 2], rtol=1e-12)
    assert np.max(np.abs(o.zdb[1:] - 10 * np.log10(lg[1:, 0]))) < 1e-9


def test_read_single_cc_live(wrp, oracle, tmp_path):
    exe = _ref_bin(oracle, "read_single_ref")
    secs = [wrp.synth.make_sector_int16(M, N, s, 1) for s in (3, 4)]
    wire = np.concatenate([wrp.synth.to_wire(x) for x in secs])
    inp = tmp_path / "wire.bin"
    wire.tofile(inp)
    r = subprocess.run([exe], stdout=subprocess.DEVNULL,
                       env=dict(os.environ, WRP_FAKE_UDP_IN=str(inp), WRP_FAKE_UDP_OUT=str(tmp_path / "udp")))
    assert r.returncode == 0
    zb = np.fromfile(str(tmp_path / "udp") + ".19002", np.uint8).reshape(2, 2 + 4 * 512)
    zr = np.fromfile(str(tmp_path / "udp") + ".19003", np.uint8).reshape(2, 2 + 4 * 512)
    assert zb[:, 1].tolist() == [0, 1]  # the reference numbers sectors from 0 itself
    out, _ = oracle.batch_wire_f32(wire, 2, M, N, 3, 2)
    for s in range(2):
        zdb = zb[s, 2:].copy().view(">f4").astype(np.float32)
        zdr = zr[s, 2:].copy().view(">f4").astype(np.float32)
        assert np.isneginf(zdb[0])
        assert np.max(np.abs(out[s, 1:, 0] - zdb[1:])) < 1e-3
        assert np.max(np.abs(out[s, :, 1] - zdr)) < 1e-3
        # the double oracle agrees with the float reference well inside the 0.01 dB product tolerance
        o = oracle.chain(wrp.synth.to_planar(secs[s], 3).astype(np.complex128))
        assert np.max(np.abs(o.zdb[1:] - zdb[1:])) < 1e-3
