"""The oracle (CPU restatement) against numpy, the reference's shipped fixtures and the golden
vectors produced by running the UNMODIFIED reference sources (oracle/gen_golden.py)."""
import hashlib

import numpy as np
import pytest

from conftest import assert_stage_close, rel_l2

M, N = 1024, 512


@pytest.mark.parametrize("n", [4, 8, 32, 512, 1024, 4096])
@pytest.mark.parametrize("sign", [-1, 1])
def test_fft_matches_numpy(oracle, n, sign):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    want = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * n  # FFTW BACKWARD is un-normalised
    assert rel_l2(oracle.fft(x, sign), want) < 1e-14
    assert rel_l2(oracle.fft(x.astype(np.complex64), sign), want) < 2e-6


def test_constants_known_answers(oracle):
    """SURVEY.md §7 known answers for read.cc:9-51 at 1024 x 512, 7 taps."""
    ham, c = oracle.hamming(M, N)
    assert c == pytest.approx(-4.159546322e-11, rel=1e-9)
    assert ham[0, 0] == pytest.approx(-2.448291662e-13, rel=1e-9)
    wr = 0.53836 - 0.46164 * np.cos(2 * np.pi * np.arange(M) / (M - 1))
    wd = 0.53836 - 0.46164 * np.cos(2 * np.pi * np.arange(N) / (N - 1))
    assert np.allclose(ham, np.outer(wr, wd) * c, rtol=1e-13)
    g = oracle.ma_taps(7)
    assert np.allclose(g[:4], [0.0044330482, 0.0540055826, 0.2420362294, 0.3990502797], atol=1e-10)
    assert np.allclose(g, g[::-1]) and g.sum() == pytest.approx(1.0, abs=1e-15)
    H = oracle.ma_fft(7, N)
    assert abs(H[256]) == pytest.approx(0.0141228898, rel=1e-8)
    assert np.allclose(H, np.fft.fft(np.r_[g, np.zeros(N - 7)]), atol=1e-15)
    # float variants (read_single.cc:17-60) agree with the double ones to float precision
    ham32, c32 = oracle.hamming(M, N, np.float32)
    assert c32 == pytest.approx(c, rel=1e-6)
    assert np.allclose(ham32, ham, rtol=2e-6)


def test_shipped_fixtures_04_to_08_to_09(oracle, golden_fixtures):
    """out/04abs.cpu.out -> out/08pow.cpu.out -> in/09zdb.altb / out/99result.cpu.out.
    The dumps carry 6 significant digits, so 08 is reproduced to ~1e-5 and ZdB to 1e-4 dB."""
    f = golden_fixtures
    got = oracle.pdop(f["s04_rows"])
    rowmax = np.abs(f["s08_rows"]).max(axis=1, keepdims=True)
    assert (np.abs(got["s08_pow"] - f["s08_rows"]) / rowmax).max() < 2e-5
    assert np.array_equal(f["s08_rows"], f["s08_in_rows"])  # in/08pow.altb == out/08pow.cpu.out
    # stage 09 from the stage-08 row sums
    P = f["s08_rowsum"]
    with np.errstate(divide="ignore"):
        zdb = 10 * np.log10((30.0 * np.arange(512)) ** 2 * 1941.05 * P)
    assert np.isneginf(zdb[0]) and np.isneginf(f["result_99"][0, 0]) and np.isneginf(f["zdb_09"][0])
    assert np.max(np.abs(zdb[1:] - f["result_99"][1:, 0])) < 1e-4
    assert np.array_equal(f["result_99"][1:, 0], f["zdb_09"][1:])
    assert np.array_equal(f["result_99"][:, 1], f["zdr_10"])
    # the taps sum to 1: the row sum of 08 equals the row sum of 04 (the fused kernel relies on it)
    assert np.allclose(f["s08_rowsum"], f["s04_rowsum"], rtol=1e-5)


def test_products_entry_point_matches_fixture(oracle, golden_fixtures):
    f = golden_fixtures
    # build a [512, N] power matrix whose row sums are the fixture's (first column carries the sum)
    pw = np.zeros((512, 4))
    pw[:, 0] = f["s08_rowsum"]
    zdb, _ = oracle.products(pw, pw)
    assert np.isneginf(zdb[0])
    assert np.max(np.abs(zdb[1:] - f["result_99"][1:, 0])) < 1e-4


def _sector0(wrp):
    iq16 = wrp.synth.make_sector_int16(M, N, 0, 0)
    return iq16, wrp.synth.to_wire(iq16)


def test_synthetic_generator_is_stable(wrp, golden_ref_run):
    """The golden vectors are only meaningful if sector 0 is regenerated bit-identically."""
    _, wire = _sector0(wrp)
    assert hashlib.sha256(wire.tobytes()).hexdigest() == str(golden_ref_run["wire_sha256"])


def test_chain_f64_matches_unmodified_read_cc(wrp, oracle, golden_ref_run):
    """Every stage of the restatement against a literal run of read.cc (double, hh+vv)."""
    g = golden_ref_run
    iq16, _ = _sector0(wrp)
    o = oracle.chain(wrp.synth.to_planar(iq16, 2).astype(np.complex128), dumps=True)
    rf, rh, cols = g["rows_full"], g["rows_half"], g["cols"]
    st = o.stages
    for name, got, want in [
        ("01", st["s01_hamm"][:, rf], g["s01_rows"]), ("02", st["s02_fft1"][:, rf], g["s02_rows"]),
        ("02c", st["s02_fft1"][:, :, cols], g["s02_cols"]), ("03", st["s03_fft2"][:, rf], g["s03_rows"]),
        ("03c", st["s03_fft2"][:, :, cols], g["s03_cols"]), ("04", st["s04_abs"][:, rh], g["s04_rows"]),
        ("05", st["s05_fft3"][:, rh], g["s05_rows"]), ("06", st["s06_mult"][:, rh], g["s06_rows"]),
        ("07", st["s07_conv"][:, rh], g["s07_rows"]), ("08", st["s08_pow"][:, rh], g["s08_rows"]),
    ]:
        assert rel_l2(got, want) < 1e-12, name
    assert np.allclose(st["power"][0], g["power_hh"], rtol=1e-12)
    assert np.allclose(st["power"][1], g["power_vv"], rtol=1e-12)
    assert np.isneginf(o.zdb[0]) and np.isneginf(g["zdb"][0])
    assert np.max(np.abs(o.zdb[1:] - g["zdb"][1:])) < 1e-9
    assert np.max(np.abs(o.zdr - g["zdr"])) < 1e-9


def test_chain_f32_matches_unmodified_read_single_cc(wrp, oracle, golden_ref_run):
    """Wire ingest + float chain against read_single.cc's product datagrams
    (2-byte sector id + 512 big-endian floats, read_single.cc:483-492)."""
    g = golden_ref_run
    iq16, wire = _sector0(wrp)
    pkt_b, pkt_r = g["rs_zdb_packet"], g["rs_zdr_packet"]
    assert pkt_b.size == 2 + 4 * 512 and tuple(pkt_b[:2]) == (0, 0)
    zdb_ref = pkt_b[2:].copy().view(">f4").astype(np.float32)
    zdr_ref = pkt_r[2:].copy().view(">f4").astype(np.float32)
    planar = oracle.decode_wire(wire, M, N, 3, np.float32)
    assert np.array_equal(planar, wrp.synth.to_planar(iq16, 3))
    o = oracle.chain(planar, precision="f32")
    assert np.isneginf(zdb_ref[0]) and np.isneginf(o.zdb[0])
    assert np.max(np.abs(o.zdb[1:] - zdb_ref[1:])) < 1e-3
    assert np.max(np.abs(o.zdr - zdr_ref)) < 1e-3
    out, used = oracle.batch_wire_f32(wire, 1, M, N, 3, 2)
    assert used == 2 and np.array_equal(out[0, :, 0], o.zdb) and np.array_equal(out[0, :, 1], o.zdr)


def test_decode_wire_edge_values(oracle):
    """sector.cpp:52-62: big-endian two's complement, extremes and channel order."""
    m, n = 4, 4
    rec = np.zeros((m * n, 6), dtype=">i2")
    rec[0] = [32767, -32768, -1, 1, 256, -256]
    rec[5] = [0x0102, 0x0304, 0x0506, 0x0708, 0x090A, 0x0B0C]
    wire = rec.view(np.uint8).reshape(-1)
    p = oracle.decode_wire(wire, m, n, 3)
    assert p[0, 0, 0] == 32767 - 32768j and p[1, 0, 0] == -1 + 1j and p[2, 0, 0] == 256 - 256j
    assert p[0, 1, 1] == complex(0x0102, 0x0304) and p[2, 1, 1] == complex(0x090A, 0x0B0C)
    p2 = oracle.decode_wire(wire, m, n, 2)
    assert p2.shape == (2, m, n) and np.array_equal(p2, p[:2])


def test_chain_small_sizes_against_numpy(oracle):
    """Run-time sizes: the restated chain equals a direct numpy transcription of Appendix A."""
    rng = np.random.default_rng(7)
    for (m, n, c) in [(8, 4, 2), (64, 32, 3), (16, 128, 1)]:
        x = (rng.integers(-2000, 2000, (c, m, n)) + 1j * rng.integers(-2000, 2000, (c, m, n))).astype(np.complex128)
        o = oracle.chain(x, dumps=True)
        ham, _ = oracle.hamming(m, n)
        x2 = np.fft.fft(x * ham, axis=1)
        y = np.conj(x2 - x2.mean(axis=2, keepdims=True))
        x3 = np.conj(np.roll(np.fft.fft(y, axis=2), n // 2, axis=2))
        x3[:, :, n - 1] = 0
        x3[:, :, n - 2] = 0
        p = np.abs(x3[:, : m // 2]) ** 2
        assert rel_l2(o.stages["s02_fft1"], x2) < 1e-12
        # x3 compares on the row maximum: the DC column is rounding noise of different size
        assert np.max(np.abs(o.stages["s03_fft2"] - x3) / np.abs(x3).max(axis=2, keepdims=True)) < 1e-12
        assert_stage_close(o.stages["s04_abs"], p, "04")
        taps = 7 if n >= 7 else n
        H = np.fft.fft(np.r_[oracle.ma_taps(7), np.zeros(max(n - 7, 0))][:n]) if n >= 7 else None
        if H is not None:
            q = np.real(np.fft.ifft(np.fft.fft(p, axis=2) * H, axis=2))
            assert_stage_close(o.stages["s08_pow"], q, "08")


def test_chain_rejects_bad_sizes(oracle):
    with pytest.raises(ValueError):
        oracle.chain(np.zeros((2, 12, 8), np.complex128))
    with pytest.raises(ValueError):
        oracle.chain(np.zeros((4, 8, 8), np.complex128))


def test_error_metric_and_float_codec(oracle):
    """error.cpp:15-32 (non-finite pairs skipped) and floats.c:3-36."""
    ref = np.array([-np.inf, 1.0, 2.0, 3.0], np.float32)
    got = np.array([-np.inf, 1.0, 2.0, 3.5], np.float32)
    assert oracle.rel_l2(ref, got) == pytest.approx(np.sqrt(0.25 / 14.0), rel=1e-6)
    for v in (0.0, -1.5, 3.14159274, -np.inf, 1e-38):
        b = oracle.ftob(v)
        assert b == np.array([v], ">f4").tobytes()
        back = oracle.btof(b)
        assert back == np.float32(v) or (np.isinf(v) and np.isinf(back))


def test_energy_form_of_the_doppler_stage_on_oracle_spectra(oracle, wrp):
    """The identity the fused kernel's Doppler block uses (Parseval): with Y the un-normalised
    Doppler transform of a range-FFT row x,  sum_{b != 0, N/2-1, N/2-2} |Y_b|^2 =
    N sum_j |x_j|^2 - |Y_0|^2 - |Y_{N/2-1}|^2 - |Y_{N/2-2}|^2, because the mean removal zeroes bin 0
    (read.cc:193-201) and the clip zeroes shifted columns N-1, N-2 (read.cc:222-224).  Checked on
    the oracle's own stage-02 rows against its stage-08 row power, in float32 arithmetic ordered
    like the kernel (16 strided partials per lane, then a tree over 32 lanes)."""
    M, N = 1024, 512
    x = wrp.synth.to_planar(wrp.synth.make_sector_int16(M, N, 2, 1))
    ref = oracle.chain(x.astype(np.complex128), dumps=True)
    x2 = ref.stages["s02_fft1"][:, :M // 2, :].astype(np.complex64)

    def lane_tree_sum(a):
        part = a.reshape(a.shape[:-1] + (N // 32, 32))
        acc = part[..., 0, :]
        for k in range(1, N // 32):
            acc = (acc + part[..., k, :]).astype(a.dtype)
        w = 16
        while w:
            acc = (acc[..., :w] + acc[..., w:2 * w]).astype(a.dtype)
            w //= 2
        return acc[..., 0]

    def abs2(z):
        return (z.real.astype(np.float32) ** 2 + z.imag.astype(np.float32) ** 2).astype(np.float32)

    j = np.arange(N)
    energy = lane_tree_sum(abs2(x2))
    removed = abs2(lane_tree_sum(x2))
    for m in (1, 2):
        tw = np.exp(2j * np.pi * j * (N // 2 - m) / N).astype(np.complex64)
        removed = removed + abs2(lane_tree_sum((x2 * tw).astype(np.complex64)))
    power = (np.float32(N) * energy - removed).astype(np.float64)
    want = np.asarray(ref.stages["power"], np.float64)[:, :M // 2]
    assert np.max(np.abs(power / want - 1)) < 5e-6  # 2e-5 dB
    # the subtraction never cancels more than ~3/4 of the row energy: the Doppler window spreads
    # any line over three bins, only bin 0 and two edge bins are removed
    assert np.max(removed / (np.float32(N) * energy)) < 0.9


@pytest.mark.parametrize("n", [512, 1024])
def test_pruned_dft_decomposition_used_by_the_kernels(n):
    """The algebra behind `dft_bins012` (wrp_fft.cuh) and the per-lane factors of the energy form:
    with j = 32 a + l, Doppler bin N/2 - m of the un-normalised inverse transform is
    sum_l (-1)^l e^{-2 pi i l m / N} D_m(l), D_m(l) = bin m of the forward (N/32)-point DFT over a, and
    bins 0, 1, 2 of that DFT follow from a decimation-in-frequency network pruned to three outputs."""
    rng = np.random.default_rng(n)
    r = n // 32
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    y = np.fft.ifft(x) * n  # Y_b = sum_j x_j e^{+2 pi i j b / N}
    xa = x.reshape(r, 32)   # xa[a][l] = x[32 a + l]
    w = np.exp(-2j * np.pi / r)
    s, d = xa[:r // 2] + xa[r // 2:], xa[:r // 2] - xa[r // 2:]
    g1 = d[:r // 4] - 1j * d[r // 4:]
    b1 = sum(g1[a] * w ** a for a in range(r // 4))
    ss, sd = s[:r // 4] + s[r // 4:], s[:r // 4] - s[r // 4:]
    b0 = ss.sum(axis=0)
    g2 = sd[:r // 8] - 1j * sd[r // 8:]
    b2 = sum(g2[a] * (w * w) ** a for a in range(r // 8))
    full = np.fft.fft(xa, axis=0)  # forward DFT over a, per lane
    assert np.allclose(b0, full[0]) and np.allclose(b1, full[1]) and np.allclose(b2, full[2])
    lane = np.arange(32)
    for m, bm in ((1, b1), (2, b2)):
        t = (-1.0) ** lane * np.exp(-2j * np.pi * lane * m / n)
        assert abs((t * bm).sum() - y[n // 2 - m]) < 1e-9 * np.abs(y).max()
    assert abs(b0.sum() - y[0]) < 1e-9 * np.abs(y).max()
    power = n * np.sum(np.abs(x) ** 2) - abs(y[0]) ** 2 - abs(y[n // 2 - 1]) ** 2 - abs(y[n // 2 - 2]) ** 2
    keep = np.ones(n, bool)
    keep[[0, n // 2 - 1, n // 2 - 2]] = False
    assert abs(power - np.sum(np.abs(y[keep]) ** 2)) < 1e-9 * power
