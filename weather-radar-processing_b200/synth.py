"""Deterministic synthetic IQ sectors in both ingest formats.

The reference's benchmark variants feed a ramp ``make_cuFloatComplex(i, j)``
(gpu_1fp.cu:295-312); that input is degenerate for parity work (one spectral line),
so tests and the bench use seeded radar-like sectors instead: point targets, a
zero-Doppler clutter line, receiver noise and a DC offset, quantised to the
reference's int16 full scale (+-16383.5 in K_wind, rpv2.cu:239).

Formats produced:
  wire    uint8[M*N*12]  records ``hhI hhQ vvI vvQ vhI vhQ`` of big-endian int16,
                         record index i*N+j (sector.cpp:52-62, read_single.cc:145-172)
  planar  complex64[3, M, N]  the reference's pinned/device slot layout
                         ``p_iq[j + i*N + ch*M*N]`` (rpv2.cu:379-381)
"""
from __future__ import annotations

import numpy as np

FULL_SCALE_LO = -16384
FULL_SCALE_HI = 16383


def sector_seed(sector: int, elevation: int = 0) -> int:
    return 0x5EC7 + 1009 * elevation + sector


def make_sector_int16(M: int, N: int, sector: int = 0, elevation: int = 0,
                      n_targets: int = 12) -> np.ndarray:
    """int16[3, M, N, 2] (channel, row, col, I/Q)."""
    rng = np.random.Generator(np.random.PCG64(sector_seed(sector, elevation)))
    i = np.arange(M, dtype=np.float64)[:, None]
    j = np.arange(N, dtype=np.float64)[None, :]
    hh = np.zeros((M, N), dtype=np.complex128)
    amp = 10.0 ** rng.uniform(1.0, 3.7, n_targets)
    f_r = rng.uniform(0.01, 0.49, n_targets)
    f_d = rng.uniform(-0.45, 0.45, n_targets)
    phi = rng.uniform(0.0, 1.0, n_targets)
    for a, fr, fd, ph in zip(amp, f_r, f_d, phi):
        hh += a * np.exp(2j * np.pi * (fr * i + fd * j + ph))
    hh += 300.0 * np.exp(2j * np.pi * 0.05 * i)  # zero-Doppler clutter line

    def noise():
        return rng.normal(0.0, 30.0, (M, N)) + 1j * rng.normal(0.0, 30.0, (M, N))

    dc = 40.0 - 25.0j
    chans = (
        hh + noise() + dc,
        0.5 * hh * np.exp(1j * np.pi / 7) + noise() + dc,
        0.05 * hh + noise() + dc,
    )
    out = np.empty((3, M, N, 2), dtype=np.int16)
    for c, x in enumerate(chans):
        out[c, :, :, 0] = np.clip(np.rint(x.real), FULL_SCALE_LO, FULL_SCALE_HI).astype(np.int16)
        out[c, :, :, 1] = np.clip(np.rint(x.imag), FULL_SCALE_LO, FULL_SCALE_HI).astype(np.int16)
    return out


def to_wire(iq16: np.ndarray) -> np.ndarray:
    """int16[3, M, N, 2] -> uint8[M*N*12] big-endian interleaved records."""
    rec = np.ascontiguousarray(np.transpose(iq16, (1, 2, 0, 3)))  # [M, N, 3, 2]
    return rec.astype(">i2").view(np.uint8).reshape(-1)


def to_planar(iq16: np.ndarray, channels: int = 3) -> np.ndarray:
    """int16[3, M, N, 2] -> complex64[channels, M, N]."""
    x = iq16[:channels].astype(np.float32)
    return np.ascontiguousarray(x[..., 0] + 1j * x[..., 1]).astype(np.complex64)


def to_text(iq16: np.ndarray, channels: int = 2) -> str:
    """The stdin text read.cc:105-123 parses: ``a b`` pairs, hh matrix then vv."""
    x = iq16[:channels].reshape(-1)
    return " ".join(map(str, x.tolist())) + "\n"


def make_batch(M: int, N: int, n_sectors: int, *, fmt: str = "wire", first_sector: int = 0,
               elevation: int = 0, distinct: int | None = None) -> np.ndarray:
    """Batch of sectors.  ``distinct`` bounds how many different sectors are
    synthesised (the rest repeat cyclically) so large bench batches build quickly."""
    distinct = n_sectors if distinct is None else max(1, min(distinct, n_sectors))
    base = [make_sector_int16(M, N, first_sector + s, elevation) for s in range(distinct)]
    if fmt == "wire":
        one = [to_wire(b) for b in base]
        out = np.empty((n_sectors, M * N * 12), dtype=np.uint8)
    elif fmt == "planar":
        one = [to_planar(b) for b in base]
        out = np.empty((n_sectors, 3, M, N), dtype=np.complex64)
    else:
        raise ValueError(f"unknown format {fmt!r}")
    for s in range(n_sectors):
        out[s] = one[s % distinct]
    return out
