"""Reader/writer for the reference's per-stage text dumps.

Format, derived from the reference's debug blocks (e.g. read.cc:258-270,
rpv2.cu:582-603) and from the shipped files: one matrix row per line; real stages
print ``value `` per element with iostream's default 6 significant digits; complex
stages print ``(re,im) ``; a trailing space precedes the newline; ``in/*.altb`` use
CRLF, ``out/*.out`` LF; ``99result`` prints ``zdb zdr`` per range gate.
"""
from __future__ import annotations

import io
import re

import numpy as np

STAGE_NAMES = ("00iq", "01hamm", "02fft1", "03fft2", "04abs", "05fft3", "06mult",
               "07conv", "08pow", "09zdb", "10zdr", "99result")
COMPLEX_STAGES = {"00iq", "01hamm", "02fft1", "03fft2", "05fft3", "06mult", "07conv"}

_CPLX = re.compile(r"\(([^,()]+),([^,()]+)\)")


def _fmt(v: float) -> str:
    """iostream default formatting (%g with 6 significant digits; inf/nan as glibc)."""
    if np.isnan(v):
        return "-nan" if np.signbit(v) else "nan"
    if np.isinf(v):
        return "-inf" if v < 0 else "inf"
    return "%g" % v


def format_dump(a: np.ndarray, *, crlf: bool = False) -> str:
    a = np.atleast_2d(a)
    eol = "\r\n" if crlf else "\n"
    out = io.StringIO()
    if np.iscomplexobj(a):
        for row in a:
            out.write("".join(f"({_fmt(z.real)},{_fmt(z.imag)}) " for z in row))
            out.write(eol)
    else:
        for row in a:
            out.write("".join(_fmt(float(v)) + " " for v in row))
            out.write(eol)
    return out.getvalue()


def write_dump(path: str, a: np.ndarray, *, crlf: bool | None = None) -> None:
    if crlf is None:
        crlf = path.endswith(".altb")
    with open(path, "w", newline="") as f:
        f.write(format_dump(a, crlf=crlf))


def format_result(zdb: np.ndarray, zdr: np.ndarray) -> str:
    """``99result`` layout: one ``zdb zdr`` line per gate (no trailing space)."""
    return "".join(f"{_fmt(float(a))} {_fmt(float(b))}\n" for a, b in zip(zdb, zdr))


def read_dump(path: str) -> np.ndarray:
    """Parse a stage dump into a 2-D array (complex128 for ``(re,im)`` stages)."""
    with open(path, "r", newline="") as f:
        text = f.read()
    lines = [ln for ln in text.replace("\r\n", "\n").split("\n") if ln.strip()]
    if lines and "(" in lines[0]:
        rows = [[complex(float(a), float(b)) for a, b in _CPLX.findall(ln)] for ln in lines]
        return np.array(rows, dtype=np.complex128)
    rows = [[float(t) for t in ln.split()] for ln in lines]
    width = max(len(r) for r in rows)
    if any(len(r) != width for r in rows):
        raise ValueError(f"{path}: ragged dump")
    return np.array(rows, dtype=np.float64)
