"""ctypes front-end to oracle/liboracle.so — the CPU restatement of the reference chain.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product package
(``weather-radar-processing_b200``) never imports this module.

Parity status: PINNED — see oracle/wrp_oracle.h for what pins it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")

STAGES_COMPLEX_FULL = ("s00_iq", "s01_hamm", "s02_fft1", "s03_fft2")
STAGES_REAL_HALF = ("s04_abs", "s08_pow")
STAGES_COMPLEX_HALF = ("s05_fft3", "s06_mult", "s07_conv")
STAGE_FIELDS = (
    "s00_iq", "s01_hamm", "s02_fft1", "s03_fft2", "s04_abs",
    "s05_fft3", "s06_mult", "s07_conv", "s08_pow", "power",
)


class _Cfg(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("C", C.c_int), ("ma_taps", C.c_int),
                ("range_res", C.c_double), ("calib", C.c_double)]


class _Dumps(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in STAGE_FIELDS]


@dataclass
class ChainResult:
    zdb: np.ndarray
    zdr: np.ndarray
    stages: dict  # name -> ndarray (complex stages as complex dtype)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when the reference checkout exists)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("wrp_oracle.c", "wrp_oracle_impl.inc", "wrp_oracle.h")
    ):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def build_ref(ref_root: str = "/root/reference") -> bool:
    """Compile the unmodified reference CPU programs into oracle/_ref (only where
    the reference checkout exists, i.e. in the build container)."""
    if not os.path.exists(os.path.join(ref_root, "read.cc")):
        return False
    subprocess.run(["make", "-C", _HERE, "ref", f"REF={ref_root}"], check=True, capture_output=True)
    return True


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_int
        L.wrpo_hamming_f64.argtypes = [ip, ip, C.c_void_p, dp]
        L.wrpo_hamming_f32.argtypes = [ip, ip, C.c_void_p, fp]
        L.wrpo_ma_f64.argtypes = [ip, C.c_void_p]
        L.wrpo_ma_f32.argtypes = [ip, C.c_void_p]
        L.wrpo_ma_fft_f64.argtypes = [ip, ip, C.c_void_p]
        L.wrpo_ma_fft_f32.argtypes = [ip, ip, C.c_void_p]
        L.wrpo_decode_wire_f64.argtypes = [C.c_void_p, ip, ip, ip, C.c_void_p]
        L.wrpo_decode_wire_f32.argtypes = [C.c_void_p, ip, ip, ip, C.c_void_p]
        for nm in ("wrpo_chain_f64", "wrpo_chain_f32"):
            f = getattr(L, nm)
            f.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.POINTER(_Dumps), C.c_void_p, C.c_void_p]
            f.restype = C.c_int
        for sfx in ("f64", "f32"):
            getattr(L, "wrpo_pdop_" + sfx).argtypes = [ip, ip, ip] + [C.c_void_p] * 5
            getattr(L, "wrpo_products_" + sfx).argtypes = [C.POINTER(_Cfg), ip] + [C.c_void_p] * 4
        L.wrpo_batch_wire_f32.argtypes = [C.POINTER(_Cfg), C.c_void_p, ip, C.c_void_p, ip]
        L.wrpo_batch_wire_f32.restype = C.c_int
        L.wrpo_rel_l2_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.wrpo_rel_l2_f32.restype = C.c_double
        L.wrpo_fft_f64.argtypes = [C.c_void_p, ip, ip]
        L.wrpo_fft_f32.argtypes = [C.c_void_p, ip, ip]
        L.wrpo_ftob.argtypes = [C.c_float, C.c_void_p]
        L.wrpo_btof.argtypes = [C.c_void_p]
        L.wrpo_btof.restype = C.c_float
        _lib = L
    return _lib


def _cfg(M, N, Cn, ma_taps=7, range_res=30.0, calib=1941.05):
    return _Cfg(M, N, Cn, ma_taps, range_res, calib)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def hamming(M: int, N: int, dtype=np.float64):
    """(ham[M,N], c) — read.cc:9-38 (f64) / read_single.cc:17-47 (f32)."""
    ham = np.empty((M, N), dtype=dtype)
    if dtype == np.float64:
        c = C.c_double()
        lib().wrpo_hamming_f64(M, N, _ptr(ham), C.byref(c))
    else:
        c = C.c_float()
        lib().wrpo_hamming_f32(M, N, _ptr(ham), C.byref(c))
    return ham, c.value


def ma_taps(taps: int = 7, dtype=np.float64):
    g = np.empty(taps, dtype=dtype)
    (lib().wrpo_ma_f64 if dtype == np.float64 else lib().wrpo_ma_f32)(taps, _ptr(g))
    return g


def ma_fft(taps: int, N: int, dtype=np.float64):
    out = np.empty((N, 2), dtype=dtype)
    (lib().wrpo_ma_fft_f64 if dtype == np.float64 else lib().wrpo_ma_fft_f32)(taps, N, _ptr(out))
    return out[:, 0] + 1j * out[:, 1]


def fft(x: np.ndarray, sign: int = -1):
    """Bare oracle FFT on a 1-D complex array (returns a new array)."""
    if x.dtype == np.complex128:
        buf = np.ascontiguousarray(x).view(np.float64).copy()
        lib().wrpo_fft_f64(_ptr(buf), x.shape[0], sign)
        return buf.view(np.complex128)
    buf = np.ascontiguousarray(x.astype(np.complex64)).view(np.float32).copy()
    lib().wrpo_fft_f32(_ptr(buf), x.shape[0], sign)
    return buf.view(np.complex64)


def decode_wire(wire: np.ndarray, M: int, N: int, Cn: int = 3, dtype=np.float64):
    """Wire bytes (M*N*12 uint8) -> planar complex [C, M, N] (sector.cpp:52-62)."""
    wire = np.ascontiguousarray(wire, dtype=np.uint8)
    assert wire.size == M * N * 12
    out = np.empty((Cn, M, N, 2), dtype=dtype)
    (lib().wrpo_decode_wire_f64 if dtype == np.float64 else lib().wrpo_decode_wire_f32)(
        _ptr(wire), M, N, Cn, _ptr(out))
    return out.view(np.complex128 if dtype == np.float64 else np.complex64)[..., 0]


def chain(iq: np.ndarray, *, dumps: bool = False, precision: str = "f64", ma_taps_n: int = 7,
          range_res: float = 30.0, calib: float = 1941.05) -> ChainResult:
    """Run the restated chain on planar complex iq[C, M, N].

    precision "f64" follows read.cc:133-345, "f32" follows read_single.cc:222-502.
    """
    Cn, M, N = iq.shape
    real = np.float64 if precision == "f64" else np.float32
    cplx = np.complex128 if precision == "f64" else np.complex64
    x = np.ascontiguousarray(iq.astype(cplx))
    cfg = _cfg(M, N, Cn, ma_taps_n, range_res, calib)
    zdb = np.empty(M // 2, dtype=real)
    zdr = np.empty(M // 2, dtype=real)
    d = _Dumps()
    bufs = {}
    if dumps:
        for name in STAGES_COMPLEX_FULL:
            bufs[name] = np.empty((Cn, M, N), dtype=cplx)
        for name in STAGES_COMPLEX_HALF:
            bufs[name] = np.empty((Cn, M // 2, N), dtype=cplx)
        for name in STAGES_REAL_HALF:
            bufs[name] = np.empty((Cn, M // 2, N), dtype=real)
        bufs["power"] = np.empty((Cn, M // 2), dtype=real)
        for name, b in bufs.items():
            setattr(d, name, b.ctypes.data)
    fn = lib().wrpo_chain_f64 if precision == "f64" else lib().wrpo_chain_f32
    rc = fn(C.byref(cfg), _ptr(x), C.byref(d), _ptr(zdb), _ptr(zdr))
    if rc != 0:
        raise ValueError(f"oracle chain rejected sizes M={M} N={N} C={Cn} (rc={rc})")
    return ChainResult(zdb, zdr, bufs)


def pdop(s04: np.ndarray, taps: int = 7, precision: str = "f64"):
    """Stage 04 [rows, N] -> dict(s05, s06, s07 complex; s08 real) (read.cc:285-301)."""
    real = np.float64 if precision == "f64" else np.float32
    cplx = np.complex128 if precision == "f64" else np.complex64
    a = np.ascontiguousarray(s04, dtype=real)
    rows, N = a.shape
    out = {k: np.empty((rows, N), dtype=cplx) for k in ("s05_fft3", "s06_mult", "s07_conv")}
    out["s08_pow"] = np.empty((rows, N), dtype=real)
    fn = lib().wrpo_pdop_f64 if precision == "f64" else lib().wrpo_pdop_f32
    fn(rows, N, taps, _ptr(a), _ptr(out["s05_fft3"]), _ptr(out["s06_mult"]), _ptr(out["s07_conv"]),
       _ptr(out["s08_pow"]))
    return out


def products(pow_hh: np.ndarray, pow_vv: np.ndarray | None = None, precision: str = "f64",
             range_res: float = 30.0, calib: float = 1941.05):
    """Stage 08 [rows, N] -> (zdb, zdr) (read.cc:335-344)."""
    real = np.float64 if precision == "f64" else np.float32
    a = np.ascontiguousarray(pow_hh, dtype=real)
    b = None if pow_vv is None else np.ascontiguousarray(pow_vv, dtype=real)
    rows, N = a.shape
    cfg = _cfg(2 * rows, N, 2, 7, range_res, calib)
    zdb = np.empty(rows, dtype=real)
    zdr = np.empty(rows, dtype=real)
    fn = lib().wrpo_products_f64 if precision == "f64" else lib().wrpo_products_f32
    fn(C.byref(cfg), rows, _ptr(a), None if b is None else _ptr(b), _ptr(zdb), _ptr(zdr))
    return zdb, zdr


def batch_wire_f32(wire: np.ndarray, n_sectors: int, M: int, N: int, Cn: int = 3,
                   n_threads: int = 0):
    """OpenMP batch of the float chain on wire-format sectors -> (out[n, M/2, 2], threads)."""
    wire = np.ascontiguousarray(wire, dtype=np.uint8)
    assert wire.size == n_sectors * M * N * 12
    out = np.empty((n_sectors, M // 2, 2), dtype=np.float32)
    cfg = _cfg(M, N, Cn)
    used = lib().wrpo_batch_wire_f32(C.byref(cfg), _ptr(wire), n_sectors, _ptr(out), n_threads)
    return out, used


def rel_l2(ref: np.ndarray, got: np.ndarray) -> float:
    """error.cpp:15-32 metric."""
    a = np.ascontiguousarray(ref, dtype=np.float32)
    b = np.ascontiguousarray(got, dtype=np.float32)
    return lib().wrpo_rel_l2_f32(_ptr(a), _ptr(b), a.size)


def ftob(f: float) -> bytes:
    b = (C.c_uint8 * 4)()
    lib().wrpo_ftob(f, b)
    return bytes(b)


def btof(b: bytes) -> float:
    arr = (C.c_uint8 * 4)(*b)
    return lib().wrpo_btof(arr)
