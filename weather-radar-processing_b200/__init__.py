"""weather-radar-processing_b200 — B200-native per-sector weather-radar chain.

Python face of the product: a thin ctypes binding over the C ABI in ``include/wrp.h``
(``libwrp.so``, hand-written CUDA for sm_100a) plus Python mirrors of the reference's
host-side types (``Dimension3``/``Dimension4`` dimension.h:4-16, ``Sector`` sector.h:8-19,
``RadarProcessor`` radar_processor.h:14-96) so that tests read like the reference's own
call sequence.  There is no CPU fallback: without ``libwrp.so`` or without a CUDA device
every compute entry point raises.

This package never imports ``oracle`` (the CPU restatement is test infrastructure).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import dumpio, synth, volume  # noqa: F401  (re-exported helpers)

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("WRP_LIB") or os.path.join(_HERE, "libwrp.so")  # WRP_LIB: A/B builds in experiments

WRP_OK = 0
FMT_C64_PLANAR = 0
FMT_WIRE_I16BE = 1
MODE_FUSED = 0
MODE_STAGED = 1
DOPPLER_ENERGY = 0
DOPPLER_FFT = 1
CHAIN_AUTO = 0
CHAIN_QUEUE = 1
CHAIN_V1 = 2

STAGE_IDS = {
    "00iq": 0, "01hamm": 1, "02fft1": 2, "03fft2": 3, "04abs": 4, "05fft3": 5,
    "06mult": 6, "07conv": 7, "08pow": 8, "09zdb": 9, "10zdr": 10, "power": 11,
}
_COMPLEX_STAGES = {0, 1, 2, 3, 5, 6, 7}
_FULL_STAGES = {0, 1, 2, 3}

EXPORTED_SYMBOLS = (
    "wrp_version", "wrp_default_config", "wrp_create", "wrp_destroy", "wrp_last_error",
    "wrp_get_info", "wrp_get_constants", "wrp_process_device", "wrp_process_host",
    "wrp_submit", "wrp_collect", "wrp_alloc_pinned", "wrp_free_pinned", "wrp_dump_stage",
    "wrp_launch_count", "wrp_profile_enable", "wrp_profile_read", "wrp_pack_products",
    "wrp_chain_kernel_name", "wrp_set_stage02_tap", "wrp_set_product_mirrors", "wrp_process_host_to_device",
    "wrp_volume_create", "wrp_volume_shard", "wrp_volume_process", "wrp_volume_last_error", "wrp_volume_destroy",
)


class WrpError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libwrp status {status}: {message}")
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("n_rows_M", C.c_int), ("n_cols_N", C.c_int), ("n_channels", C.c_int),
        ("n_streams", C.c_int), ("ma_taps", C.c_int), ("range_res_m", C.c_float),
        ("calib", C.c_float), ("input_fmt", C.c_int), ("mode", C.c_int), ("max_batch", C.c_int),
        ("doppler_form", C.c_int), ("chain_impl", C.c_int), ("x2_lag", C.c_int), ("x2_ring", C.c_int),
        ("evict_first", C.c_int), ("debug", C.c_int),
    ]


class Info(C.Structure):
    _fields_ = [
        ("version", C.c_int), ("device", C.c_int), ("sm_count", C.c_int), ("l2_bytes", C.c_int),
        ("input_bytes_per_sector", C.c_size_t), ("output_floats_per_sector", C.c_size_t),
        ("intermediate_bytes_per_sector", C.c_size_t), ("chunk_sectors", C.c_int),
        ("kernels_per_chunk", C.c_int),
    ]


class Profile(C.Structure):
    _fields_ = [
        ("ms_decode", C.c_double), ("ms_range", C.c_double), ("ms_doppler", C.c_double),
        ("ms_staged", C.c_double), ("ms_chain", C.c_double), ("n_decode", C.c_ulonglong),
        ("n_range", C.c_ulonglong), ("n_doppler", C.c_ulonglong), ("n_staged", C.c_ulonglong),
        ("n_chain", C.c_ulonglong), ("sectors", C.c_ulonglong),
    ]


def build(verbose: bool = False) -> str:
    """Compile libwrp.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", REPO_ROOT, os.path.relpath(LIB_PATH, REPO_ROOT)],
                       capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libwrp.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load libwrp.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not built: run __graft_entry__.build() or `make`; there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, ip = C.c_void_p, C.c_int
        L.wrp_version.restype = ip
        L.wrp_default_config.argtypes = [C.POINTER(Config)]
        L.wrp_default_config.restype = None
        L.wrp_create.argtypes = [C.POINTER(Config), ip, C.POINTER(vp)]
        L.wrp_destroy.argtypes = [vp]
        L.wrp_destroy.restype = None
        L.wrp_last_error.argtypes = [vp]
        L.wrp_last_error.restype = C.c_char_p
        L.wrp_get_info.argtypes = [vp, C.POINTER(Info)]
        L.wrp_get_constants.argtypes = [vp, vp, vp, vp]
        L.wrp_process_device.argtypes = [vp, vp, ip, vp, vp]
        L.wrp_process_host.argtypes = [vp, vp, ip, vp]
        L.wrp_submit.argtypes = [vp, vp, ip, vp, vp]
        L.wrp_collect.argtypes = [vp, vp, vp, vp, ip, C.POINTER(ip)]
        L.wrp_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.wrp_free_pinned.argtypes = [vp]
        L.wrp_dump_stage.argtypes = [vp, ip, ip, ip, vp, C.POINTER(C.c_size_t)]
        L.wrp_launch_count.argtypes = [vp]
        L.wrp_launch_count.restype = C.c_ulonglong
        L.wrp_chain_kernel_name.argtypes = [vp]
        L.wrp_chain_kernel_name.restype = C.c_char_p
        L.wrp_set_stage02_tap.argtypes = [vp, vp]
        L.wrp_set_product_mirrors.argtypes = [vp, C.POINTER(vp), ip]
        L.wrp_process_host_to_device.argtypes = [vp, vp, ip, vp]
        L.wrp_volume_create.argtypes = [C.POINTER(Config), C.POINTER(ip), ip, ip, ip, C.POINTER(vp)]
        L.wrp_volume_shard.argtypes = [vp, ip, C.POINTER(ip), C.POINTER(ip)]
        L.wrp_volume_process.argtypes = [vp, vp, vp]
        L.wrp_volume_last_error.argtypes = [vp]
        L.wrp_volume_last_error.restype = C.c_char_p
        L.wrp_volume_destroy.argtypes = [vp]
        L.wrp_volume_destroy.restype = None
        L.wrp_profile_enable.argtypes = [vp, ip]
        L.wrp_profile_read.argtypes = [vp, C.POINTER(Profile), ip]
        L.wrp_pack_products.argtypes = [vp, ip, ip, ip, ip, vp, vp]
        _lib = L
    return _lib


def default_config(**overrides) -> Config:
    cfg = Config()
    lib().wrp_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"wrp_config has no field {k!r}")
        setattr(cfg, k, v)
    return cfg


def bind_host_to_gpu(device: int) -> list[int]:
    """Pin the calling process to the CPU cores NVML reports as local to ``device`` (same NUMA node /
    PCIe root), so that pinned host buffers allocated afterwards are node-local and H2D copies do not
    cross the socket interconnect.  Matters when several ranks stream from host memory at once (the
    reference is single-GPU and never had to care).  Returns the cores chosen ([] = left unchanged)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = [c for c in cores if c in allowed]
        if cores and len(cores) < len(allowed):
            os.sched_setaffinity(0, cores)
            return cores
    except Exception:
        pass
    return []


class PinnedBuffer:
    """Page-locked host memory (the reference's cudaMallocHost'ed p_iq, rpv2.cu:291)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        rc = lib().wrp_alloc_pinned(nbytes, C.byref(p))
        if rc != WRP_OK:
            raise WrpError(rc, lib().wrp_last_error(None).decode())
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            lib().wrp_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RadarChain:
    """One libwrp handle (one device).  Mirrors the rpv2 call sequence
    read_matrix -> copy_matrix_to_device -> perform_stage_1/2/3 -> copy_result_to_host
    (rpv2.cu:665-683) behind process_host / submit / collect."""

    def __init__(self, device: int = 0, **cfg_overrides):
        self.cfg = default_config(**cfg_overrides)
        h = C.c_void_p()
        rc = lib().wrp_create(C.byref(self.cfg), device, C.byref(h))
        if rc != WRP_OK:
            raise WrpError(rc, lib().wrp_last_error(None).decode())
        self._h = h
        self.M, self.N, self.C = self.cfg.n_rows_M, self.cfg.n_cols_N, self.cfg.n_channels

    # -- plumbing ---------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != WRP_OK:
            raise WrpError(rc, lib().wrp_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().wrp_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self) -> Info:
        i = Info()
        self._check(lib().wrp_get_info(self._h, C.byref(i)))
        return i

    @property
    def input_bytes_per_sector(self) -> int:
        return int(self.info.input_bytes_per_sector)

    @property
    def launch_count(self) -> int:
        return int(lib().wrp_launch_count(self._h))

    @property
    def chain_kernel(self) -> str:
        """Name of the kernel that carries the chain for this configuration."""
        return lib().wrp_chain_kernel_name(self._h).decode()

    def set_stage02_tap(self, dev_ptr: int | None):
        """Streaming kernel only: also store the range-FFT rows k < M/2 it folds to device memory
        [sector][channel][M/2][N] complex64 at dev_ptr (None switches the tap off)."""
        self._check(lib().wrp_set_stage02_tap(self._h, dev_ptr))

    def set_product_mirrors(self, dev_ptrs=()):
        """The fused gather (wrp_set_product_mirrors): process_device / process_host_to_device also store
        every product at the same float index of each mirror — device pointers, normally peer-mapped
        slices of the product volume on the other devices of the box.  () switches it off."""
        ptrs = [int(p) for p in dev_ptrs]
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        self._check(lib().wrp_set_product_mirrors(self._h, arr, len(ptrs)))

    def constants(self):
        """(hamming[M,N], taps[ma_taps], fft_ma[N] complex) — rpv2.cu:222-281."""
        ham = np.empty((self.M, self.N), np.float32)
        taps = np.empty(self.cfg.ma_taps, np.float32)
        fm = np.empty((self.N, 2), np.float32)
        self._check(lib().wrp_get_constants(self._h, ham.ctypes.data, taps.ctypes.data, fm.ctypes.data))
        return ham, taps, fm[:, 0] + 1j * fm[:, 1]

    # -- compute ----------------------------------------------------------------------
    def process_device(self, dev_iq_ptr: int, n_sectors: int, dev_out_ptr: int, stream: int = 0):
        """HBM-resident batch on raw device pointers (e.g. torch tensors' data_ptr())."""
        self._check(lib().wrp_process_device(self._h, dev_iq_ptr, n_sectors, dev_out_ptr, stream))

    def process_host(self, host_iq, n_sectors: int, out: np.ndarray | None = None) -> np.ndarray:
        """Host-buffer batch -> float32[n_sectors, M/2, 2] (ZdB, ZDR)."""
        if isinstance(host_iq, PinnedBuffer):
            ptr, nbytes = host_iq.ptr, host_iq.nbytes
        else:
            host_iq = np.ascontiguousarray(host_iq)
            ptr, nbytes = host_iq.ctypes.data, host_iq.nbytes
        need = n_sectors * self.input_bytes_per_sector
        if n_sectors < 0:
            self._check(lib().wrp_process_host(self._h, ptr, n_sectors, None))
        if nbytes < need:
            raise ValueError(f"input holds {nbytes} bytes, {n_sectors} sectors need {need}")
        if out is None:
            out = np.empty((n_sectors, self.M // 2, 2), np.float32)
        self._check(lib().wrp_process_host(self._h, ptr, n_sectors, out.ctypes.data))
        return out

    def process_host_to_device(self, host_iq, n_sectors: int, dev_out_ptr: int):
        """Host-buffer batch whose products stay on the device: dev_out_ptr -> float32[n_sectors, M/2, 2]."""
        if isinstance(host_iq, PinnedBuffer):
            ptr, nbytes = host_iq.ptr, host_iq.nbytes
        else:
            host_iq = np.ascontiguousarray(host_iq)
            ptr, nbytes = host_iq.ctypes.data, host_iq.nbytes
        if nbytes < n_sectors * self.input_bytes_per_sector:
            raise ValueError("input buffer too small")
        self._check(lib().wrp_process_host_to_device(self._h, ptr, n_sectors, dev_out_ptr))

    def submit(self, host_iq, n_sectors: int, sector_ids=None, elev_ids=None):
        if isinstance(host_iq, PinnedBuffer):
            ptr = host_iq.ptr
        else:
            host_iq = np.ascontiguousarray(host_iq)
            ptr = host_iq.ctypes.data
        sid = None if sector_ids is None else np.ascontiguousarray(sector_ids, np.int32)
        eid = None if elev_ids is None else np.ascontiguousarray(elev_ids, np.int32)
        self._check(lib().wrp_submit(self._h, ptr, n_sectors,
                                     None if sid is None else sid.ctypes.data,
                                     None if eid is None else eid.ctypes.data))

    def collect(self):
        """-> (products[n, M/2, 2], sector_ids[n], elev_ids[n]); n == 0 when idle."""
        cap = self.cfg.max_batch
        out = np.empty((cap, self.M // 2, 2), np.float32)
        sid = np.empty(cap, np.int32)
        eid = np.empty(cap, np.int32)
        n = C.c_int(0)
        self._check(lib().wrp_collect(self._h, out.ctypes.data, sid.ctypes.data, eid.ctypes.data, cap,
                                      C.byref(n)))
        return out[: n.value], sid[: n.value], eid[: n.value]

    def dump_stage(self, stage, sector_in_batch: int = 0, channel: int = 0) -> np.ndarray:
        """Stage dump of the last WRP_MODE_STAGED batch as a [rows, N] (or [M/2]) array."""
        sid = STAGE_IDS[stage] if isinstance(stage, str) else int(stage)
        nbytes = C.c_size_t(0)
        self._check(lib().wrp_dump_stage(self._h, sector_in_batch, sid, channel, None, C.byref(nbytes)))
        buf = np.empty(nbytes.value // 4, np.float32)
        self._check(lib().wrp_dump_stage(self._h, sector_in_batch, sid, channel, buf.ctypes.data,
                                         C.byref(nbytes)))
        if sid in (9, 10, 11):
            return buf
        rows = self.M if sid in _FULL_STAGES else self.M // 2
        if sid in _COMPLEX_STAGES:
            return buf.view(np.complex64).reshape(rows, self.N)
        return buf.reshape(rows, self.N)

    def profile_enable(self, on: bool = True):
        self._check(lib().wrp_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> Profile:
        p = Profile()
        self._check(lib().wrp_profile_read(self._h, C.byref(p), 1 if reset else 0))
        return p


class VolumeScan:
    """wrp_volume_*: one volume scan (n_elevations x n_sectors units, rpv2.cu:572-579 order) sharded in
    contiguous unit blocks over the listed devices of this box, one host thread per shard, the product
    volume gathered on devices[0] by peer copies.  No torch, no NCCL: the C ABI alone."""

    def __init__(self, devices, n_sectors: int, n_elevations: int, **cfg_overrides):
        self.cfg = default_config(**cfg_overrides)
        self.devices = list(devices)
        self.n_sectors, self.n_elevations = n_sectors, n_elevations
        arr = (C.c_int * len(self.devices))(*self.devices)
        v = C.c_void_p()
        rc = lib().wrp_volume_create(C.byref(self.cfg), arr, len(self.devices), n_sectors, n_elevations, C.byref(v))
        if rc != WRP_OK:
            raise WrpError(rc, lib().wrp_volume_last_error(None).decode())
        self._v = v

    def shard(self, g: int) -> tuple[int, int]:
        lo, n = C.c_int(0), C.c_int(0)
        rc = lib().wrp_volume_shard(self._v, g, C.byref(lo), C.byref(n))
        if rc != WRP_OK:
            raise WrpError(rc, "wrp_volume_shard: bad shard index")
        return lo.value, n.value

    def process(self, host_iq, out: np.ndarray | None = None) -> np.ndarray:
        """-> float32[n_elevations * n_sectors, M/2, 2] in the reference's sitdim order."""
        units = self.n_sectors * self.n_elevations
        ptr = host_iq.ptr if isinstance(host_iq, PinnedBuffer) else np.ascontiguousarray(host_iq).ctypes.data
        if out is None:
            out = np.empty((units, self.cfg.n_rows_M // 2, 2), np.float32)
        rc = lib().wrp_volume_process(self._v, ptr, out.ctypes.data)
        if rc != WRP_OK:
            raise WrpError(rc, lib().wrp_volume_last_error(self._v).decode())
        return out

    def close(self):
        if getattr(self, "_v", None):
            lib().wrp_volume_destroy(self._v)
            self._v = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_products(zdb_zdr: np.ndarray, sector: int, elev: int = 0, with_elev: bool = True):
    """send_results' packets (rpv2.cu:620-663): -> (zdb_packet, zdr_packet) bytes."""
    a = np.ascontiguousarray(zdb_zdr, np.float32)
    gates = a.shape[0]
    size = (4 if with_elev else 2) + 4 * gates
    zb = np.empty(size, np.uint8)
    zr = np.empty(size, np.uint8)
    n = lib().wrp_pack_products(a.ctypes.data, gates, sector, elev, 1 if with_elev else 0,
                                zb.ctypes.data, zr.ctypes.data)
    if n != size:
        raise WrpError(-n, "wrp_pack_products failed")
    return zb.tobytes(), zr.tobytes()


# ---- Python mirrors of the reference's host types ---------------------------------------
class Dimension3:
    """dimension.h:4-9 / dimension.cpp:3-11."""

    def __init__(self, w: int, h: int, d: int):
        self.width, self.height, self.depth = w, h, d
        self.m_size = w * h
        self.total_size = w * h * d

    def at_depth(self, x: int, y: int, depth: int) -> int:
        return y * self.width + x + depth * self.width * self.height


class Dimension4:
    """dimension.h:11-16 / dimension.cpp:13-21."""

    def __init__(self, w: int, h: int, c: int, d: int):
        self.width, self.height, self.copies, self.depth = w, h, c, d
        self.m_size = w * h
        self.total_size = w * h * c * d

    def copy_at_depth(self, x: int, y: int, copy: int, depth: int) -> int:
        return (y * self.width + x + copy * self.width * self.height
                + depth * self.width * self.height * self.copies)


class Sector:
    """sector.h:8-19: three int16 arrays of 2*sweeps*samples (I,Q interleaved)."""

    def __init__(self, num_sweeps: int, num_samples: int):
        self.sweeps, self.samples = num_sweeps, num_samples
        n = 2 * num_sweeps * num_samples
        self.hh = np.zeros(n, np.int16)
        self.vv = np.zeros(n, np.int16)
        self.vh = np.zeros(n, np.int16)

    def fromByteArray(self, buff) -> None:  # noqa: N802 (reference spelling, sector.cpp:52-62)
        raw = np.frombuffer(bytes(buff), dtype=">i2", count=6 * self.sweeps * self.samples)
        rec = raw.reshape(-1, 6).astype(np.int16)
        self.hh[:] = rec[:, 0:2].reshape(-1)
        self.vv[:] = rec[:, 2:4].reshape(-1)
        self.vh[:] = rec[:, 4:6].reshape(-1)
