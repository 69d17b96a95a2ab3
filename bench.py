#!/usr/bin/env python
"""bench.py — sectors/s of the per-sector weather-radar chain on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic sectors:
default sector shape 1024 x 512 x 3 (rpv2.cu:38-45), ``--sectors`` sectors per GPU per step.

ours:
  value      whole-job sectors/s with the batch already resident in HBM as planar complex
             float — the reference's own device format (rpv2.cu:379-381) — through
             wrp_process_device (CUDA events on the launching stream, max over ranks).  A burst
             figure (K steps of 0.4 ms); `sustained` repeats the same step back to back for
             >= 2.5 s with the clocks and the power sampled meanwhile.
  e2e        the same metric through the public host-buffer call wrp_process_host with
             pinned HOST buffers in the radar's wire format (what the reference's
             read_matrix receives, sector.cpp:52-62): H2D of every step's input and D2H of
             its products are inside the timed region.
  volume     BASELINE config 4 at this N: one volume scan (9 x 143 wire sectors in pinned host
             memory) sharded contiguously over the ranks, products gathered — strong scaling.
  stress     BASELINE config 5 at this N: 4096 x 1024 x 3 sectors, 64 resident per GPU, against
             the HBM roofline — weak scaling.
  roofline   dominant kernel (chain_stream_kernel) against the measured HBM copy bandwidth in
             MEASURED_PEAKS.json: algorithmic bytes per launch / mean launch time (CUDA events
             recorded around every launch inside the timed region).
  cpu_baseline  the oracle's float chain (CPU port of read_single.cc) on all host cores,
             bounded sample.  The oracle is only the baseline/checker here, never the product.
reference:
  the reference's own CPU program (oracle/_ref/read_single_ref = unmodified read_single.cc
  built against the FFTW/UDP shims; falls back to the oracle port when _ref was not built)
  on all host cores; rank 0 only.

Multi-GPU (torchrun, one rank per GPU): sectors are independent, each rank
processes its own shard — weak scaling — and the product volume (9 elevations) is gathered by the
chain kernel itself: its epilogue stores every product into all ranks' volume buffers (CUDA-IPC
mapped peers over NVLink, wrp_set_product_mirrors); per volume one 4-byte all-reduce on a side
stream is the completion handshake, and the timed region ends after the last one.  `--gather nccl`
(or a box where the peers cannot be mapped) all-gathers the volume with NCCL once per volume on the
side stream instead.  There is no other collective.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, N, C = 1024, 512, 3
ALGO_BYTES_C64 = C * M * N * 8 + (M // 2) * 8      # 12 587 008
ALGO_BYTES_WIRE = M * N * 12 + (M // 2) * 8        # 6 295 552
WORKLOAD = ("default sector 1024x512x3 (rpv2.cu:38-45), all stages fused; one step = one 360-degree PPI "
            "elevation of 143 sectors (rpv2.cu:39 n_sectors) per GPU")
# the SAME object in both arms' `config` (the driver compares them); run details go under `run`
CONFIG = {"workload": WORKLOAD, "M": M, "N": N, "channels": C, "sectors_per_elevation": 143}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


def cpu_baseline_port(sample_sectors: int, target_seconds: float = 12.0):
    """Oracle float chain (port of read_single.cc) on all host cores, wire-format input.
    A buffer of `sample_sectors` synthetic sectors is processed repeatedly for ~target_seconds."""
    import oracle
    synth = importlib.import_module("weather-radar-processing_b200.synth")
    cores = os.cpu_count() or 1
    wire = synth.make_batch(M, N, sample_sectors, fmt="wire", distinct=2)
    oracle.batch_wire_f32(wire[:1], 1, M, N, C, 1)  # warm the page cache / libm
    done, used = 0, cores
    t0 = time.perf_counter()
    while True:
        _, used = oracle.batch_wire_f32(wire, sample_sectors, M, N, C, cores)
        done += sample_sectors
        dt = time.perf_counter() - t0
        if dt >= target_seconds or done >= 100000:
            break
    return {"value": done / dt, "unit": "sectors/s", "cores": used, "kind": "port",
            "sample": f"{done} synthetic wire-format sectors 1024x512x3 ({sample_sectors}-sector buffer repeated), "
                      f"oracle float chain = CPU port of read_single.cc, OpenMP over sectors, {used} threads, {dt:.1f} s"}


def reference_gpu_leg(streams: int = 3):
    """The reference's OWN GPU path on this box: oracle/_ref/gpu_1fp_unistream_ref = unmodified
    gpu_1fp_unistream.cu (cuFFT cascade + 11 element-wise kernels, pinned buffers, N streams; built with
    nvcc for sm_100a against cuFFT; FFTW only for the 7-tap init transform -> shim).  Its stock main()
    processes 127 synthetic 1024x512x3 sectors (host fill + H2D of planar c64 + ~40 launches per sector)
    and prints its own cudaEvent time.  A side figure next to the CPU reference arm."""
    exe = os.path.join(ROOT, "oracle", "_ref", "gpu_1fp_unistream_ref")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/gpu_1fp_unistream_ref not built (needs nvcc + the reference checkout)"}
    best = None
    try:
        for _ in range(3):  # the first run pays cuFFT/plan and context start-up outside its timed loop anyway
            r = subprocess.run([exe, str(streams)], capture_output=True, text=True, timeout=180)
            for line in r.stdout.splitlines():
                if "transfer and execute (ms)" in line:
                    ms = float(line.split(":")[1])
                    best = ms if best is None else min(best, ms)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"run failed: {e}"}
    if not best:
        return {"unavailable": "no timing line in the program's output"}
    return {"value": 127 / (best * 1e-3), "unit": "sectors/s", "ms_for_127_sectors": best, "streams": streams,
            "program": "gpu_1fp_unistream.cu (unmodified, nvcc sm_100a + cuFFT), best of 3 runs",
            "note": "the reference's own timed loop: host-side synthetic fill, async H2D of 12.6 MB planar c64 per "
                    "sector from pinned memory, cuFFT + element-wise cascade, 4 KiB D2H"}


def run_reference(args):
    """--impl reference: the reference's CPU program on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = importlib.import_module("weather-radar-processing_b200.synth")
    cores = os.cpu_count() or 1
    exe = os.path.join(ROOT, "oracle", "_ref", "read_single_ref")
    per_proc = 8  # amortises the reference's start-up (window tables, plans: ~35 ms) to ~4 % of a process's run
    steps, warmup = args.steps, args.warmup
    if os.path.exists(exe):
        kind = "reference"
        tmp = tempfile.mkdtemp(prefix="wrp_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        inp = os.path.join(tmp, "wire.bin")
        synth.make_batch(M, N, per_proc, fmt="wire").tofile(inp)

        def one_step():
            procs = []
            for p in range(cores):
                env = dict(os.environ, WRP_FAKE_UDP_IN=inp, WRP_FAKE_UDP_OUT=os.path.join(tmp, f"o{p}"))
                procs.append(subprocess.Popen([exe], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env))
            for p in procs:
                if p.wait() != 0:
                    raise RuntimeError("read_single_ref failed")
            return cores * per_proc
        desc = (f"{cores} concurrent processes of oracle/_ref/read_single_ref (unmodified read_single.cc, "
                f"float, hh+vv+vh, FFTW replaced by oracle/shim), {per_proc} wire-format sector each per step")
    else:
        kind = "port"
        import oracle
        n = max(cores, 8)
        wire = synth.make_batch(M, N, n, fmt="wire", distinct=2)

        def one_step():
            oracle.batch_wire_f32(wire, n, M, N, C, cores)
            return n
        desc = f"oracle float chain (port of read_single.cc), OpenMP {cores} threads, {n} sectors per step"
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        done += one_step()
    dt = time.perf_counter() - t0
    v = done / dt
    line = {
        "impl": "reference", "metric": "sectors_per_s", "value": v, "unit": "sectors/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "run": {"sectors_per_step": done // steps, "input_fmt": "wire_i16be",
                "note": "the reference's own CPU chain (read_single.cc) on the host cores; FFTW is absent from the image, "
                        "the stand-in is oracle/shim/fftw3.h — a CPU baseline, NOT the reference's cuFFT cascade"},
        "cpu_baseline": {"value": v, "unit": "sectors/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": "sectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_gpu": reference_gpu_leg() if args.gpus == 1 and not args.skip_reference_gpu else None,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class _DeviceArray:
    """A raw device allocation seen by torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def peer_mapped_buffers(torch, dist, dev, world, rank, shape, count):
    """`count` zero-filled float32 buffers of `shape` on this rank's device, each mapped into every other rank's
    address space (CUDA IPC; the ranks are processes of one box, the mapping goes over NVLink).  Returns
    (tensors, ptrs): tensors[i] = this rank's buffer i, ptrs[i][r] = rank r's buffer i as a device pointer valid
    in THIS process (ptrs[i][rank] is the local one) — the mirrors of wrp_set_product_mirrors.  Returns None on
    every rank if any rank could not allocate or map (the ranks agree through an all-reduce after each phase, so
    nobody is left waiting in a collective)."""

    def ck(ret):
        err, *rest = ret if isinstance(ret, tuple) else (ret,)
        if int(err) != 0:
            raise RuntimeError(f"CUDA runtime error {err}")
        return rest[0] if rest else None

    def all_ok(ok):
        t = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    nbytes = 4 * int(np.prod(shape))
    local, mine = [], []
    try:
        from cuda.bindings import runtime as cudart

        ck(cudart.cudaSetDevice(dev.index))
        for _ in range(count):
            ptr = int(ck(cudart.cudaMalloc(nbytes)))
            ck(cudart.cudaMemset(ptr, 0, nbytes))
            local.append(ptr)
            mine.append(bytes(ck(cudart.cudaIpcGetMemHandle(ptr)).reserved))
        ok = True
    except Exception as ex:  # noqa: BLE001 - reported, then the NCCL gather takes over
        print(f"bench.py: rank {rank}: cannot export the volume buffers ({ex})", file=sys.stderr)
        ok = False
    if not all_ok(ok):
        return None
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    tensors, ptrs = [], []
    try:
        for i in range(count):
            row = []
            for r in range(world):
                if r == rank:
                    row.append(local[i])
                    continue
                h = cudart.cudaIpcMemHandle_t()
                h.reserved = handles[r][i]
                row.append(int(ck(cudart.cudaIpcOpenMemHandle(h, cudart.cudaIpcMemLazyEnablePeerAccess))))
            tensors.append(torch.as_tensor(_DeviceArray(local[i], shape), device=dev))
            ptrs.append(row)
        torch.cuda.synchronize()
        ok = True
    except Exception as ex:  # noqa: BLE001
        print(f"bench.py: rank {rank}: cannot map the peers' volume buffers ({ex})", file=sys.stderr)
        ok = False
    if not all_ok(ok):
        return None
    return tensors, ptrs


def _dist_env():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def _max_over_ranks(x, dev, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _timed_launches(fn, n, stream):
    """n back-to-back calls of fn() on `stream`, device time in ms (CUDA events)."""
    import torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(n):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def leg_volume(wrp, args, dev, world, rank, local_rank, steps):
    """BASELINE config 4 at this N: one volume scan = 9 elevations x 143 sectors of wire records in pinned
    host memory, contiguous (elevation, sector) shards (rpv2.cu:572-579 order), products gathered on the
    device (one all-gather) — strong scaling.  Returns the JSON object (rank 0) or None."""
    import torch
    import torch.distributed as dist
    synth = wrp.synth
    S, E = 143, 9
    U = S * E
    lo, hi = wrp.volume.shard_bounds(U, rank, world)
    base = [synth.to_wire(synth.make_sector_int16(M, N, s, 0)) for s in range(4)]
    pin = wrp.PinnedBuffer((hi - lo) * M * N * 12)
    view = pin.array.reshape(hi - lo, M * N * 12)
    for k in range(lo, hi):
        view[k - lo] = base[k % 4].reshape(-1)
    chain = wrp.RadarChain(local_rank, input_fmt=wrp.FMT_WIRE_I16BE, max_batch=args.host_piece, n_streams=args.streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    vol = wrp.volume.process_volume(chain, pin, U, dev)  # warm-up: pinned ring, NCCL channels
    barrier()
    l0 = chain.launch_count
    t0 = time.perf_counter()
    for _ in range(steps):
        vol = wrp.volume.process_volume(chain, pin, U, dev)
    barrier()
    dt = _max_over_ranks(time.perf_counter() - t0, dev, world)
    v = vol.cpu().numpy()
    # the gathered volume is checked: shape, finite, and unit k equals unit k mod 4 (same synthetic sector)
    ok = v.shape == (U, M // 2, 2) and np.isfinite(v[:, 1:]).all() and all(
        np.allclose(v[k, 1:], v[k % 4, 1:], rtol=0, atol=1e-3) for k in range(0, U, 61))
    launches = chain.launch_count - l0
    chain.close()
    pin.close()
    if not ok:
        raise SystemExit("bench.py: gathered volume is wrong")
    value = U * steps / dt
    return {"value": value, "unit": "sectors/s", "scaling": "strong", "ms_per_volume": dt / steps * 1e3,
            "units": U, "steps": steps, "volume_bytes": int(v.nbytes), "gathered_volume_checked": True,
            "h2d_gbs_per_gpu": value / world * M * N * 12 / 1e9, "gpu_launches": int(launches),
            "note": "9 elevations x 143 wire sectors from pinned host memory, contiguous (elevation, sector) shards, "
                    "products stay on the device and are all-gathered once per volume"}


def leg_stress(wrp, args, dev, world, local_rank, stream, peak):
    """BASELINE config 5 at this N: 4096 x 1024 x 3 planar sectors, `--stress-sectors` resident per GPU."""
    import torch
    synth = wrp.synth
    SM_, SN_, SS_ = 4096, 1024, args.stress_sectors
    sx = synth.to_planar(synth.make_sector_int16(SM_, SN_, 0, 0), C)
    d_sx = torch.from_numpy(np.ascontiguousarray(sx).view(np.float32).reshape(-1)).to(dev).repeat(SS_)
    d_so = torch.empty((SS_, SM_ // 2, 2), dtype=torch.float32, device=dev)
    reps = 5
    with wrp.RadarChain(local_rank, n_rows_M=SM_, n_cols_N=SN_, n_channels=C, max_batch=1) as sch:
        kernel = sch.chain_kernel
        run = lambda: sch.process_device(d_sx.data_ptr(), SS_, d_so.data_ptr(), stream.cuda_stream)
        for _ in range(2):
            run()
        ms = _max_over_ranks(_timed_launches(run, reps, stream), dev, world)
    if not torch.isfinite(d_so[:, 1:]).all():
        raise SystemExit("bench.py: non-finite stress products")
    sv = world * SS_ * reps / (ms * 1e-3)
    sbytes = C * SM_ * SN_ * 8 + (SM_ // 2) * 2 * 4
    del d_sx, d_so
    return {"value": sv, "unit": "sectors/s", "scaling": "weak", "shape": f"{SM_}x{SN_}x{C} c64",
            "sectors_resident_per_gpu": SS_, "kernel": kernel, "algorithmic_bytes_per_sector": sbytes,
            "gbs_per_gpu": sv / world * sbytes / 1e9, "hbm_frac": sv / world * sbytes / 1e9 / peak}


def run_ours(args):
    import torch
    import torch.distributed as dist

    wrp = importlib.import_module("weather-radar-processing_b200")
    synth = wrp.synth
    world, rank, local_rank = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libwrp has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = wrp.bind_host_to_gpu(local_rank) if world > 1 and not os.environ.get("WRP_NO_BIND") else []
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = args.sectors
    steps, warmup = args.steps, args.warmup
    peak, peak_src = load_peaks()

    # ---- inputs: a few distinct synthetic sectors tiled to the batch; larger than L2 --------
    planar = synth.make_batch(M, N, S, fmt="planar", first_sector=rank * 7, distinct=4)
    d_in = torch.from_numpy(planar.view(np.float32).reshape(-1)).to(dev)
    # products land elevation by elevation in a volume buffer [E][S][gates][2] (the reference's result[] order,
    # rpv2.cu:607, 736); two volumes so that one can be gathered while the next one fills
    E = 9
    vol = [torch.empty((E, S, M // 2, 2), dtype=torch.float32, device=dev) for _ in range(2)]
    d_out = [vol[0][0], vol[1][0]]  # scratch views for the side legs below
    fused = world > 1 and args.gather == "fused"
    if fused:
        # the fused gather: every rank's kernels store their products into all ranks' volume buffers themselves
        mapped = peer_mapped_buffers(torch, dist, dev, world, rank, (world, E, S, M // 2, 2), 2)
        if mapped is None:
            fused = False  # said on stderr; the line's run.gather then reads "nccl"
        else:
            gathered, peer_ptrs = mapped
            flag = torch.zeros(1, device=dev)
    if not fused:
        gathered = [torch.zeros((world, E, S, M // 2, 2), dtype=torch.float32, device=dev) for _ in range(2)] if world > 1 else None
    chain = wrp.RadarChain(local_rank, max_batch=args.host_piece)
    info = chain.info
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    gather_done = [None, None]
    step_no = [0]
    n_gathers = [0]
    slot_bytes = S * (M // 2) * 2 * 4  # one elevation of one rank

    def gather(v):
        """End of a volume.  --gather nccl: the finished product volume is all-gathered on the side stream.
        --gather fused (default): the kernels have already stored it into every rank's buffer through the product
        mirrors; what is left is the completion handshake a consumer needs — one 4-byte all-reduce on the side
        stream, behind this rank's last kernel of the volume.  (Also measured: a DMA gather, cudaMemcpyAsync into
        the IPC-mapped peers — equal at N = 2, 31 % slower at N = 8; profiles/r02_ab_variants.md.)"""
        ready = torch.cuda.Event()
        ready.record(stream)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            if fused:
                dist.all_reduce(flag)
            else:
                dist.all_gather_into_tensor(gathered[v], vol[v])
            gather_done[v] = torch.cuda.Event()
            gather_done[v].record(side)
        n_gathers[0] += 1

    def step():
        """One PPI elevation per GPU.  After the ninth elevation the product volume is complete on every rank
        (fused: stored by the kernels; nccl: all-gathered on the side stream) while the next volume's first
        elevations already run."""
        e, v = step_no[0] % E, (step_no[0] // E) & 1
        step_no[0] += 1
        if world > 1 and e == 0 and gather_done[v] is not None:
            stream.wait_event(gather_done[v])  # everybody is done with this volume buffer (two volumes ago)
        if fused:
            off = (rank * E + e) * slot_bytes  # this rank's slice [rank][e] of a volume buffer
            chain.set_product_mirrors([peer_ptrs[v][r] + off for r in range(world) if r != rank])
            chain.process_device(d_in.data_ptr(), S, peer_ptrs[v][rank] + off, stream.cuda_stream)
        else:
            chain.process_device(d_in.data_ptr(), S, vol[v][e].data_ptr(), stream.cuda_stream)
        if world > 1 and e == E - 1:
            gather(v)

    def flush():
        """End of a timed region: gather the volume that is still filling, wait for the side stream."""
        if world > 1:
            if step_no[0] % E:
                gather((step_no[0] // E) & 1)
                step_no[0] += E - step_no[0] % E  # the next step starts a fresh volume
            stream.wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 1)):
        step()
    flush()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    chain.profile_read(reset=True)
    chain.profile_enable(True)
    l0 = chain.launch_count
    g0 = n_gathers[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    flush()  # the timed region ends after the last gather
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = chain.launch_count - l0
    n_gathers_timed = n_gathers[0] - g0
    chain.profile_enable(False)
    prof = chain.profile_read(reset=True)
    ms = _max_over_ranks(ms, dev, world)
    value = world * S * steps / (ms * 1e-3)
    chain_kernel = chain.chain_kernel
    if fused:
        chain.set_product_mirrors(())
        # the volume buffers the kernels filled over NVLink against an NCCL all-gather of each rank's own slice
        for v in range(2):
            check = torch.empty_like(gathered[v])
            dist.all_gather_into_tensor(check, gathered[v][rank].contiguous())
            if not torch.equal(check, gathered[v]) or not torch.isfinite(gathered[v][:, 0, :, 1:]).all():
                raise SystemExit("bench.py: the fused gather left a wrong product volume")
        del check
    elif world > 1:  # the last gathered volume holds every rank's products
        v = ((step_no[0] - 1) // E) & 1
        if not torch.equal(gathered[v][rank], vol[v]) or not torch.isfinite(gathered[v][:, 0, :, 1:]).all():
            raise SystemExit("bench.py: all-gathered product volume is wrong")

    # ---- sustained leg: the same step back to back for >= 2.5 s, clocks and power sampled meanwhile ----
    run = lambda: chain.process_device(d_in.data_ptr(), S, d_out[0].data_ptr(), stream.cuda_stream)
    n_sus = max(int(args.sustain_seconds / (ms / steps * 1e-3)), steps)
    sus_sampler = ClockSampler(local_rank)
    sus_sampler.start()
    time.sleep(0.25)
    ms_sus = _max_over_ranks(_timed_launches(run, n_sus, stream), dev, world)
    time.sleep(0.15)
    sus_clocks = sus_sampler.stop()
    sustained = {"value": world * S * n_sus / (ms_sus * 1e-3), "unit": "sectors/s", "seconds": ms_sus * 1e-3,
                 "launches": n_sus, "hbm_frac": S * n_sus / (ms_sus * 1e-3) * ALGO_BYTES_C64 / 1e9 / peak,
                 "clocks": sus_clocks}

    # ---- side figures: the other forms of the same chain on the same resident batch (N = 1 only) ----
    alt = {}
    if world == 1:
        for name, cfg in (("two_kind_queue_energy_form", {"chain_impl": wrp.CHAIN_QUEUE}),
                          ("two_kind_queue_doppler_fft", {"doppler_form": wrp.DOPPLER_FFT})):
            with wrp.RadarChain(local_rank, max_batch=args.host_piece, **cfg) as alt_chain:
                f = lambda: alt_chain.process_device(d_in.data_ptr(), S, d_out[1].data_ptr(), stream.cuda_stream)
                for _ in range(3):
                    f()
                v = S * steps / (_timed_launches(f, steps, stream) * 1e-3)
                alt[name] = {"value": v, "unit": "sectors/s", "kernel": alt_chain.chain_kernel, "config": cfg,
                             "hbm_frac": v * ALGO_BYTES_C64 / 1e9 / peak}
    run()
    torch.cuda.synchronize()
    host = d_out[0].cpu().numpy()
    if not np.isfinite(host[:, 1:, :]).all():
        raise SystemExit("bench.py: non-finite products")

    # ---- e2e: host wire-format buffers through wrp_process_host -------------------------------
    wire_chain = wrp.RadarChain(local_rank, input_fmt=wrp.FMT_WIRE_I16BE, max_batch=args.host_piece,
                                n_streams=args.streams)
    S2 = args.e2e_sectors
    wire_np = synth.make_batch(M, N, S2, fmt="wire", first_sector=rank * 7, distinct=4)
    pin_in = wrp.PinnedBuffer(wire_np.nbytes)
    pin_in.array[:] = wire_np.reshape(-1)
    out_e2e = np.empty((S2, M // 2, 2), np.float32)
    for _ in range(max(warmup, 1)):
        wire_chain.process_host(pin_in, S2, out_e2e)
    barrier()
    l1 = wire_chain.launch_count
    t0 = time.perf_counter()
    for _ in range(steps):
        wire_chain.process_host(pin_in, S2, out_e2e)
    torch.cuda.synchronize()
    dt = _max_over_ranks(time.perf_counter() - t0, dev, world)
    launches_e2e = wire_chain.launch_count - l1
    e2e_value = world * S2 * steps / dt

    # ---- the ceiling of that leg: plain cudaMemcpyAsync of the same pinned bytes, all ranks at once ----
    d_wire = torch.empty(wire_np.nbytes, dtype=torch.uint8, device=dev)
    h_wire = torch.from_numpy(pin_in.array)  # a view of the pinned buffer
    copy = lambda: d_wire.copy_(h_wire, non_blocking=True)
    copy()
    barrier()
    ms_copy = _max_over_ranks(_timed_launches(copy, 5, stream), dev, world) / 5
    h2d_ceiling_gbs = wire_np.nbytes / (ms_copy * 1e-3) / 1e9  # per GPU, with every rank copying

    # ---- BASELINE config 2 ("single sector, all stages fused"): latency of ONE sector, N = 1 only ----
    single = None
    if world == 1:
        sector_floats = C * M * N * 2
        one = lambda k: chain.process_device(d_in.data_ptr() + (k * 7 % S) * sector_floats * 4, 1, d_out[0].data_ptr(),
                                             stream.cuda_stream)
        for k in range(5):
            one(k)
        dev_us = []
        for k in range(60):  # a different resident sector every time: its 12.6 MB are not in L2
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            one(k)
            b.record(stream)
            b.synchronize()
            dev_us.append(a.elapsed_time(b) * 1e3)
        out_one = np.empty((1, M // 2, 2), dtype=np.float32)
        host_us = []
        for k in range(40):  # wire records in pinned host memory -> products in host memory, wall clock
            view = pin_in.array[(k * 7 % S2) * M * N * 12:][:M * N * 12]
            t0 = time.perf_counter()
            wire_chain.process_host(view, 1, out_one)
            host_us.append((time.perf_counter() - t0) * 1e6)
        single = {"device_us": statistics.median(dev_us), "device_us_min": min(dev_us),
                  "host_to_host_us": statistics.median(host_us), "host_to_host_us_min": min(host_us),
                  "note": "one 1024x512x3 sector per call: HBM-resident planar input (CUDA events around one launch, input "
                          "not in L2) and wire records from pinned host memory through wrp_process_host (wall clock, "
                          "H2D 6.3 MB + one launch + D2H 4 KiB)"}

    # ---- side figure: the same batch resident in HBM in the wire format (int16 ingest) ----
    d_out2 = torch.empty((S2, M // 2, 2), dtype=torch.float32, device=dev)
    fw = lambda: wire_chain.process_device(d_wire.data_ptr(), S2, d_out2.data_ptr(), stream.cuda_stream)
    for _ in range(3):
        fw()
    wire_resident = S2 * steps / (_timed_launches(fw, steps, stream) * 1e-3)
    clocks = sampler.stop()
    wire_kernels = wire_chain.info.kernels_per_chunk
    del d_wire, d_out2
    wire_chain.close()

    # cross-check: the wire path and the planar path see the same sectors -> same products
    n_chk = min(S, S2, 4)
    if not np.allclose(out_e2e[:n_chk, 1:], host[:n_chk, 1:], rtol=0, atol=1e-3):
        raise SystemExit("bench.py: e2e (wire) and HBM-resident (planar) products disagree")
    pin_in.close()

    # ---- BASELINE config 5 (stress shape) and config 4 (volume scan) at this N ----
    stress = leg_stress(wrp, args, dev, world, local_rank, stream, peak) if args.stress_sectors > 0 else None
    volume = leg_volume(wrp, args, dev, world, rank, local_rank, args.volume_steps) if args.volume_steps > 0 else None

    if rank == 0:
        n_k, ms_k = max(int(prof.n_chain), 1), prof.ms_chain
        sectors_per_launch = S * steps / n_k
        launch_ms = ms_k / n_k
        achieved = sectors_per_launch * ALGO_BYTES_C64 / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "latest_summary.json")) as f:
                traffic = json.load(f).get(chain_kernel + "_dram_bytes_per_launch")
        except Exception:
            pass
        # reported on rank 0 at N = 1 only (torchrun also pins OMP_NUM_THREADS=1 on its workers)
        cpu = None
        if world == 1:
            cpu = cpu_baseline_port(max(4 * (os.cpu_count() or 1), 16),
                                    target_seconds=12.0 if args.cpu_sample <= 0 else 0.5)
        line = {
            "metric": "sectors_per_s", "value": value, "unit": "sectors/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG),
            "run": {"sectors_per_step_per_gpu": S, "input_fmt": "c64_planar", "chunk_sectors": int(info.chunk_sectors),
                    "l2": f"input batch {d_in.numel() * 4 / 1e6:.0f} MB per GPU > L2 {info.l2_bytes / 1e6:.0f} MB, no flush needed",
                    "parallelism": f"sectors sharded over {world} GPU(s); " + (
                        "fused gather: the chain kernel's epilogue stores every product into all ranks' volume buffers "
                        "(peer-mapped, NVLink) itself; per volume (9 elevations x 143 sectors per GPU) one 4-byte all-reduce "
                        "on a side stream as the completion handshake; the timed region ends after the last handshake; the "
                        "volume buffers are checked against an NCCL all-gather afterwards" if fused else
                        "the product volume (9 elevations x 143 sectors per GPU) is all-gathered once per volume on a side "
                        "stream, overlapped with the next volume's kernels; the timed region ends after the last gather"),
                    "gather": "fused" if fused else ("nccl" if world > 1 else None),
                    "gathers_in_timed_region": n_gathers_timed},
            "iq_gbs": value * ALGO_BYTES_C64 / 1e9,
            "chain_hbm_frac": value / world * ALGO_BYTES_C64 / 1e9 / peak,
            "sustained": sustained,
            "wire_resident": {"value": wire_resident, "unit": "sectors/s per GPU", "input_fmt": "wire_i16be",
                              "hbm_frac": wire_resident * ALGO_BYTES_WIRE / 1e9 / peak, "kernels_per_launch": int(wire_kernels),
                              "note": "HBM-resident int16 wire sectors, decoded on the streaming kernel's load path (no decode pre-pass)"},
            "chain_forms": {"default": {"kernel": chain_kernel,
                                        "note": "streaming kernel: range tiles folded into per-gate sums, no range->Doppler "
                                                "hand-off; stages 03-08 in energy form (Parseval)"},
                            **alt},
            "single_sector": single,
            "stress": stress,
            "volume": volume,
            "roofline": {"bound": "hbm", "kernel": chain_kernel, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": sectors_per_launch * ALGO_BYTES_C64,
                         "mean_launch_ms": launch_ms, "kernel_share_of_step": ms_k / ms,
                         "traffic_source": "profiles/latest_summary.json (ncu --set full of this command, dram__bytes_read+write per launch)"},
            "e2e": {"value": e2e_value, "unit": "sectors/s", "h2d_bytes_per_step": S2 * M * N * 12,
                    "d2h_bytes_per_step": S2 * M * 4, "input_fmt": "wire_i16be", "sectors_per_step_per_gpu": S2,
                    "h2d_gbs": e2e_value / world * M * N * 12 / 1e9, "api": "wrp_process_host",
                    "h2d_ceiling_gbs": h2d_ceiling_gbs,
                    "frac_of_h2d_ceiling": (e2e_value / world * M * N * 12 / 1e9) / h2d_ceiling_gbs,
                    "h2d_ceiling_note": "plain cudaMemcpyAsync of the same pinned buffer, every rank copying at once, per GPU",
                    "host_cores_bound_per_rank": len(numa_cores)},
            "gpu_launches": int(launches + launches_e2e),
            "clocks": clocks,
            "cpu_baseline": cpu,
            "reference_gpu": reference_gpu_leg() if world == 1 and not args.skip_reference_gpu else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_volume(args):
    """`--workload volume`: only the config-4 leg, as its own JSON line (strong scaling)."""
    import torch
    import torch.distributed as dist

    wrp = importlib.import_module("weather-radar-processing_b200")
    world, rank, local_rank = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libwrp has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        wrp.bind_host_to_gpu(local_rank)
        dist.init_process_group("nccl", device_id=dev)
    vol = leg_volume(wrp, args, dev, world, rank, local_rank, max(args.steps, 1))
    if rank == 0:
        print(json.dumps({
            "metric": "sectors_per_s", "value": vol["value"], "unit": "sectors/s", "n_gpus": world, "steps": args.steps,
            "warmup": 1, "ms_per_step": vol["ms_per_volume"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG, workload="volume scan 9 elevations x 143 sectors, wire int16 from pinned host memory"),
            "e2e": {"value": vol["value"], "unit": "sectors/s", "h2d_bytes_per_step": vol["units"] * M * N * 12,
                    "d2h_bytes_per_step": 0}, "volume": vol, "gpu_launches": vol["gpu_launches"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sectors", type=int, default=143, help="sectors per GPU per step (HBM-resident leg)")
    ap.add_argument("--e2e-sectors", type=int, default=143, help="sectors per GPU per step (host leg)")
    ap.add_argument("--host-piece", type=int, default=8, help="sectors per pinned-ring piece")
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--sustain-seconds", type=float, default=2.5, help="length of the sustained leg")
    ap.add_argument("--stress-sectors", type=int, default=64,
                    help="sectors of the 4096x1024 stress shape kept resident per GPU (BASELINE config 5; 0 = skip)")
    ap.add_argument("--volume-steps", type=int, default=3,
                    help="volume scans timed for BASELINE config 4 (9 x 143 wire sectors from pinned host memory; 0 = skip)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N > 1: product volume stored into every rank's buffer by the kernels (fused) or all-gathered by NCCL")
    ap.add_argument("--workload", default="sector", choices=["sector", "volume"],
                    help="sector: the default line (every leg); volume: only config 4 as its own line")
    ap.add_argument("--cpu-sample", type=int, default=0, help="non-zero: shorten the CPU baseline leg (profiling runs)")
    ap.add_argument("--skip-reference-gpu", action="store_true",
                    help="do not run the reference's cuFFT program as a side figure (profiling runs: it is a child process)")
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 untimed warm-up steps
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "volume":
        run_volume(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
