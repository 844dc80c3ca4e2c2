#!/usr/bin/env python
"""Benchmark of the feature-extraction hot path (BASELINE.json metric: IQ frames/s, all 18 features,
2048-sample complex128 frames).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # the CUDA path (one rank per GPU)
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path over one batch of synthetic frames: BASELINE config 2,
6 modulations x 16 SNRs x 500 frames x 2048 samples = 48,000 frames (1.573 GB, > the 126 MB L2, so
every step streams from HBM) PER GPU (weak scaling: every rank owns such a batch, no collective on
the data path).  Prints ONE JSON line on rank 0:

  value / ms_per_step   K steps between one CUDA-event pair, max over ranks (device-resident input)
  roofline              algorithmic bytes per step / average step time / measured HBM peak; DRAM traffic from the
                        committed ncu capture while it still belongs to these kernel sources
  sustained             >= 5 s of back-to-back steps with the NVML clock / power record
  e2e                   the same metric through amc_extract_host with pinned HOST buffers (copies inside the timed
                        region), the copy-only ceiling of the same buffers, complex64 transport
  strong_scaling        BASELINE config 3 at its real size (2.52 M frames, 82.6 GB over the ranks), ms per pass and the
                        SHA-256 of the gathered feature matrix (must equal the recorded N = 1 hash)
  gather (N > 1)        NCCL all_gather vs the fused peer-store epilogue
  stage (N = 1)         run_extraction(cfg) on a 1.57 GB all_modulations.mat, wall clock
  cpu_baseline, cpu_baseline_as_shipped (N = 1)   the oracle port on all cores / in the reference's own 6-process shape
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

N_MODS, N_SNR, N_FRAMES, FRAME = 6, 16, 500, 2048
BYTES_PER_FRAME = FRAME * 16 + 18 * 8  # algorithmic bytes: one read of the c128 frame + 18 f64 out
METRIC = "IQ frames/sec (18 features, 2048 samples)"
SNRS = [-10.0 + 2.0 * i for i in range(N_SNR)]


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _physical_gpu_index(local_rank: int) -> int:
    """NVML index of the GPU torch calls cuda:<local_rank> (CUDA_VISIBLE_DEVICES may renumber them)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


# ------------------------------------------------------------------------------------------
# clocks sampler (pynvml), runs during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.power = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:  # noqa: BLE001
                    pass
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {
            "sm_mhz": float(np.median(self.samples)),
            "sm_min_mhz": int(min(self.samples)),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
            "power_w_max": float(max(self.power)) if self.power else None,
        }


# ------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/): the reference algorithm restated in numpy
# ------------------------------------------------------------------------------------------
def _cpu_cell(args):
    """Worker: features of `nf` frames of one (modulation, snr) cell with the reference's work
    pattern (every feature function rebuilding its inputs, scipy kurtosis)."""
    mod, si, nf, seed = args
    from amcpy_b200 import synth
    from oracle import amc_oracle as orc

    x = synth.cell(mod, SNRS[si], si, range(nf), FRAME, seed)
    ids = list(range(1, 19))
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        for f in range(nf):
            orc.calculate_features_faithful(ids, x[f])
    return time.perf_counter() - t0


def cpu_pass(pool, procs: int, frames_per_cell: int, seed: int):
    """One bounded pass: 6 x 16 cells x frames_per_cell frames, spread over `procs` processes.
    Returns (frames, seconds of wall clock for the feature work incl. pool overhead)."""
    jobs = [(m, s, frames_per_cell, seed) for m in range(N_MODS) for s in range(N_SNR)]
    t0 = time.perf_counter()
    pool.map(_cpu_cell, jobs, chunksize=max(1, len(jobs) // (procs * 4)))
    # data generation happens inside the workers; subtract nothing: generation is ~1% of the work
    return len(jobs) * frames_per_cell, time.perf_counter() - t0


def run_cpu_baseline(target_s: float = 12.0):
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        cpu_pass(pool, procs, 1, 7)  # warm imports
        f0, s0 = cpu_pass(pool, procs, 4, 8)          # measured rate of this host
        fpc = max(1, min(N_FRAMES, int((f0 / s0) * target_s / (N_MODS * N_SNR))))   # ~target_s of wall clock
        frames, secs = cpu_pass(pool, procs, fpc, 11)
    return {
        "value": frames / secs,
        "unit": "frames/s",
        "cores": procs,
        "kind": "port",
        "sample": f"{frames} frames ({N_MODS}x{N_SNR}x{fpc} of the 6x16x500x2048 c128 set), oracle faithful form, "
                  f"{procs} processes, {secs:.1f}s wall",
    }


# The reference's extraction stage AS SHIPPED (feature_extraction.py:22-99): one process per modulation, each of
# which loads the WHOLE .mat file (:46-47), takes its variable as the Fortran-ordered array loadmat returns (:48) and
# feeds zero-copy strided views of it (:68) through an unbounded Queue to `num_threads` daemon threads (:58-61) that
# write float32 rows (:35); the parent starts all six and joins them (:89-97).  The per-frame function is the oracle's
# faithful restatement of calculate_features.  config.py:98 ships num_threads = 8; 1 is the GIL-free setting.
def _shipped_worker_thread(q, fm, ids):
    from oracle import amc_oracle as orc

    while True:
        item = q.get()
        if item is None:
            q.task_done()
            return
        sig, si, fi = item
        with np.errstate(all="ignore"):
            fm[si, fi, :] = orc.calculate_features_faithful(ids, sig)
        q.task_done()


def _shipped_process(path, key, n_snr, n_frames, frame_size, num_threads):
    import queue
    import threading as th

    import scipy.io

    parsed = scipy.io.loadmat(path)[key]                       # the whole file, once per process
    fm = np.zeros((n_snr, n_frames, 18), dtype=np.float32)
    q = queue.Queue()
    ids = list(range(1, 19))
    workers = [th.Thread(target=_shipped_worker_thread, args=(q, fm, ids), daemon=True) for _ in range(num_threads)]
    for w in workers:
        w.start()
    for si in range(n_snr):
        for fi in range(n_frames):
            q.put((parsed[si, fi, 0:frame_size], si, fi))
    q.join()
    for _ in workers:
        q.put(None)


def run_cpu_baseline_as_shipped(frames_per_cell: int = 48):
    import multiprocessing as mp
    import tempfile

    from amcpy_b200 import synth
    from amcpy_b200.config import SignalConfig

    info = SignalConfig().mat_info
    res = {"kind": "port", "shape": "6 processes x num_threads threads, Fortran-order views, 6 x loadmat "
                                    "(feature_extraction.py:42-99)",
           "cores": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "all_modulations.mat")
        synth.write_all_modulations_mat(path, synth.dataset(SNRS, frames_per_cell, FRAME, 31), info)
        ctx = mp.get_context("fork")
        frames = N_MODS * N_SNR * frames_per_cell
        for nt in (8, 1):
            procs = [ctx.Process(target=_shipped_process, args=(path, info[m], N_SNR, frames_per_cell, FRAME, nt))
                     for m in synth.MODULATIONS]
            t0 = time.perf_counter()
            for p_ in procs:
                p_.start()
            for p_ in procs:
                p_.join()
            secs = time.perf_counter() - t0
            ok = all(p_.exitcode == 0 for p_ in procs)
            res[f"num_threads_{nt}"] = {"value": frames / secs if ok else None, "unit": "frames/s", "seconds": secs,
                                        "frames": frames}
    res["sample"] = f"{frames} frames (6x16x{frames_per_cell} of the 6x16x500x2048 c128 set) per setting"
    return res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        cpu_pass(pool, procs, 1, 7)
        f0, s0 = cpu_pass(pool, procs, 2, 8)          # measured rate of this host
        # each step: a bounded sample; the whole run (warmup + steps) is kept under ~2 minutes
        step_budget = min(3.0, 120.0 / max(1, args.steps + args.warmup))
        fpc = max(1, min(N_FRAMES, int((f0 / s0) * step_budget / (N_MODS * N_SNR))))
        for w in range(args.warmup):
            cpu_pass(pool, procs, fpc, 100 + w)
        tot_f, tot_s = 0, 0.0
        for k in range(args.steps):
            f, s = cpu_pass(pool, procs, fpc, 200 + k)
            tot_f += f
            tot_s += s
    val = tot_f / tot_s
    sample = (f"each step {N_MODS}x{N_SNR}x{fpc} = {N_MODS * N_SNR * fpc} frames of the 6x16x500x2048 c128 workload, "
              f"reference algorithm restated in numpy (oracle faithful form), {procs} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: 6 mods x 16 SNR x 500 frames x 2048 samples complex128, all 18 features",
                   "frames_per_step": N_MODS * N_SNR * fpc},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------
CONFIG3_SNRS = [-20.0 + 2.0 * i for i in range(21)]
CONFIG3_FRAMES = 20000          # per (modulation, SNR) cell: 6 x 21 x 20,000 = 2.52 M frames, 82.6 GB complex128


def _kernel_source_hash() -> str:
    """sha256 over the CUDA sources: ties committed ncu numbers (DRAM traffic) to the kernels they were taken from."""
    import hashlib

    h = hashlib.sha256()
    for f in sorted((ROOT / "amcpy_b200" / "csrc").glob("*.cu*")):
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def _measured_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` capture - only while the capture belongs to the
    kernels being benchmarked (same source hash); otherwise None (a stale number is worse than none)."""
    tf = ROOT / "profiles" / "traffic_per_launch.json"
    try:
        d = json.loads(tf.read_text())
        if d.get("kernel_source_sha256_16") == _kernel_source_hash():
            return d.get("dram_bytes_per_launch"), d.get("source")
        return None, "profiles/traffic_per_launch.json was captured from different kernel sources: not reported"
    except Exception:  # noqa: BLE001
        return None, "no capture committed"


def _timed_max_over_ranks(fn, reps, dev, dist):
    """ms per call of `fn` (device time, CUDA events on the current stream), max over ranks."""
    import torch

    fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def strong_scaling_config3(rank, world, dev, dist, frames_per_cell=CONFIG3_FRAMES, reps=3):
    """BASELINE config 3 (fixed total work): every rank generates and extracts a contiguous slice of the frame axis of
    every (modulation, SNR) cell; the (2.52 M, 18) matrix gathered on rank 0 is hashed - it must not depend on N."""
    import hashlib

    import torch

    from amcpy_b200 import ops, synth

    if frames_per_cell % world:
        return {"skipped": f"{frames_per_cell} frames per cell do not divide over {world} ranks"}
    per = frames_per_cell // world
    n_cells = N_MODS * len(CONFIG3_SNRS)
    x = synth.dataset_device(N_MODS, CONFIG3_SNRS, per, FRAME, dev, seed=3, first_frame=rank * per)
    out = torch.empty((x.shape[0], 18), dtype=torch.float64, device=dev)
    ms = _timed_max_over_ranks(lambda: ops.extract_features(x, out=out), reps, dev, dist)
    del x
    if dist is not None:
        parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
        dist.gather(out, parts, dst=0)
    else:
        parts = [out]
    res = None
    if rank == 0:
        full = torch.stack([p_.view(n_cells, per, 18) for p_ in parts], dim=1).reshape(n_cells * frames_per_cell, 18)
        sha = hashlib.sha256(full.cpu().numpy().tobytes()).hexdigest()
        frames = n_cells * frames_per_cell
        rec = ROOT / "profiles" / "config3_sha256.json"
        want = None
        try:
            d = json.loads(rec.read_text())
            if d.get("kernel_source_sha256_16") == _kernel_source_hash() and d.get("frames_per_cell") == frames_per_cell:
                want = d.get("sha256_n1")
        except Exception:  # noqa: BLE001
            pass
        res = {
            "workload": f"BASELINE config 3: 6 mods x 21 SNR x {frames_per_cell} frames x 2048 complex128 = "
                        f"{frames} frames, {frames * FRAME * 16 / 1e9:.1f} GB, on-device generator, fixed total work",
            "frames": frames, "ms_per_pass": ms, "frames_per_s": frames / (ms * 1e-3),
            "hbm_frac_per_gpu": frames * BYTES_PER_FRAME / (ms * 1e-3) / 1e9 / world / _peaks()[0],
            "sha256_of_features": sha, "sha256_recorded_at_n1": want,
            "matches_n1": (sha == want) if want else None, "finite": bool(torch.isfinite(full).all()),
        }
    del out, parts
    torch.cuda.empty_cache()
    return res


def gather_timing(rank, world, dev, dist, x, n_frames):
    """The one optional exchange of the path - the (frames, 18) matrix to the consumer's rank - both ways: NCCL
    all_gather_into_tensor after the kernel vs the kernel's own epilogue storing rows into rank 0's matrix through
    NVLink peer mappings (torch symmetric memory).  Every rank's shard = 1/N of this rank's bench batch."""
    import torch

    from amcpy_b200 import ops, sharding

    total = (n_frames // world) * world
    lo, hi = sharding.shard_range(total, rank, world)
    mine = x[lo:hi]
    res = {"frames_total": total, "rows_per_rank": hi - lo}
    try:
        res["extract_only_ms"] = _timed_max_over_ranks(lambda: ops.extract_features(mine), 20, dev, dist)
        res["extract_plus_nccl_all_gather_ms"] = _timed_max_over_ranks(
            lambda: sharding.gather_features(ops.extract_features(mine), total), 20, dev, dist)
        a = sharding.gather_features(ops.extract_features(mine), total)
        res["fused_peer_store_ms"] = _timed_max_over_ranks(
            lambda: sharding.extract_sharded_to_root(mine, total, root=0), 20, dev, dist)
        b = sharding.extract_sharded_to_root(mine, total, root=0)
        torch.cuda.synchronize()
        ok = torch.tensor([1 if (rank != 0 or torch.equal(a, b)) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res["bitwise_equal"] = bool(ok.item())
    except Exception as e:  # noqa: BLE001 - symmetric memory needs P2P access between all ranks
        res["error"] = f"{type(e).__name__}: {str(e)[:200]}"
    return res


def stage_timing(dev, frames_per_cell=N_FRAMES):
    """The drop-in stage end to end on BASELINE config 1: run_extraction(cfg) on a 1.57 GB all_modulations.mat
    (parse + host->device + kernels + device->host + six savemat), wall clock, against the PCIe floor."""
    import tempfile

    import torch

    from amcpy_b200 import synth
    from amcpy_b200.config import Config, Paths, SignalConfig
    from amcpy_b200.feature_extraction import run_extraction

    with tempfile.TemporaryDirectory() as td:
        cfg = Config(paths=Paths(root=Path(td)), signals=SignalConfig(num_frames=frames_per_cell))
        cfg.paths.ensure_dirs()
        x = synth.dataset_device(N_MODS, SNRS, frames_per_cell, FRAME, dev, seed=1).view(N_MODS, N_SNR, frames_per_cell, FRAME)
        synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, x.cpu().numpy(), cfg.signals.mat_info)
        del x
        torch.cuda.empty_cache()
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            run_extraction(cfg)                      # warm: staging buffers, page cache
            times = []
            for _ in range(3):
                t0 = time.perf_counter()
                run_extraction(cfg)
                times.append(time.perf_counter() - t0)
    frames = N_MODS * N_SNR * frames_per_cell
    return {"workload": "BASELINE config 1: run_extraction on all_modulations.mat (6 x 16 x 500 x 2048 complex128, 1.57 GB)",
            "seconds": min(times), "seconds_all": times, "frames_per_s": frames / min(times),
            "input_gbs": frames * FRAME * 16 / min(times) / 1e9}


def run_cuda_arm(args):
    import torch

    from amcpy_b200 import _native as nat
    from amcpy_b200 import ops, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py --gpus {args.gpus} must be launched with torchrun (one rank per GPU)", file=sys.stderr)
            return 2
    if local_rank == 0:
        nat.build_if_missing()   # the .so normally travels with the snapshot; a source-only checkout builds it once
    else:
        for _ in range(600):     # other ranks wait for rank 0's build instead of compiling the same file
            if nat.LIB_PATH.exists():
                break
            time.sleep(0.5)
    nat.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n_frames = N_MODS * N_SNR * N_FRAMES
    # hand-written on-device generator (counter-based Philox): rank r owns frames [r*500, (r+1)*500) of every cell
    # (weak scaling: every rank streams its own 48,000-frame batch)
    x = synth.dataset_device(N_MODS, SNRS, N_FRAMES, FRAME, dev, seed=2024, first_frame=rank * N_FRAMES)
    out = torch.empty((n_frames, 18), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    lib = nat.lib()
    nat.check(lib.amc_init(local_rank))

    torch.cuda.synchronize()
    torch.cuda.profiler.start()   # no-op unless run under `ncu --profile-from-start off`
    for _ in range(max(3, args.warmup)):
        ops.extract_features(x, out=out)
    barrier()

    # ---- device-resident timing: EXACTLY K steps between one CUDA-event pair on the launching stream (barrier +
    # synchronize on both sides).  No events between the steps: the launches are chained by programmatic dependent
    # launch, and an event record between two kernels would serialise what a real caller's stream does not.
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.amc_launch_count()   # counted from here: the generator / warm-up launches are excluded
    physical = _physical_gpu_index(local_rank)
    with ClockSampler(physical) as clocks:
        barrier()
        t_all0.record(stream)
        for k in range(args.steps):
            ops.extract_features(x, out=out)
        t_all1.record(stream)
        barrier()
    torch.cuda.profiler.stop()
    launches = lib.amc_launch_count() - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    # the same step bracketed by its own event pair (what round 1 reported: isolated launches, no overlap)
    n_iso = min(args.steps, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_iso)]
    for k in range(n_iso):
        ev[k][0].record(stream)
        ops.extract_features(x, out=out)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    isolated_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    kernel_ms = total_ms / args.steps     # average duration of a step (fused kernel + careful-path scan) in the timed region
    value = world * n_frames * args.steps / (total_ms * 1e-3)

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"profiling_only": True, "kernel_ms_per_launch": isolated_ms,
                              "kernel_ms_per_launch_chained": kernel_ms, "value": value}), flush=True)
        return 0

    # ---- sustained: >= 5 s of back-to-back steps with the clock / power record (the headline run above is a burst)
    sustained = None
    if not args.quick:
        with ClockSampler(physical) as sclk:
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_sus, t_host0 = 0, time.perf_counter()
            s0.record(stream)
            while time.perf_counter() - t_host0 < args.sustain_seconds:
                for _ in range(200):
                    ops.extract_features(x, out=out)
                n_sus += 200
                torch.cuda.synchronize()
            s1.record(stream)
            barrier()
        sus_ms = torch.tensor([s0.elapsed_time(s1) / n_sus], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(sus_ms, op=dist.ReduceOp.MAX)
        sus_ms = float(sus_ms.item())
        sustained = {"seconds": n_sus * sus_ms * 1e-3, "steps": n_sus, "ms_per_step": sus_ms,
                     "value": world * n_frames / (sus_ms * 1e-3), "unit": "frames/s",
                     "hbm_frac_per_gpu": n_frames * BYTES_PER_FRAME / (sus_ms * 1e-3) / 1e9 / _peaks()[0],
                     "clocks": sclk.summary()}

    # ---- end to end through the public host API: pinned host frames -> features on the host
    # one rank per GPU: allocate (first-touch) the pinned staging buffers on the GPU's own NUMA node
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_cpus = nat.bind_host_thread_to_gpu(physical)
    xh = torch.empty((n_frames, FRAME), dtype=torch.complex128, pin_memory=True)
    xh.copy_(x)
    oh = torch.empty((n_frames, 18), dtype=torch.float64, pin_memory=True)
    xh_np, oh_np = xh.numpy(), oh.numpy()
    e2e_steps = max(2, min(args.steps, 5))

    def wall_max(fn, reps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / reps

    e2e_s = wall_max(lambda: ops.extract_features_host(xh_np, device=local_rank, out=oh_np), e2e_steps)
    e2e_value = world * n_frames / e2e_s
    same = bool(torch.equal(torch.from_numpy(oh_np).to(dev), out))
    # copy-only ceiling of the same transfer: the same pinned buffer to the device and the result back, no kernel,
    # all ranks at once - what amc_extract_host could reach at best on this box at this N
    xd_tmp = torch.empty_like(x)
    od_tmp = torch.empty_like(out)

    def copy_only():
        xd_tmp.copy_(xh, non_blocking=True)
        oh.copy_(od_tmp, non_blocking=True)
        torch.cuda.synchronize()

    od_tmp.copy_(out)
    copy_s = wall_max(copy_only, e2e_steps)
    oh.copy_(out)
    del xd_tmp, od_tmp
    bytes_step = n_frames * FRAME * 16 + n_frames * 18 * 8
    # complex64 transport (captures are 8-bit / 16-bit IQ: complex64 holds them exactly): half the PCIe bytes
    xh64 = torch.empty((n_frames, FRAME), dtype=torch.complex64, pin_memory=True)
    xh64.copy_(x.to(torch.complex64))
    oh64 = np.empty((n_frames, 18), dtype=np.float64)
    e2e64_s = wall_max(lambda: ops.extract_features_host(xh64.numpy(), device=local_rank, out=oh64), e2e_steps)
    del xh64
    if affinity0 is not None and numa_cpus:
        os.sched_setaffinity(0, affinity0)   # the CPU baselines below use every core of the box again

    def guarded(name, fn):
        """An auxiliary leg must never cost the headline line: its failure is reported under its own key.  (Collective
        legs run on every rank, so a failure inside one of them is the same exception on every rank.)"""
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            print(f"bench.py: leg '{name}' failed: {type(e).__name__}: {e}", file=sys.stderr)
            return {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    strong = None if args.quick else guarded("strong_scaling", lambda: strong_scaling_config3(rank, world, dev, dist))
    gather = guarded("gather", lambda: gather_timing(rank, world, dev, dist, x, n_frames)) if (world > 1 and not args.quick) else None
    del xh, oh, x
    torch.cuda.empty_cache()
    stage = guarded("stage", lambda: stage_timing(dev)) if (world == 1 and rank == 0 and not args.quick) else None

    if rank == 0:
        peak, peak_src = _peaks()
        achieved = n_frames * BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = _measured_traffic()
        cpu = guarded("cpu_baseline", run_cpu_baseline) if (world == 1 and not args.no_cpu_baseline) else None
        shipped = (guarded("cpu_baseline_as_shipped", run_cpu_baseline_as_shipped)
                   if (world == 1 and not args.no_cpu_baseline and not args.quick) else None)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "BASELINE config 2: 6 mods x 16 SNR x 500 frames x 2048 samples complex128, all 18 features",
                "frames_per_gpu_per_step": n_frames, "bytes_per_gpu_per_step": n_frames * FRAME * 16,
                "l2": "inputs (1.57 GB per GPU) larger than L2 (126 MB); no flush needed",
                "sharding": f"{world} ranks x 48000 independent frames, no data-path collective",
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "fused16_features_kernel<2048,double2,15> (+ the careful-path scan kernel that follows it)",
                "kernel_ms_per_launch": kernel_ms, "kernel_ms_per_launch_isolated": isolated_ms,
                "timing": "kernel_ms_per_launch = CUDA-event time of the K-step timed region / K (max over ranks); "
                          "_isolated = mean of per-step event pairs (an event between two launches defeats their "
                          "programmatic overlap)",
                "algorithmic_bytes_per_launch": n_frames * BYTES_PER_FRAME,
                "frac_of_nominal_8TBs": achieved / 8000.0,
            },
            "e2e": {
                "value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": n_frames * FRAME * 16,
                "d2h_bytes_per_step": n_frames * 18 * 8, "steps": e2e_steps,
                "api": "amcpy_b200.ops.extract_features_host -> C ABI amc_extract_host (pinned host buffers)",
                "matches_device_path": same, "host_cpus_bound": len(numa_cpus),
                "pcie_gbs_per_gpu": bytes_step / e2e_s / 1e9,
                "copy_only_gbs_per_gpu": bytes_step / copy_s / 1e9,
                "frac_of_copy_only": copy_s / e2e_s,
                "complex64_transport": {"value": world * n_frames / e2e64_s, "unit": "frames/s",
                                        "h2d_bytes_per_step": n_frames * FRAME * 8},
            },
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if sustained is not None:
            line["sustained"] = sustained
        if strong is not None:
            line["strong_scaling"] = strong
        if gather is not None:
            line["gather"] = gather
        if stage is not None:
            line["stage"] = stage
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if shipped is not None:
            line["cpu_baseline_as_shipped"] = shipped
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--quick", action="store_true", help="skip sustained / config-3 / gather / stage / as-shipped legs")
    ap.add_argument("--sustain-seconds", type=float, default=5.0)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else 300
    if args.warmup is None:
        args.warmup = 1 if args.impl == "reference" else 10
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
