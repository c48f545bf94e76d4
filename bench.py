#!/usr/bin/env python
"""bench.py — fingerprinted audio-seconds per second (BASELINE.json metric M1).

Workload (BASELINE.json configs[1]): batch fingerprinting of 1,000 synthetic 3-minute
44.1 kHz mono int16 tracks per GPU, fan 15, amp_min 10, 4096-point window, 50 % overlap.
A "step" is one pass of the whole path (K1 STFT->dB, K2 peaks, K3 pairs+SHA-1) over the
batch.

  value : PCM already resident in HBM, digests left in HBM (CUDA events, max over ranks)
  e2e   : the same call through the public host API — pinned host PCM in, digests out to
          pinned host memory, H2D and D2H inside the timed region
  roofline : the dominant kernel (K1) against the measured HBM copy peak
  cpu_baseline / --impl reference : the reference's CPU algorithm (oracle port:
          numpy FFT + scipy.ndimage + hashlib, one process per track like
          fingerprint_directory's Pool) on this box's host cores, bounded sample.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 44100
TRACK_SAMPLES = 7_938_000          # 3 min
FAN, AMP_MIN = 15, 10
K1_BYTES_PER_AUDIO_S = 264_684     # 88 200 B PCM read + 176 484 B float32 spectrogram write


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=1000, help="tracks per GPU per step")
    ap.add_argument("--track-samples", type=int, default=TRACK_SAMPLES)
    ap.add_argument("--compute", default="f64", choices=["f64", "f32"])
    ap.add_argument("--chunk-frames", type=int, default=262144)
    ap.add_argument("--host-pool-tracks", type=int, default=250)
    ap.add_argument("--cpu-sample-tracks", type=int, default=0, help="0 = 4 per host core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic tracks (SURVEY.md §8d Config 2 recipe), generated on the GPU
# ------------------------------------------------------------------------------------------
def synth_tracks_gpu(dev, first_seed: int, n_tracks: int, n_samples: int, out):
    """6 voices, note change every 11 025 samples, pitch 110*2^(U{0..59}/12), amplitude
    U(500, 4000), harmonics (1, 1/2, 1/4), + N(0, 200^2) noise, clipped to int16."""
    import torch
    seg = 11025
    nseg = -(-n_samples // seg)
    t = torch.arange(n_samples, device=dev, dtype=torch.float64) / FS
    seg_idx = torch.arange(n_samples, device=dev) // seg
    for i in range(n_tracks):
        g = torch.Generator(device=dev)
        g.manual_seed(first_seed + i)
        x = torch.zeros(n_samples, device=dev, dtype=torch.float64)
        semis = torch.randint(0, 60, (6, nseg), device=dev, generator=g)
        amps = torch.rand((6, nseg), device=dev, generator=g, dtype=torch.float64) * 3500 + 500
        for v in range(6):
            f = 110.0 * torch.pow(2.0, semis[v].double() / 12.0)
            ph = (2 * np.pi) * f[seg_idx] * t
            a = amps[v][seg_idx]
            x += a * (torch.sin(ph) + 0.5 * torch.sin(2 * ph) + 0.25 * torch.sin(3 * ph))
        x += torch.randn(n_samples, device=dev, generator=g, dtype=torch.float64) * 200
        out[i].copy_(torch.clamp(torch.round(x), -32768, 32767).to(torch.int16))


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle port, one process per track (fingerprint_directory's Pool)
# ------------------------------------------------------------------------------------------
_CPU_TRACKS = None


def _cpu_worker(i):
    from oracle import sia_oracle as O
    hs = O.fingerprint(_CPU_TRACKS[i], Fs=FS, fan_value=FAN, amp_min=AMP_MIN)
    return len(hs)


def cpu_reference_run(tracks, procs: int, steps: int = 1, warmup: int = 0):
    """Returns (audio_s_per_s, seconds_per_step, hashes) timing `steps` passes over `tracks`."""
    import multiprocessing as mp
    global _CPU_TRACKS
    _CPU_TRACKS = tracks
    ctx = mp.get_context("fork")
    audio_s = sum(len(t) for t in tracks) / FS
    with ctx.Pool(procs) as pool:
        for _ in range(warmup):
            list(pool.imap_unordered(_cpu_worker, range(min(len(tracks), procs))))
        t0 = time.perf_counter()
        nh = 0
        for _ in range(steps):
            nh = sum(pool.imap_unordered(_cpu_worker, range(len(tracks))))
        dt = (time.perf_counter() - t0) / max(steps, 1)
    return audio_s / dt, dt, nh


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(gpu_index: int):
    """Multi-GPU runs: pin this rank to the CPUs NVML lists as local to its GPU, so the pinned host buffers of
    the e2e leg are first-touched on the GPU's own NUMA node (8 ranks sharing one socket's memory and PCIe
    root halve each other's copy bandwidth).  Returns the number of CPUs bound, or None if NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    audio_s_per_track = args.track_samples / FS
    config = {"workload": f"configs[1]: batch fingerprinting of {args.tracks} synthetic {audio_s_per_track:.0f}-s "
                          f"44.1 kHz mono int16 tracks per GPU (fan {FAN}, amp_min {AMP_MIN}, wsize 4096, overlap 0.5)",
              "tracks_per_gpu": args.tracks, "track_samples": args.track_samples, "fan_value": FAN,
              "amp_min": AMP_MIN, "sharding": "by track, no collective", "l2": "inputs larger than L2 (15.9 GB PCM per step)"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import sia_oracle as O
        procs = os.cpu_count() or 1
        ntr = args.cpu_sample_tracks or max(procs, 8)
        tracks = [O.synth_track(10_000 + i, args.track_samples) for i in range(ntr)]
        v, dt, nh = cpu_reference_run(tracks, procs, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
        sample = (f"{ntr} of the {args.tracks} tracks per step ({ntr * audio_s_per_track:.0f} audio-s), "
                  f"Pool({procs}) one task per track like fingerprint_directory (__init__.py:341,357)")
        print(json.dumps({
            "impl": "reference", "metric": "fingerprint_audio_seconds_per_second", "value": v, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "hashes_per_step_sample": nh}))
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from shazam_b200 import _native as N
    from shazam_b200.fingerprinter import Fingerprinter

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    host_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, L = args.tracks, args.track_samples
    stride = (L + 7) // 8 * 8
    d_pcm = torch.empty(B * stride, dtype=torch.int16, device=dev)
    rows = d_pcm.view(B, stride)
    t_gen = time.perf_counter()
    synth_tracks_gpu(dev, 1_000_000 * rank, B, L, [rows[i, :L] for i in range(B)])
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    starts = np.arange(B, dtype=np.int64) * stride
    lens = np.full(B, L, np.int64)

    fp = Fingerprinter(local_rank, max_chunk_frames=args.chunk_frames)
    p = fp.params(Fs=FS, fan_value=FAN, amp_min=AMP_MIN, compute=args.compute)
    frames_per_track = N.num_frames(L)
    cap = int(B * frames_per_track * 5.5 * (FAN - 1))           # ~4 peaks/frame on this signal class
    d_hash = torch.empty((cap, 10), dtype=torch.uint8, device=dev)
    d_t1 = torch.empty(cap, dtype=torch.int32, device=dev)
    audio_s = B * audio_s_per_track

    def step_device():
        return fp.fingerprint_device(d_pcm, starts, lens, p, out=(d_hash, d_t1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        res = step_device()
    n_hashes = len(res.t1)
    fp.timing(True)                                   # reset + enable per-kernel CUDA-event timers
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    kms, klaunch = fp.timing(False)
    t_dev = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_step = float(t_dev.item()) / args.steps
    value = world * audio_s / (ms_step * 1e-3)

    # ---- end to end through the host API ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pool = min(args.host_pool_tracks, B)
        h_pcm = torch.empty(pool * stride, dtype=torch.int16).pin_memory()
        h_pcm.copy_(d_pcm[: pool * stride])
        h_starts = (np.arange(B, dtype=np.int64) % pool) * stride      # the batch cycles over the pinned pool
        h_hash = torch.empty((cap, 10), dtype=torch.uint8).pin_memory()
        h_t1 = torch.empty(cap, dtype=torch.int32).pin_memory()

        def step_host():
            return fp.fingerprint_host(h_pcm, h_starts, lens, p, out=(h_hash, h_t1))
        for _ in range(max(min(args.warmup, 2), 1)):
            r = step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = step_host()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dt = float(t_e.item())
        n_out = len(r.t1)
        e2e = {"value": world * audio_s / dt, "unit": "audio-s/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(B * L * 2), "d2h_bytes_per_step": int(n_out * 14 + (B + 1) * 8),
               "api": "Fingerprinter.fingerprint_host -> sia_fingerprint_batch_host (pinned host PCM in, "
                      "digests + offsets out to pinned host memory)",
               "host_pool_tracks": pool}
        if host_cpus:
            e2e["host_cpus_bound_per_rank"] = host_cpus
        del h_pcm, h_hash, h_t1

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1) ----------------------------------------------
    peak, peak_src = measured_peak_gbs()
    k1_ms = kms[0]
    k1_launches = max(klaunch[0], 1)
    algo_bytes_per_launch = K1_BYTES_PER_AUDIO_S * audio_s * args.steps / k1_launches
    achieved = algo_bytes_per_launch / (k1_ms / k1_launches * 1e-3) / 1e9 if k1_ms > 0 else None
    names = ["stft_db(K1)", "peaks_bitmap(K2)", "peaks_compact(K2)", "pairs_sha1(K3)", "scans"]
    kernel_ms = {n: round(m / args.steps, 4) for n, m in zip(names, kms)}
    roofline = {"bound": "hbm", "kernel": "stft_db_kernel<%s>" % ("double" if args.compute == "f64" else "float"),
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes_per_launch, "launches_timed": k1_launches,
                "avg_launch_ms": k1_ms / k1_launches, "kernel_ms_per_step": kernel_ms,
                "share_of_step": round(k1_ms / max(sum(kms), 1e-9), 4)}
    if args.compute == "f64" and k1_ms > 0:
        # what actually bounds K1: the FP64 pipe (+ the integer pipe).  854 DP arithmetic instructions
        # per thread and frame (static SASS count: 455 DADD, 249 DFMA, 150 DMUL) x 4 warps = 3 416 DP warp
        # instructions per frame; a B200 SM sub-partition issues one DP warp instruction every 2 cycles
        # (tools/ubench/fp64_rate.cu: 2.13) -> 148 x 4 / 2 per clock at the sampled SM clock
        dp_warp_inst = 3416.0 * B * frames_per_track * args.steps
        clk = (clocks or {}).get("sm_mhz") or 1965.0
        dp_peak = 148 * 4 / 2 * clk * 1e6
        roofline["fp64_pipe"] = {"achieved_warp_inst_per_s": dp_warp_inst / (k1_ms * 1e-3), "peak_warp_inst_per_s": dp_peak,
                                 "frac": dp_warp_inst / (k1_ms * 1e-3) / dp_peak,
                                 "note": "K1 computes in float64 (1e-3 dB bound on every bin); 3416 DP warp instructions "
                                         "per frame (SASS), DP issue rate 1 per 2 cycles per sub-partition (tools/ubench)"}
    prof = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(prof):
        try:
            per_frame = json.load(open(prof)).get("dram_bytes_per_frame")
            frames = B * frames_per_track * args.steps / k1_launches
            roofline["traffic"] = per_frame * frames if per_frame else None
            roofline["traffic_source"] = "profiles/k1_traffic.json (ncu --set full, dram read+write per frame) x frames per launch"
        except Exception:
            pass

    # ---- CPU baseline on a bounded sample of the same tracks ----------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        ntr = args.cpu_sample_tracks or min(B, 4 * procs)
        tracks = [rows[i, :L].cpu().numpy() for i in range(ntr)]
        v, dt, nh = cpu_reference_run(tracks, procs)
        cpu = {"value": v, "unit": "audio-s/s", "cores": procs, "kind": "port",
               "sample": f"first {ntr} of the {B} tracks ({ntr * audio_s_per_track:.0f} audio-s, {dt:.1f} s wall), "
                         f"oracle port of the reference CPU path, Pool({procs}) one task per track"}

    out = {"metric": "fingerprint_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": args.compute, "data": "synthetic",
           "config": config, "clocks": clocks, "e2e": e2e,
           "gpu_launches": int(sum(klaunch)), "hashes_per_step_per_gpu": int(n_hashes),
           "roofline": roofline, "cpu_baseline": cpu, "synth_seconds": round(t_gen, 1)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
