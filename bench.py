#!/usr/bin/env python3
"""bench.py — tracks/s of the per-track analysis hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

A "step" is one pass of analyze_audio over one batch of synthetic 3-minute 44.1 kHz tracks
(BASELINE.json configs[1]: 1024 tracks per B200, BPM 70-180, random keys — SURVEY.md §8d C2).  The
batch is generated ON the device before the timed region, so `value` is the device-resident
throughput; `e2e` is the same metric through the host-buffer C-ABI call (pinned host samples, H2D and
result D2H inside the timed region).  One process per GPU under torchrun; tracks are independent, so
ranks share nothing but the barrier and the max-over-ranks time (weak scaling: 1024 tracks per GPU).

The reference arm times the CPU oracle port of the reference algorithm (the Rust crate cannot be built
in this image: no cargo) with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SR = 44100
N_SAMPLES = 7_938_000  # 3 minutes
ALGO_BYTES_PER_TRACK = 4 * N_SAMPLES  # one read of the input samples
FLOPS_KEY_STFT = 15_488 * 532_480  # 8192-point frames, 5 N log2 N convention
FLOPS_BASE_STFT = 15_500 * 112_640  # 2048-point frames at hop 512
FLOPS_ESCALATION = 54_250 * 112_640  # hop 256 + 512 + 1024 re-analysis of an escalated track (SURVEY.md §8d counts all three)
FLOPS_BASE_STFT = 15_500 * 112_640


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(device_index: int) -> str:
    """Pins this process (and the library's upload thread, which inherits the mask) to the CPUs of the NUMA node the GPU hangs
    off, before any pinned host buffer is allocated: pinned pages are placed on the allocating thread's node, and an upload
    that has to cross the socket interconnect runs at about half the PCIe rate.  Returns a short note for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip())
        if node < 0:
            return "gpu numa node unknown (-1)"
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"gpu on numa node {node}, none of its cpus in this process's cpuset"
        os.sched_setaffinity(0, allowed)
        return f"bound to numa node {node} ({len(allowed)} cpus) of GPU {device_index}"
    except Exception as e:  # placement is an optimisation, never a reason to fail the measurement
        return f"not bound ({type(e).__name__})"


def track_params(first: int, count: int) -> np.ndarray:
    import synth

    out = np.zeros((count, 5), np.float32)
    for i in range(count):
        p = synth.c2_params(first + i)
        out[i] = (p.bpm, p.tonic, p.minor, p.phase_frac, p.chord_amp)
    return out


def oracle_batch(samples: np.ndarray, n_tracks: int, jobs: int) -> float:
    """Wall seconds of the CPU oracle over n_tracks concatenated tracks with `jobs` threads
    (mirrors rayon par_iter over tracks, examples/analyze_batch.rs:239-326)."""
    import oracle_lib as O

    L = O.lib(fast=True)
    offs = (np.arange(n_tracks + 1, dtype=np.uint64) * np.uint64(samples.size // n_tracks))
    srs = np.full(n_tracks, SR, np.uint32)
    return float(L.so_batch_timed(O.f32ptr(samples), offs.ctypes.data_as(C.POINTER(C.c_uint64)), srs.ctypes.data_as(C.POINTER(C.c_uint32)), n_tracks, jobs,
                                  None, None, None))


def host_jobs() -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)
    return max(1, min(n - 1, 32))  # default_jobs = CPUs - 1 (examples/analyze_batch.rs:180-185), capped to bound the sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth

    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "-j8"], check=True)
    jobs = host_jobs()
    n = jobs  # one track per worker and step: a bounded sample of the C2 workload
    x = np.concatenate([synth.render(synth.c2_params(i)) for i in range(n)])
    for _ in range(max(args.warmup, 0) and 1):  # one warm-up pass is enough for a CPU code (bounded run time)
        oracle_batch(x, n, jobs)
    times = [oracle_batch(x, n, jobs) for _ in range(args.steps)]
    ms = 1000.0 * sum(times) / len(times)
    val = n / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": "tracks_per_sec_3min_44k1", "value": val, "unit": "tracks/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: batch of synthetic 3-min 44.1 kHz mono tracks, BPM 70-180, random keys (BASELINE.json configs[1]); analyze_audio defaults",
                   "tracks_per_step": n, "note": "CPU oracle port of stratum-dsp 1.0.0 (Rust toolchain absent), std::thread pool over tracks"},
        "cpu_baseline": {"value": val, "unit": "tracks/s", "cores": jobs, "kind": "port", "sample": f"{n} C2 tracks per step, {args.steps} steps"},
        "e2e": {"value": val, "unit": "tracks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=1024, help="tracks per GPU and step (BASELINE.json configs[1]: 1024)")
    ap.add_argument("--e2e-tracks", type=int, default=1024, help="tracks per step of the host-buffer (e2e) leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import stratum_dsp_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not S.LIB_PATH.exists():
        raise SystemExit("libstratum_b200.so is missing — run __graft_entry__.build(); there is no fallback path")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    nt = args.tracks
    # ---- device-resident batch: generated on the device, outside the timed region ----
    buf = torch.empty(nt * N_SAMPLES, dtype=torch.float32, device="cuda")
    S.synth_batch(buf.data_ptr(), nt, N_SAMPLES, SR, track_params(rank * nt, nt), local_rank)
    offsets = np.arange(nt + 1, dtype=np.uint64) * np.uint64(N_SAMPLES)
    srs = np.full(nt, SR, np.uint32)

    waves = [1]

    def step():
        res = S.analyze_batch_device(buf.data_ptr(), offsets, srs, None, local_rank, convert=False)
        ms = S.last_call_device_ms()
        waves[0] = max(1, S.last_call_waves())
        return res, ms

    for _ in range(args.warmup):
        res, _ms = step()
        S.free_results(res)
    S.stage_timing(True)
    S.stage_times(reset=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = S.launch_count()
    t0 = time.perf_counter()
    dev_ms = 0.0
    last = None
    for _ in range(args.steps):
        if last is not None:
            S.free_results(last)
        last, ms = step()
        dev_ms += ms
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    launches = S.launch_count() - launches0
    clocks = sampler.stop()
    stages = S.stage_times(reset=True)
    S.stage_timing(False)
    # sanity on the last step's results (no oracle here: parity lives in tests/ and smoke())
    ok = sum(1 for r in last if r.status == 0)
    bpm_hit = 0
    params = track_params(rank * nt, nt)
    for i, r in enumerate(last):
        b, t = r.bpm, params[i, 0]
        if r.status == 0 and b > 0 and min(abs(b - t), abs(2 * b - t), abs(b - 2 * t)) <= 2.0:
            bpm_hit += 1
    n_escalated = sum(1 for r in last if r.status == 0 and r.tempogram_multi_res_triggered == 1)
    S.free_results(last)
    L = S.lib()
    L.stratum_b200_fp32_peak_tflops.restype = C.c_double
    fp32_peak = float(L.stratum_b200_fp32_peak_tflops(C.c_int32(local_rank)))

    # ---- e2e: host buffers through the reference-facing call, H2D + result D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_tracks if world == 1 else min(args.e2e_tracks, 256), nt)  # N ranks pin N host buffers: keep them modest
        host = None
        while host is None:  # a box with little lockable host memory gets a smaller (still pinned) e2e batch rather than no number
            try:
                host = torch.empty(ne * N_SAMPLES, dtype=torch.float32, pin_memory=True)
            except RuntimeError:
                if ne <= 32:
                    raise
                ne //= 2
        host.copy_(buf[: ne * N_SAMPLES])
        torch.cuda.synchronize()
        hnp = host.numpy()
        eoff = np.arange(ne + 1, dtype=np.uint64) * np.uint64(N_SAMPLES)
        import ctypes as _C
        srs_e = np.full(ne, SR, np.uint32)

        def host_step():
            # the reference-facing C-ABI call itself (include/stratum_b200.h): host samples in, StratumResult array out
            res = (S.StratumResult * ne)()
            st = S.lib().stratum_b200_analyze_batch(hnp.ctypes.data, eoff.ctypes.data_as(_C.POINTER(_C.c_uint64)), srs_e.ctypes.data_as(_C.POINTER(_C.c_uint32)), ne,
                                                    None, (_C.c_int32 * 1)(local_rank), 1, res)
            assert st == 0, S.last_error()
            n_ok = sum(1 for r in res if r.status == 0)
            S.free_results(res)
            return n_ok

        # pinned H2D rate of this box (outside the timed region): the ceiling of the f32 leg is this rate / 31.75 MB per track
        probe = torch.empty(min(ne, 32) * N_SAMPLES, dtype=torch.float32, device="cuda")
        probe.copy_(host[: probe.numel()], non_blocking=True)
        torch.cuda.synchronize()
        tb0 = time.perf_counter()
        probe.copy_(host[: probe.numel()], non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = probe.numel() * 4 / (time.perf_counter() - tb0) / 1e9
        del probe
        host_step()  # warm-up (staging buffer allocation)
        h0, d0 = S.transfer_bytes()
        barrier()
        te0 = time.perf_counter()
        e_steps = max(1, min(args.steps, 3))
        for _ in range(e_steps):
            assert host_step() == ne
        barrier()
        e_ms = (time.perf_counter() - te0) * 1000.0 / e_steps
        h1, d1 = S.transfer_bytes()
        e2e = {"tracks_per_step": ne, "ms_per_step": e_ms, "h2d": (h1 - h0) // e_steps, "d2h": (d1 - d0) // e_steps, "h2d_gbs": h2d_gbs}
        # informational: the decoder-side entry (16-bit PCM uploaded as is, converted on the device): half the H2D bytes
        hnp = None
        del host  # the f32 buffer is unpinned before the PCM one is pinned
        pcm = torch.empty(ne * N_SAMPLES, dtype=torch.int16, pin_memory=True)
        for i in range(0, ne, 8):  # converted in slices: the analysis arenas own most of the device memory
            a, b = i * N_SAMPLES, min(i + 8, ne) * N_SAMPLES
            pcm[a:b].copy_((buf[a:b] * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
        torch.cuda.synchronize()
        pnp = pcm.numpy()
        ptracks = [pnp[i * N_SAMPLES:(i + 1) * N_SAMPLES] for i in range(ne)]
        chans = np.ones(ne, np.uint32)

        def pcm_step():
            res = (S.StratumResult * ne)()
            st = S.lib().stratum_b200_analyze_batch_pcm16(pnp.ctypes.data, eoff.ctypes.data_as(_C.POINTER(_C.c_uint64)), srs_e.ctypes.data_as(_C.POINTER(_C.c_uint32)),
                                                          chans.ctypes.data_as(_C.POINTER(_C.c_uint32)), ne, None, (_C.c_int32 * 1)(local_rank), 1, res)
            assert st == 0, S.last_error()
            S.free_results(res)

        pcm_step()
        barrier()
        tp0 = time.perf_counter()
        for _ in range(e_steps):
            pcm_step()
        barrier()
        e2e["pcm16_ms_per_step"] = (time.perf_counter() - tp0) * 1000.0 / e_steps
        del pcm, ptracks

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "-j8"], check=True)
            jobs = host_jobs()
            sample = buf[: jobs * N_SAMPLES].cpu().numpy()
            sec = oracle_batch(sample, jobs, jobs)
            cpu = {"value": jobs / sec, "unit": "tracks/s", "cores": jobs, "kind": "port",
                   "sample": f"first {jobs} tracks of the batch, one per thread, {sec:.1f} s wall (CPU oracle, -O3)"}
        except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU measurement
            cpu = {"value": None, "unit": "tracks/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    # ---- reduce over ranks: max time, summed units ----
    ms_step = dev_ms / args.steps
    wall_step = wall_ms / args.steps
    e_ms_step = e2e["ms_per_step"] if e2e else 0.0
    p_ms_step = e2e["pcm16_ms_per_step"] if e2e else 0.0
    if dist:
        t = torch.tensor([ms_step, wall_step, e_ms_step, p_ms_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, wall_step, e_ms_step, p_ms_step = (float(v) for v in t.tolist())
        c = torch.tensor([ok, bpm_hit, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        ok, bpm_hit, launches = (int(v) for v in c.tolist())
    if rank == 0:
        peak, peak_src = load_peaks()
        total_tracks = nt * world
        value = total_tracks / (ms_step / 1000.0)
        # dominant kernel of the step, by device time on the launching stream.  Only single-kernel stages qualify: the
        # tail stages of the tempo path (tempograms, final BPM, beats, legacy, energy onsets) run on the second stream beside
        # the key path, so their event spans include time spent waiting for SM slots and say nothing about kernel cost.
        single = {k: v for k, v in stages.items() if k in ("stft_8192_key", "key_mask", "stft_2048_hop512")}
        dom = max(single, key=single.get) if single else (max(stages, key=stages.get) if stages else None)
        roof = None
        if dom:
            k_ms = stages[dom] / args.steps  # per step (all waves of the step)
            launches_per_step = waves[0]     # one launch of the kernel per wave
            ach = nt * ALGO_BYTES_PER_TRACK / (k_ms / 1000.0) / 1e9  # = per-launch bytes / per-launch time
            traffic = None
            tp = ROOT / "profiles" / "roofline_traffic.json"
            if tp.exists():  # DRAM bytes per track of this kernel from the committed `ncu --set full` capture -> bytes per launch
                try:
                    per_track = json.loads(tp.read_text()).get(dom, {}).get("dram_bytes_per_track")
                    traffic = per_track * nt / launches_per_step if per_track else None
                except Exception:
                    traffic = None
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_track": ALGO_BYTES_PER_TRACK, "kernel_ms_per_step": k_ms,
                    "launches_per_step": launches_per_step, "algorithmic_bytes_per_launch": nt * ALGO_BYTES_PER_TRACK / launches_per_step,
                    "kernel_ms_per_launch": k_ms / launches_per_step,
                    "share_of_step": k_ms / ms_step,
                    "note": "path is FP32/shared-memory bound (SURVEY §8d: ~315 flop/B); the HBM fraction is reported as the contract asks",
                    "fp32_tflops_key_stft": (nt * FLOPS_KEY_STFT / (stages["stft_8192_key"] / args.steps / 1000.0) / 1e12) if "stft_8192_key" in stages else None}
            # second denominator: STFT flops of the whole step by the 5 N log2 N convention against the measured FMA rate
            step_flops = nt * (FLOPS_BASE_STFT + FLOPS_KEY_STFT) + n_escalated * FLOPS_ESCALATION
            ach_tf = step_flops / (ms_step / 1000.0) / 1e12
            roof["fp32"] = {"achieved": ach_tf, "peak": fp32_peak or None, "unit": "TFLOP/s", "frac": (ach_tf / fp32_peak) if fp32_peak else None,
                            "peak_source": "measured: FMA microbenchmark of this run (stratum_b200_fp32_peak_tflops), 2 flops per FMA",
                            "flops_per_step": step_flops, "escalated_tracks": n_escalated,
                            "note": "STFT flops only (5 N log2 N for the complex transform of N points; the kernels use the real-input packing, "
                                    "so they execute about half of that); the path compiles with -fmad=false, FMAs only inside the FFT"}
        line = {
            "metric": "tracks_per_sec_3min_44k1", "value": value, "unit": "tracks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: batch of synthetic 3-min 44.1 kHz mono tracks, BPM 70-180, random keys (BASELINE.json configs[1]); analyze_audio defaults",
                       "tracks_per_gpu_per_step": nt, "samples_per_track": N_SAMPLES, "l2": "inputs (32.5 GB per GPU) far larger than L2; no flush needed",
                       "timing": "CUDA events on the library stream around each batch call, max over ranks", "host_placement": numa_note},
            "ms_per_track": ms_step / nt, "wall_ms_per_step": wall_step,
            "gpu_launches": launches, "clocks": clocks,
            "stages_ms_per_step": {k: v / args.steps for k, v in stages.items()},
            "stages_note": "CUDA-event spans per stream; onsets_energy, legacy_bpm, multires_tempogram, final_bpm and beats run on the second stream "
                           "beside other kernels (late split), so their spans overlap the key-path stages and do not add up to the step",
            "results_ok": ok, "bpm_within_2_or_octave": bpm_hit, "tracks_total": total_tracks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": ({"value": e2e["tracks_per_step"] * world / (e_ms_step / 1000.0), "unit": "tracks/s", "h2d_bytes_per_step": int(e2e["h2d"]),
                     "d2h_bytes_per_step": int(e2e["d2h"]), "tracks_per_step": e2e["tracks_per_step"], "ms_per_step": e_ms_step,
                     "pcm16_value": e2e["tracks_per_step"] * world / (p_ms_step / 1000.0), "pinned_h2d_gbs_rank0": e2e["h2d_gbs"],
                     "note": "stratum_b200_analyze_batch on pinned host f32 samples: H2D of the samples and D2H of the results inside the timed region; "
                             "pcm16_value = same through stratum_b200_analyze_batch_pcm16 (int16 upload, conversion on the device)"}
                    if e2e else None),
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
