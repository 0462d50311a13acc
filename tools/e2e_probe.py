"""GPU-box probe of the host-buffer (e2e) path: pinned H2D bandwidth, wave-size sensitivity of the device-resident path
and chunk-size sensitivity of stratum_b200_analyze_batch.  Prints one line per measurement."""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import stratum_dsp_b200 as S  # noqa: E402
from bench import N_SAMPLES, SR, track_params  # noqa: E402

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 256
buf = torch.empty(ne * N_SAMPLES, dtype=torch.float32, device="cuda")
S.synth_batch(buf.data_ptr(), ne, N_SAMPLES, SR, track_params(0, ne), 0)
host = torch.empty(ne * N_SAMPLES, dtype=torch.float32, pin_memory=True)
host.copy_(buf)
torch.cuda.synchronize()
# raw pinned H2D bandwidth
dst = torch.empty(64 * N_SAMPLES, dtype=torch.float32, device="cuda")
for _ in range(2):
    t0 = time.perf_counter()
    dst.copy_(host[: 64 * N_SAMPLES], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"h2d pinned 2GB: {64 * N_SAMPLES * 4 / dt / 1e9:.1f} GB/s")
del dst
offsets = np.arange(ne + 1, dtype=np.uint64) * np.uint64(N_SAMPLES)
srs = np.full(ne, SR, np.uint32)
hnp = host.numpy()


def dev_step():
    res = S.analyze_batch_device(buf.data_ptr(), offsets, srs, None, 0, convert=False)
    ms = S.last_call_device_ms()
    S.free_results(res)
    return ms


def host_step():
    res = (S.StratumResult * ne)()
    t0 = time.perf_counter()
    st = S.lib().stratum_b200_analyze_batch(hnp.ctypes.data, offsets.ctypes.data_as(C.POINTER(C.c_uint64)), srs.ctypes.data_as(C.POINTER(C.c_uint32)), ne, None,
                                            (C.c_int32 * 1)(0), 1, res)
    dt = (time.perf_counter() - t0) * 1000
    assert st == 0, S.last_error()
    S.free_results(res)
    return dt


for wm in (256, 128, 64, 32):
    os.environ["STRATUM_B200_WAVE_MAX_TRACKS"] = str(wm)
    dev_step()
    ms = min(dev_step() for _ in range(2))
    print(f"device-resident {ne} tracks, waves of {wm}: {ms:.1f} ms  ({ne / ms * 1000:.0f} tracks/s)")
os.environ.pop("STRATUM_B200_WAVE_MAX_TRACKS")
for extra in sys.argv[2:]:
    k, v = extra.split("=")
    os.environ[k] = v
for mb in (2048, 3876):
    os.environ["STRATUM_B200_STAGE_MB"] = str(mb)
    host_step()
    ms = min(host_step() for _ in range(2))
    print(f"host f32 {ne} tracks, stage {mb} MB: {ms:.1f} ms  ({ne / ms * 1000:.0f} tracks/s)")
    S.stage_timing(True)
    S.stage_times(reset=True)
    ms = host_step()
    st = S.stage_times(reset=True)
    S.stage_timing(False)
    dev = sum(v for k, v in st.items() if not k.startswith("host_"))
    print(f"   one call {ms:.1f} ms: device stages {dev:.1f} ms; host spans:", {k: round(v, 1) for k, v in st.items() if k.startswith("host_")})

# decoder-side entry: int16 PCM uploaded as is, converted on the device
pcm = torch.empty(ne * N_SAMPLES, dtype=torch.int16, pin_memory=True)
for i in range(0, ne, 8):
    a, b = i * N_SAMPLES, min(i + 8, ne) * N_SAMPLES
    pcm[a:b].copy_((buf[a:b] * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
torch.cuda.synchronize()
pnp = pcm.numpy()
chans = np.ones(ne, np.uint32)


def pcm_step():
    res = (S.StratumResult * ne)()
    t0 = time.perf_counter()
    st = S.lib().stratum_b200_analyze_batch_pcm16(pnp.ctypes.data, offsets.ctypes.data_as(C.POINTER(C.c_uint64)), srs.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                  chans.ctypes.data_as(C.POINTER(C.c_uint32)), ne, None, (C.c_int32 * 1)(0), 1, res)
    dt = (time.perf_counter() - t0) * 1000
    assert st == 0, S.last_error()
    S.free_results(res)
    return dt


for mb in (2048, 3876):
    os.environ["STRATUM_B200_STAGE_MB"] = str(mb)
    pcm_step()
    ms = [pcm_step() for _ in range(3)]
    print(f"host pcm16 {ne} tracks, stage {mb} MB: {min(ms):.1f} ms  ({ne / min(ms) * 1000:.0f} tracks/s)  all={[round(x) for x in ms]}")
    S.stage_timing(True)
    S.stage_times(reset=True)
    pcm_step()
    print("   stages:", {k: round(v, 1) for k, v in S.stage_times(reset=True).items() if v > 5})
    S.stage_timing(False)
