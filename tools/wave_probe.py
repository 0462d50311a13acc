"""GPU-box probe: device-resident throughput of a 1024-track batch against the wave size (STRATUM_B200_WAVE_MAX_TRACKS)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import stratum_dsp_b200 as S  # noqa: E402
from bench import N_SAMPLES, SR, track_params  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
buf = torch.empty(nt * N_SAMPLES, dtype=torch.float32, device="cuda")
S.synth_batch(buf.data_ptr(), nt, N_SAMPLES, SR, track_params(0, nt), 0)
offsets = np.arange(nt + 1, dtype=np.uint64) * np.uint64(N_SAMPLES)
srs = np.full(nt, SR, np.uint32)


def dev_step():
    res = S.analyze_batch_device(buf.data_ptr(), offsets, srs, None, 0, convert=False)
    ms = S.last_call_device_ms()
    S.free_results(res)
    return ms


for wm in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "256,192,160,128,112,96,80,64".split(","))]:
    os.environ["STRATUM_B200_WAVE_MAX_TRACKS"] = str(wm)
    dev_step()
    ms = min(dev_step() for _ in range(2))
    print(f"device-resident {nt} tracks, waves of {wm}: {ms:.1f} ms  ({nt / ms * 1000:.0f} tracks/s)", flush=True)
