"""Short single-GPU command for ncu captures: N synthetic 3-minute tracks through stratum_b200_analyze_batch_device once."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import stratum_dsp_b200 as S  # noqa: E402
from bench import N_SAMPLES, SR, c2_param_rows as track_params  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
buf = torch.empty(nt * N_SAMPLES, dtype=torch.float32, device="cuda")
S.synth_batch(buf.data_ptr(), nt, N_SAMPLES, SR, track_params(0, nt), 0)
offsets = np.arange(nt + 1, dtype=np.uint64) * np.uint64(N_SAMPLES)
srs = np.full(nt, SR, np.uint32)
for _ in range(reps):
    res = S.analyze_batch_device(buf.data_ptr(), offsets, srs, None, 0, convert=False)
    ok = sum(1 for r in res if r.status == 0)
    S.free_results(res)
print(f"{ok}/{nt} ok, {S.last_call_device_ms():.1f} ms")
