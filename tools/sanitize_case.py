"""Small mixed workload for compute-sanitizer runs (memcheck / racecheck): ragged batch, both sample rates, an
escalating track, a silent track, RMS and LUFS normalisation, the PCM16 entry and the raw STFT entry."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import synth  # noqa: E402
import stratum_dsp_b200 as S  # noqa: E402

SR = 44100
tracks, srs = [], []
for i, (bpm, secs, sr) in enumerate([(128.0, 14, 44100), (74.0, 16, 44100), (150.0, 13, 48000)]):
    p = synth.TrackParams(bpm, i * 3, i % 2, 0.3, 0.1, sr, secs * sr)
    tracks.append(synth.render(p))
    srs.append(sr)
tracks.append(np.zeros(30000, np.float32))
srs.append(44100)
tracks.append((np.random.default_rng(0).standard_normal(5000) * 0.1).astype(np.float32))
srs.append(44100)
for cfg in (None, S.AnalysisConfig(normalization=S.NORM_RMS), S.AnalysisConfig(normalization=S.NORM_LOUDNESS)):
    res = S.analyze_batch(tracks, srs, cfg)
    print([None if r.error else round(r.bpm, 2) for r in res], [None if r.error else r.key.name() for r in res])
pcm = [np.clip(np.round(t * 32767), -32768, 32767).astype(np.int16) for t in tracks[:2]]
print([round(r.bpm, 2) for r in S.analyze_batch_pcm16(pcm, srs[:2])])
print(S.stft(tracks[0][: 3 * SR], 8192, 512).shape, S.stft(tracks[0][: 2 * SR], 2048, 256).shape)
# one long track per call: the time-segmented mask (Fk >= 8192 with a handful of tracks) and, at 25 minutes, the consensus vote's
# global-memory work arrays (more than 12 288 onsets)
for secs in (120, 25 * 60):
    p = synth.TrackParams(126.0, 5, 0, 0.2, 0.1, SR, secs * SR)
    r = S.analyze_audio(synth.render(p), SR)
    print(secs, round(r.bpm, 2), r.key.name(), len(r.onsets) if hasattr(r, "onsets") else None)
S.shutdown()
print("done")
