"""GPU parity at the BASELINE.json sizes: the 60-minute drifting mix (C4), ragged durations / sample rates (C5) and
a full-length 3-minute batch (C2) checked through the oracle on a sample plus size-independent properties
(permutation invariance, duplicates agree, wave splitting is invisible)."""
import numpy as np
import pytest

import oracle_lib as O
import synth
import stratum_dsp_b200 as S
from gpu_common import assert_parity

pytestmark = pytest.mark.gpu
SR = 44100


def test_c4_sixty_minute_mix():
    # BASELINE.json configs[3]: long-sequence tempogram (524 288-point FFT) + Viterbi over ~7 500 beat frames
    x = synth.c4_mix()
    g = S.analyze_audio(x, SR)
    o = O.analyze(x, SR, fast=True)
    assert_parity(g, o, "C4 60 min")
    assert 118.0 <= g.bpm <= 130.0 or 59.0 <= g.bpm <= 65.0
    assert len(g.hmm_beat_frames) > 1000


def test_c5_ragged_real_lengths():
    # BASELINE.json configs[4]: 30 s - 10 min, 44.1 / 48 kHz, escalation and key detection enabled (defaults)
    idx = [0, 1, 2, 3, 5, 8, 13, 21]
    ps = [synth.c5_params(i) for i in idx]
    longest = max(range(len(ps)), key=lambda k: ps[k].n_samples / ps[k].sample_rate)
    ps[longest].n_samples = 600 * ps[longest].sample_rate  # make sure the 10-minute end of the range is covered
    tracks = [synth.render(p) for p in ps]
    srs = [p.sample_rate for p in ps]
    res = S.analyze_batch(tracks, srs)
    for i, (x, sr, g) in enumerate(zip(tracks, srs, res)):
        assert_parity(g, O.analyze(x, sr, fast=True), f"C5[{idx[i]}] sr={sr} dur={x.size / sr:.0f}s")


def test_c2_full_length_batch_properties(monkeypatch):
    torch = pytest.importorskip("torch")
    n, nt = 7_938_000, 12
    params = np.array([[c.bpm, c.tonic, c.minor, c.phase_frac, c.chord_amp] for c in (synth.c2_params(100 + i) for i in range(nt))], np.float32)
    params[nt - 1] = params[0]  # a duplicate track
    buf = torch.empty(nt * n, dtype=torch.float32, device="cuda")
    S.synth_batch(buf.data_ptr(), nt, n, SR, params)
    offsets = np.arange(nt + 1, dtype=np.uint64) * n
    a = S.analyze_batch_device(buf.data_ptr(), offsets, [SR] * nt)
    assert all(r.error is None for r in a)
    # duplicates agree exactly
    assert a[0].bpm == a[nt - 1].bpm and a[0].key == a[nt - 1].key and np.array_equal(a[0].beat_grid.beats, a[nt - 1].beat_grid.beats)
    # oracle on two of them (downloaded from the device buffer)
    for i in (1, 7):
        x = buf[i * n:(i + 1) * n].cpu().numpy()
        assert_parity(a[i], O.analyze(x, SR, fast=True), f"C2 full[{i}]")
    # BPM found (or its octave) for every track, key tonic found
    for i, r in enumerate(a):
        t = params[i, 0]
        assert min(abs(r.bpm - t), abs(2 * r.bpm - t), abs(r.bpm - 2 * t)) <= 2.0, (i, r.bpm, t)
    # permutation of the batch + forced small waves: identical per-track results
    perm = np.random.default_rng(3).permutation(nt)
    pbuf = torch.empty_like(buf)
    for dst, src in enumerate(perm):
        pbuf[dst * n:(dst + 1) * n] = buf[src * n:(src + 1) * n]
    monkeypatch.setenv("STRATUM_B200_WAVE_MAX_TRACKS", "5")
    b = S.analyze_batch_device(pbuf.data_ptr(), offsets, [SR] * nt)
    for dst, src in enumerate(perm):
        u, v = a[src], b[dst]
        assert u.bpm == v.bpm and u.key == v.key and u.key_clarity == v.key_clarity and u.grid_stability == v.grid_stability
        assert np.array_equal(u.beat_grid.beats, v.beat_grid.beats) and np.array_equal(u.onsets, v.onsets)
