"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs."""
import numpy as np
import pytest

import oracle_lib as O
import synth
import stratum_dsp_b200 as S
from gpu_common import assert_parity, close

pytestmark = pytest.mark.gpu
SR = 44100


# 2048 and 8192 points: the register-fused kernels; every other power of two: the generic shared-memory form of the same DAG
@pytest.mark.parametrize("frame,hop", [(2048, 512), (2048, 256), (2048, 1024), (8192, 512), (1024, 256), (4096, 512), (512, 100), (256, 256), (64, 32), (16384, 4096)])
def test_stft_bit_exact(frame, hop):
    rng = np.random.default_rng(frame + hop)
    x = (rng.standard_normal(3 * SR) * 0.25).astype(np.float32)
    g, o = S.stft(x, frame, hop), O.stft(x, frame, hop)
    assert g.shape == o.shape
    assert np.array_equal(g, o)  # same FFT DAG (oracle/so_fft.cpp) => identical bits


def test_stft_short_and_gain():
    x = np.ones(1000, np.float32)
    assert S.stft(x, 2048, 512).shape == (0, 1025)
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(8192) * 0.1).astype(np.float32)
    assert np.array_equal(S.stft(y, 2048, 512, gain=0.5), O.stft((y * np.float32(0.5)).astype(np.float32), 2048, 512))


@pytest.mark.parametrize("seconds", [12.0, 30.0])
def test_c1_click_chord(seconds):
    x = synth.render(synth.c1_params(int(seconds * SR), SR))
    assert_parity(S.analyze_audio(x, SR), O.analyze(x, SR, fast=True), f"C1 {seconds}s")


@pytest.mark.parametrize("i", range(12))
def test_c2_tracks_30s(i):
    p = synth.c2_params(i, 30 * SR, SR)
    x = synth.render(p)
    assert_parity(S.analyze_audio(x, SR), O.analyze(x, SR, fast=True), f"C2[{i}] bpm={p.bpm}")


def test_c1_full_three_minutes():
    # BASELINE.json configs[0]: one 3-min 44.1 kHz 128 BPM click+chord track
    x = synth.render(synth.c1_params())
    g, o = S.analyze_audio(x, SR), O.analyze(x, SR, fast=True)
    assert_parity(g, o, "C1 3min")
    assert abs(g.bpm - 128.0) <= 1.0 and g.key.id == 0


def test_reference_fixtures():
    for name, x in [("120bpm", synth.fixture_kick(120.0, 8.0)), ("128bpm", synth.fixture_kick(128.0, 7.5)), ("cmajor", synth.fixture_cmajor_scale()),
                    ("mixed_silence", synth.fixture_mixed_silence())]:
        g = S.analyze_audio(x, SR)
        assert_parity(g, O.analyze(x, SR), name)
    # tests/integration_tests.rs assertions on the CUDA path itself
    g = S.analyze_audio(synth.fixture_kick(120.0, 8.0), SR)
    assert g.bpm == 0 or abs(g.bpm - 120) <= 2
    g = S.analyze_audio(synth.fixture_mixed_silence(), SR)
    assert 4.0 <= g.metadata.duration_seconds <= 6.0


def test_errors_match_reference():
    with pytest.raises(S.AnalysisError) as e:
        S.analyze_audio(np.zeros(SR, np.float32), SR)
    assert e.value.kind == "ProcessingError" and "silent" in e.value.message  # tests/integration_tests.rs:264-274
    with pytest.raises(S.AnalysisError) as e:
        S.analyze_audio(np.zeros(0, np.float32), SR)
    assert e.value.kind == "InvalidInput" and "Empty" in e.value.message  # lib.rs:100-104
    with pytest.raises(S.AnalysisError) as e:
        S.analyze_audio(np.ones(100, np.float32), 0)
    assert e.value.kind == "InvalidInput"
    with pytest.raises(S.AnalysisError) as e:
        S.analyze_audio(np.ones(100000, np.float32), SR, S.AnalysisConfig(frame_size=1024))
    assert e.value.kind == "NotImplemented"  # unsupported switches are rejected, never silently ignored


@pytest.mark.parametrize("method", [S.NORM_RMS, S.NORM_LOUDNESS])
def test_rms_and_lufs_normalisation(method):
    # a5: normalize_rms (normalization.rs:325-398) and normalize_lufs (:401-484, K-weighting biquad evaluated as a
    # block-parallel affine scan).  Quiet and loud renderings so that both the gain and the clip limiter are hit.
    for i, scale in enumerate((0.05, 0.6, 1.4)):
        p = synth.c2_params(40 + i, 25 * SR, SR)
        x = (synth.render(p) * np.float32(scale)).astype(np.float32)
        g = S.analyze_audio(x, SR, S.AnalysisConfig(normalization=method))
        o = O.analyze(x, SR, {"normalization": method}, fast=True)
        assert_parity(g, o, f"norm={method} scale={scale}")
    # 48 kHz, ragged length (last LUFS block partial)
    p = synth.c2_params(44, 11 * 48000 + 1234, 48000)
    p.sample_rate = 48000
    x = synth.render(p)
    assert_parity(S.analyze_audio(x, 48000, S.AnalysisConfig(normalization=method)), O.analyze(x, 48000, {"normalization": method}), "48k")


def test_lufs_all_gated_falls_back_to_peak():
    x = (synth.render(synth.c2_params(45, 12 * SR, SR)) * np.float32(1e-5)).astype(np.float32)  # every 400 ms block below -70 LUFS
    cfg = S.AnalysisConfig(normalization=S.NORM_LOUDNESS, enable_silence_trimming=False)
    o = O.analyze(x, SR, {"normalization": 2, "enable_silence_trimming": 0})
    assert_parity(S.analyze_audio(x, SR, cfg), o, "gated")


def test_key_segment_fallback_to_whole_track():
    # lib.rs:1385-1411: when no segment reaches key_segment_min_clarity the whole-track detection is used
    x = synth.render(synth.c2_params(90, 30 * SR, SR))
    g = S.analyze_audio(x, SR, S.AnalysisConfig(key_segment_min_clarity=0.99))
    o = O.analyze(x, SR, {"key_segment_min_clarity": 0.99}, fast=True)
    assert_parity(g, o, "all segments rejected")
    batch = S.analyze_batch([x, synth.render(synth.c2_params(91, 30 * SR, SR))], SR, S.AnalysisConfig(key_segment_min_clarity=0.99))
    assert batch[0].key == g.key and batch[0].key_clarity == g.key_clarity
    # voting disabled: whole-track detection directly
    g2 = S.analyze_audio(x, SR, S.AnalysisConfig(enable_key_segment_voting=False))
    assert_parity(g2, O.analyze(x, SR, {"enable_key_segment_voting": 0}, fast=True), "voting off")


def test_short_inputs():
    rng = np.random.default_rng(5)
    for n in (100, 2047, 2048, 4096, 8191, 8192, 12000):
        x = (rng.standard_normal(n) * 0.2).astype(np.float32)
        o = O.analyze(x, SR)
        try:
            g = S.analyze_audio(x, SR)
        except S.AnalysisError as e:
            assert o.status == e.code, (n, e, o.error)
            continue
        assert_parity(g, o, f"noise n={n}")


def test_ragged_batch_matches_single_and_oracle():
    # C5 shape: mixed durations and sample rates in one call; one bad track must not abort the batch
    rng = np.random.default_rng(9)
    tracks, srs = [], []
    for i in range(6):
        p = synth.c5_params(i)
        p.n_samples = min(p.n_samples, 40 * p.sample_rate)
        tracks.append(synth.render(p))
        srs.append(p.sample_rate)
    tracks.insert(3, np.zeros(30000, np.float32))  # silent -> per-track error
    srs.insert(3, 44100)
    tracks.append((rng.standard_normal(5000) * 0.1).astype(np.float32))
    srs.append(48000)
    res = S.analyze_batch(tracks, srs)
    assert len(res) == len(tracks)
    for i, (x, sr, g) in enumerate(zip(tracks, srs, res)):
        o = O.analyze(x, sr, fast=True)
        if o.status:
            assert g.error is not None and g.error.code == o.status, i
            continue
        assert_parity(g, o, f"ragged[{i}] sr={sr} n={x.size}")
        single = S.analyze_audio(x, sr)
        assert single.bpm == g.bpm and single.key == g.key and np.array_equal(single.beat_grid.beats, g.beat_grid.beats)


def test_waves_do_not_change_results(monkeypatch):
    xs = [synth.render(synth.c2_params(20 + i, 15 * SR, SR)) for i in range(5)]
    a = S.analyze_batch(xs, SR)
    monkeypatch.setenv("STRATUM_B200_WAVE_MAX_TRACKS", "2")
    b = S.analyze_batch(xs, SR)
    for u, v in zip(a, b):
        assert u.bpm == v.bpm and u.key == v.key and u.key_clarity == v.key_clarity and np.array_equal(u.beat_grid.beats, v.beat_grid.beats)


def test_escalation_path_is_exercised():
    # BPM 70-80 / 170-180 fall in the trap zones of lib.rs:412-413 -> multi-resolution pass
    hit = 0
    for bpm in (72.0, 76.0, 174.0, 178.0):
        p = synth.TrackParams(bpm, 2, 0, 0.25, 0.1, SR, 30 * SR)
        x = synth.render(p)
        g, o = S.analyze_audio(x, SR), O.analyze(x, SR, fast=True)
        assert_parity(g, o, f"trap bpm={bpm}")
        hit += bool(g.metadata.tempogram_multi_res_triggered)
    assert hit >= 1


def test_stage_intermediates_single_track():
    x = synth.render(synth.c2_params(1, 20 * SR, SR))
    o = O.analyze(x, SR, dump=True)
    S.debug_enable(True)
    try:
        S.analyze_audio(x, SR)
        D = S.debug_array
        assert np.array_equal(D("onset.spectral_flux"), o.farray("onset.spectral_flux"))  # FFT + sqrt + div only: exact
        for nm in ("onset.energy", "onset.spectral", "onset.hfc"):
            assert np.array_equal(D(nm).astype(np.int64), o.iarray(nm)), nm
        for v in ("full", "low", "mid", "high", "mel"):  # logf differs by an ulp between libm and CUDA
            assert np.abs(D(f"base.nov.{v}") - o.farray(f"base.nov.{v}")).max() < 2e-6, v
        assert np.abs(D("key.hpcp_raw") - o.farray("key.hpcp_raw")).max() < 5e-6
        assert np.array_equal(D("hmm.path").astype(np.int64), o.iarray("hmm.path"))  # Viterbi path incl. underflow behaviour
    finally:
        S.debug_enable(False)


def test_time_segmented_mask_is_bit_identical():
    # A wave with few long tracks cuts the harmonic mask's time axis into segments that start from exact prefixes (k_key.cu: mask_kernel
    # seg_len, mask_prefix_kernel); a wave with many tracks walks every bin in one piece.  Same track, both ways: every chroma value, frame
    # energy and key weight must be bit-identical (reference: extractor.rs:1246-1349, a sequential f32 prefix per bin).
    x = synth.render(synth.c2_params(3, 100 * SR, SR))  # 8 598 key frames: segmented when analysed alone
    fill = [synth.render(synth.c2_params(40 + i, 4 * SR, SR)) for i in range(11)]
    S.debug_enable(True)
    try:
        names = ("key.hpcp_raw", "key.energy", "key.weights", "key.band_head")
        S.analyze_audio(x, SR)
        alone = {nm: S.debug_array(nm).copy() for nm in names}
        S.analyze_batch([x] + fill, SR)
        batch = {nm: S.debug_array(nm).copy() for nm in names}
    finally:
        S.debug_enable(False)
    assert alone["key.hpcp_raw"].size == batch["key.hpcp_raw"].size > 8000 * 12  # the dumped track of the batch is the long one
    for nm in names:
        assert np.array_equal(alone[nm], batch[nm]), nm


def test_device_resident_batch_and_synth():
    torch = pytest.importorskip("torch")
    n, nt = 20 * SR, 4
    params = np.array([[c.bpm, c.tonic, c.minor, c.phase_frac, c.chord_amp] for c in (synth.c2_params(i) for i in range(nt))], np.float32)
    buf = torch.empty(nt * n, dtype=torch.float32, device="cuda")
    S.synth_batch(buf.data_ptr(), nt, n, SR, params)
    torch.cuda.synchronize()
    host = buf.cpu().numpy().reshape(nt, n)
    for i in range(nt):  # device generator == numpy generator up to one f32 rounding of the last bit
        p = synth.c2_params(i, n, SR)
        p.phase_frac, p.chord_amp = float(params[i, 3]), float(params[i, 4])
        assert np.abs(host[i] - synth.render(p)).max() < 1e-5
    offsets = np.arange(nt + 1, dtype=np.uint64) * n
    res = S.analyze_batch_device(buf.data_ptr(), offsets, [SR] * nt)
    for i in range(nt):
        assert_parity(res[i], O.analyze(host[i], SR, fast=True), f"device[{i}]")
    assert S.launch_count() > 0


def test_multi_device_sharding_matches_single_device():
    # stratum_b200_analyze_batch(device_ids=...) shards contiguous track ranges over devices, one host thread each
    if S.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    xs = [synth.render(synth.c2_params(60 + i, 15 * SR, SR)) for i in range(7)]
    one = S.analyze_batch(xs, SR, devices=[0])
    two = S.analyze_batch(xs, SR, devices=[0, 1])
    for u, v in zip(one, two):
        assert u.bpm == v.bpm and u.key == v.key and u.key_clarity == v.key_clarity and np.array_equal(u.beat_grid.beats, v.beat_grid.beats)


def _to_pcm16(x):
    return np.clip(np.round(x.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def test_pcm16_ingestion_matches_reference_decoder_arithmetic():
    # int16 PCM uploaded as is, converted + mixed down on the device like examples/analyze_batch.rs:96-113
    mono = _to_pcm16(synth.render(synth.c2_params(70, 14 * SR, SR)))
    left = _to_pcm16(synth.render(synth.c2_params(71, 12 * SR, SR)) * 0.7)
    right = _to_pcm16(synth.render(synth.c2_params(72, 12 * SR, SR)) * 0.7)
    stereo = np.stack([left, right], axis=1)
    res = S.analyze_batch_pcm16([mono, stereo, np.zeros(20000, np.int16)], [SR, 48000, SR])
    f_mono = mono.astype(np.float32) / np.float32(32768.0)
    f_st = ((np.float32(0.0) + left.astype(np.float32) / np.float32(32768.0)) + right.astype(np.float32) / np.float32(32768.0)) / np.float32(2.0)
    assert_parity(res[0], O.analyze(f_mono, SR), "pcm16 mono")
    assert_parity(res[1], O.analyze(f_st.astype(np.float32), 48000), "pcm16 stereo")
    assert res[2].error is not None and "silent" in res[2].error.message
    same = S.analyze_audio(f_mono, SR)
    assert same.bpm == res[0].bpm and np.array_equal(same.onsets, res[0].onsets) and same.key_clarity == res[0].key_clarity


def test_analyze_batch_cli_jsonl(tmp_path):
    import json
    import subprocess
    import sys
    import wave

    paths = []
    for i in range(3):
        x = _to_pcm16(synth.render(synth.c2_params(80 + i, 10 * SR, SR)))
        p = tmp_path / f"t{i}.wav"
        with wave.open(str(p), "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(SR)
            w.writeframes(x.tobytes())
        paths.append(str(p))
    paths.append(str(tmp_path / "missing.wav"))
    root = __import__("pathlib").Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "examples" / "analyze_batch.py"), "--json", *paths], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [json.loads(l) for l in out.stdout.strip().splitlines()]
    assert len(lines) == 4
    for l in lines[:3]:  # examples/analyze_batch.rs:331-343 keys
        assert set(l) == {"file", "bpm", "bpm_confidence", "key", "key_confidence", "processing_time_ms", "tempogram_multi_res_triggered",
                          "tempogram_multi_res_used", "tempogram_percussive_triggered", "tempogram_percussive_used"}
        assert l["bpm"] > 0
    assert set(lines[3]) == {"file", "error"}
    assert "Done: ok=3/4" in out.stderr


def test_hpss_onsets_fourth_detector():
    # enable_hpss_onsets (lib.rs:222-235) — iterative median filtering on the device, 4-way consensus
    x = synth.render(synth.c2_params(95, 7 * SR, SR))
    cfg = {"enable_hpss_onsets": 1}
    o = O.analyze(x, SR, cfg, dump=True, fast=True)
    S.debug_enable(True)
    try:
        g = S.analyze_audio(x, SR, S.AnalysisConfig(**cfg))
        assert np.array_equal(S.debug_array("onset.hpss").astype(np.int64), o.iarray("onset.hpss"))
        ph = o.farray("hpss.perc_head")
        assert np.array_equal(S.debug_array("hpss.perc_head")[: ph.size], ph)  # medians, divisions, products: exact
    finally:
        S.debug_enable(False)
    assert_parity(g, o, "hpss onsets")
    batch = S.analyze_batch([x, synth.render(synth.c2_params(96, 6 * SR, SR))], SR, S.AnalysisConfig(**cfg))
    assert np.array_equal(batch[0].onsets, g.onsets) and batch[0].bpm == g.bpm


def test_percussive_tempogram_fallback():
    # lib.rs:587-683: HPSS + tempogram on the percussive component for low-tempo-trap tracks
    cfg = {"enable_tempogram_percussive_fallback": 1}
    for bpm in (74.0, 128.0):
        x = synth.render(synth.TrackParams(bpm, 2, 0, 0.25, 0.1, SR, 8 * SR))
        o = O.analyze(x, SR, cfg, fast=True)
        g = S.analyze_audio(x, SR, S.AnalysisConfig(**cfg))
        assert_parity(g, o, f"perc fallback bpm={bpm}")
        opt = lambda v: None if v < 0 else bool(v)
        assert g.metadata.tempogram_percussive_triggered == opt(o.percussive_triggered)
        assert g.metadata.tempogram_percussive_used == opt(o.percussive_used)


@pytest.mark.parametrize("cfg", [
    {"enable_key_hpcp": 0},                                 # a33: chroma folding, Gaussian soft mapping (extractor.rs:393-487)
    {"enable_key_hpcp": 0, "soft_chroma_mapping": 0},       # hard nearest-class assignment
    {"enable_key_hpcp": 0, "chroma_sharpening_power": 1.5}, # sharpen_chroma (chroma/normalization.rs:41-65)
    {"enable_key_harmonic_mask": 0},                        # time smoothing alone (lib.rs:1043-1060)
    {"enable_key_harmonic_mask": 0, "enable_key_spectrogram_time_smoothing": 0},  # raw key spectrogram
    {"chroma_sharpening_power": 2.0},
])
def test_key_path_variants(cfg):
    for i, sr in ((97, SR), (98, 48000)):
        p = synth.c2_params(i, 22 * sr, sr)
        p.sample_rate = sr
        x = synth.render(p)
        assert_parity(S.analyze_audio(x, sr, S.AnalysisConfig(**cfg)), O.analyze(x, sr, cfg, fast=True), f"{cfg} sr={sr}")


@pytest.mark.parametrize("cfg", [
    {"key_template_set": 1},                                 # Temperley profiles (key/templates.rs:145-222)
    {"enable_key_edge_trim": 1},                             # lib.rs:1216-1233
    {"enable_key_edge_trim": 1, "key_edge_trim_fraction": 0.3, "enable_key_segment_voting": 0},
    {"enable_key_mode_heuristic": 1},                        # detector.rs:326-518 (flip only reaches the result through whole-track detection)
    {"enable_key_minor_harmonic_bonus": 1},
    {"enable_key_mode_heuristic": 1, "enable_key_minor_harmonic_bonus": 1, "key_mode_third_ratio_margin": 0.1, "key_minor_leading_tone_bonus_weight": 0.5},
    {"enable_key_segment_voting": 0, "enable_key_mode_heuristic": 1, "key_mode_flip_min_score_ratio": 0.3},
    {"enable_key_segment_voting": 0, "enable_key_mode_heuristic": 1, "enable_key_frame_weighting": 0, "key_template_set": 1},
    {"enable_key_ensemble": 1},                              # detector.rs:881-976
    {"enable_key_ensemble": 1, "key_ensemble_kk_weight": 0.8, "key_ensemble_temperley_weight": 0.1, "enable_key_edge_trim": 1},
    {"enable_key_multi_scale": 1},                           # detector.rs:546-700
    {"enable_key_multi_scale": 1, "enable_key_mode_heuristic": 1, "enable_key_minor_harmonic_bonus": 1, "key_multi_scale_weights": [0.5, 1.0, 0.0],
     "key_multi_scale_hop": 45},
    {"enable_key_multi_scale": 1, "key_multi_scale_lengths": [5000, 200], "key_multi_scale_min_clarity": 0.6},
    {"enable_key_multi_scale": 1, "key_multi_scale_lengths": [100000]},   # shorter than every scale: falls through to segment voting (lib.rs:1304-1308)
    {"enable_key_multi_scale": 1, "key_multi_scale_min_clarity": 1.0},   # every window rejected: whole-track fallback (detector.rs:648-668)
    {"enable_key_median": 1},                                # read by nothing in analyze_audio: no effect, as in the reference
])
def test_key_scoring_variants(cfg):
    # template set, edge trim, mode heuristic / minor bonus, ensemble, multi-scale — on material whose modes do not tie
    ocfg = _oracle_cfg(cfg)
    xs = [synth.render_progression(1, 30, SR, tonic=2, minor=True, bpm=124), synth.render_progression(2, 26, 48000, tonic=9, minor=False, bpm=96)]
    srs = [SR, 48000]
    res = [S.analyze_audio(x, sr, S.AnalysisConfig(**cfg)) for x, sr in zip(xs, srs)]
    for i, (x, sr, g) in enumerate(zip(xs, srs, res)):
        assert_parity(g, O.analyze(x, sr, ocfg, fast=True), f"{cfg} track {i}")


def _oracle_cfg(cfg):
    ocfg = {}
    for k, v in cfg.items():
        if isinstance(v, list):
            ocfg["key_multi_scale_n_" + k.rsplit("_", 1)[1]] = len(v)
            for i, e in enumerate(v):
                ocfg[f"{k}[{i}]"] = e
        else:
            ocfg[k] = v
    return ocfg


@pytest.mark.parametrize("cfg", [
    {"enable_key_tuning_compensation": 1},                                           # extractor.rs:66-170; clamps to +-0.08 (lib.rs:1109-1113)
    {"enable_key_tuning_compensation": 1, "key_tuning_max_abs_semitones": 0.5},
    {"enable_key_tuning_compensation": 1, "key_tuning_max_abs_semitones": 0.5, "enable_key_hpcp": 0},   # tuned chroma folding
    {"enable_key_tuning_compensation": 1, "key_tuning_max_abs_semitones": 0.5, "enable_key_hpcp": 0, "soft_chroma_mapping": 0,
     "key_tuning_frame_step": 7, "key_tuning_peak_rel_threshold": 0.6},
    {"enable_key_hpcp_whitening": 1},                                                # extractor.rs:558-580
    {"enable_key_hpcp_whitening": 1, "key_hpcp_whitening_smooth_bins": 9, "key_hpcp_peaks_per_frame": 12},
    {"enable_key_hpcp_whitening": 1, "key_hpcp_whitening_smooth_bins": 2},           # < 3: whitening stays off (extractor.rs:562)
    {"enable_key_hpcp_bass_blend": 1},                                               # extractor.rs:1154-1239
    {"enable_key_hpcp_bass_blend": 1, "enable_key_hpcp_whitening": 1, "key_hpcp_bass_weight": 0.6, "key_hpcp_bass_fmin_hz": 40.0,
     "key_hpcp_bass_fmax_hz": 400.0, "enable_key_tuning_compensation": 1},
    {"enable_key_log_frequency": 1},                                                 # extractor.rs:701-807, 941-984
    {"enable_key_log_frequency": 1, "enable_key_tuning_compensation": 1, "enable_key_beat_synchronous": 1, "enable_key_harmonic_mask": 0},
    {"enable_key_beat_synchronous": 1},                                              # extractor.rs:830-922
    {"enable_key_beat_synchronous": 1, "enable_key_tuning_compensation": 1, "key_tuning_max_abs_semitones": 0.5, "soft_chroma_mapping": 0,
     "enable_key_segment_voting": 0},
    {"enable_key_hpss_harmonic": 1},                                                 # extractor.rs:1369-1501
    {"enable_key_hpss_harmonic": 1, "key_hpss_frame_step": 2, "key_hpss_time_margin": 5, "key_hpss_freq_margin": 10, "key_hpss_mask_power": 1.5},
    {"enable_key_hpss_harmonic": 1, "key_hpss_frame_step": 1, "enable_key_hpcp": 0, "enable_key_multi_scale": 1},
])
def test_key_chroma_variants(cfg):
    # tuning estimate, HPCP whitening / bass blend, log-frequency and beat-synchronous chroma, median-HPSS mask
    ocfg = _oracle_cfg(cfg)
    xs = [synth.render_progression(3, 30, SR, tonic=2, minor=True, bpm=124, detune_cents=30),
          synth.render_progression(4, 24, 48000, tonic=7, minor=False, bpm=100, detune_cents=-20),
          synth.render(synth.c2_params(41, 20 * SR, SR))]   # click track: a dense beat grid for the beat-synchronous branch
    srs = [SR, 48000, SR]
    for i, (x, sr) in enumerate(zip(xs, srs)):
        g = S.analyze_audio(x, sr, S.AnalysisConfig(**cfg))
        assert_parity(g, O.analyze(x, sr, ocfg, fast=True), f"{cfg} track {i}")


@pytest.mark.parametrize("cfg", [
    {"enable_key_stft_override": 0},                                      # key path on the shared 2048/512 geometry (lib.rs:984-1009)
    {"key_stft_frame_size": 2048, "key_stft_hop_size": 256},
    {"key_stft_hop_size": 1024},
    {"key_stft_hop_size": 333, "key_segment_len_frames": 600, "key_segment_hop_frames": 200},   # odd hop: unaligned frame starts
    {"enable_key_stft_override": 0, "enable_key_hpcp_whitening": 1, "enable_key_hpcp_bass_blend": 1, "enable_key_tuning_compensation": 1,
     "key_tuning_max_abs_semitones": 0.5},
    {"enable_key_stft_override": 0, "enable_key_beat_synchronous": 1, "enable_key_harmonic_mask": 0},
    {"key_stft_frame_size": 2048, "key_stft_hop_size": 1024, "enable_key_hpss_harmonic": 1, "enable_key_hpcp": 0},
    {"key_stft_hop_size": 2048, "enable_key_log_frequency": 1, "enable_key_multi_scale": 1},
    {"key_stft_frame_size": 4096},                                        # other powers of two: the generic STFT kernel in front of the same key kernels
    {"key_stft_frame_size": 1024, "key_stft_hop_size": 256},
    {"key_stft_frame_size": 4096, "key_stft_hop_size": 1024, "enable_key_harmonic_mask": 0, "enable_key_hpcp": 0},
    {"key_stft_frame_size": 100},                                         # clamped to 256 (lib.rs:986)
])
def test_key_stft_geometry(cfg):
    # key STFT frames of any power of two from 256 to 8192 at any hop, or no override at all: every key kernel takes the geometry from the configuration
    xs = [synth.render_progression(6, 26, SR, tonic=5, minor=True, bpm=118, detune_cents=15), synth.render(synth.c2_params(44, 14 * 48000, 48000))]
    srs = [SR, 48000]
    for i, (x, sr) in enumerate(zip(xs, srs)):
        g = S.analyze_audio(x, sr, S.AnalysisConfig(**cfg))
        assert_parity(g, O.analyze(x, sr, _oracle_cfg(cfg), fast=True), f"{cfg} track {i}")


@pytest.mark.parametrize("cfg", [
    {"hop_size": 256},
    {"hop_size": 1024},
    {"hop_size": 400},                                                    # not a divisor of the frame: unaligned frame starts, generic frame-RMS kernel
    {"hop_size": 384, "enable_hpss_onsets": 1, "enable_tempogram_percussive_fallback": 1},
    {"hop_size": 256, "enable_key_stft_override": 0, "enable_bpm_fusion": 1},  # key path on the shared 2048 / hop_size geometry
])
def test_hop_size_variants(cfg):
    # hop_size other than 512 (config.rs; lib.rs:156-166, 181-190, 310, 391): the base path runs in a slot of its own at that hop — energy
    # flux, STFT, spectral flux / HFC / HPSS onsets, frame -> sample conversion, legacy estimator, base tempogram — while the
    # multi-resolution pass keeps recomputing hops 256 / 512 / 1024 from the samples (lib.rs:493-509).  76 BPM sits in the low trap zone, so
    # the first track escalates.
    xs = [synth.render(synth.TrackParams(76.0, 3, 0, 0.3, 0.1, SR, 20 * SR)), synth.render(synth.c2_params(45, 13 * 48000, 48000))]
    srs = [SR, 48000]
    res = S.analyze_batch(xs, srs, S.AnalysisConfig(**cfg))
    for i, (x, sr, g) in enumerate(zip(xs, srs, res)):
        assert_parity(g, O.analyze(x, sr, _oracle_cfg(cfg), fast=True), f"{cfg} track {i}")


def test_key_variants_in_a_ragged_batch():
    # per-track decisions (beat grid present or not, tuned or untuned lists, window counts) inside one wave
    cfg = {"enable_key_beat_synchronous": 1, "enable_key_tuning_compensation": 1, "key_tuning_max_abs_semitones": 0.5, "enable_key_mode_heuristic": 1,
           "enable_key_edge_trim": 1}
    xs = [synth.render_progression(5, 28, SR, tonic=4, minor=False, bpm=110, detune_cents=25), synth.render(synth.c2_params(42, 16 * SR, SR)),
          (0.2 * np.sin(2 * np.pi * 220.0 * np.arange(6 * SR) / SR)).astype(np.float32),   # no beats: falls back to the HPCP front end
          synth.render(synth.c2_params(43, 9 * SR, SR))]
    res = S.analyze_batch(xs, SR, S.AnalysisConfig(**cfg))
    for i, (x, g) in enumerate(zip(xs, res)):
        assert_parity(g, O.analyze(x, SR, cfg, fast=True), f"ragged key variants track {i}")


@pytest.mark.parametrize("cfg", [
    {"force_legacy_bpm": 1},
    {"enable_bpm_fusion": 1},
    {"enable_tempogram_multi_resolution": 0},
    {"enable_tempogram_band_fusion": 0},
    {"enable_tempogram_mel_novelty": 0},
    {"enable_tempogram_band_fusion": 0, "enable_tempogram_mel_novelty": 0},
    {"tempogram_band_seed_only": 0},                         # bands and mel take part in the scoring (tempogram.rs:464-484, 590-603)
    {"tempogram_band_seed_only": 0, "tempogram_band_w_low": 0.5, "tempogram_band_w_high": 0.0, "tempogram_mel_weight": 0.3, "emit_tempogram_candidates": 1},
    {"enable_onset_consensus": 0},
    {"enable_silence_trimming": 0},
    {"enable_normalization": 0},
    {"enable_legacy_bpm_guardrails": 0, "force_legacy_bpm": 1},
    {"onset_threshold_percentile": 0.9, "onset_consensus_tolerance_ms": 30},
    {"enable_key_frame_weighting": 0},
    {"enable_key_segment_voting": 0, "enable_key_frame_weighting": 0},
    {"key_segment_len_frames": 512, "key_segment_hop_frames": 128, "key_segment_min_clarity": 0.5},
    {"key_spectrogram_smooth_margin": 6, "key_harmonic_mask_power": 1.5},
    {"key_hpcp_peaks_per_frame": 8, "key_hpcp_num_harmonics": 2, "key_hpcp_harmonic_decay": 0.4, "key_hpcp_mag_power": 0.8},
    {"key_hpcp_num_harmonics": 8, "key_hpcp_harmonic_decay": 0.83},  # decay.powi(h - 1) beyond h = 4: __powisf2's square-and-multiply roundings
    {"enable_key_harmonic_mask": 0},  # time smoothing alone through the compact mask path
    {"enable_key_harmonic_mask": 0, "enable_key_spectrogram_time_smoothing": 0},  # HPCP straight from the STFT rows
    {"enable_key_hpcp_bass_blend": 1},  # two peak bands in the compact band
    {"tempogram_superflux_max_filter_bins": 2, "tempogram_mel_max_filter_bins": 1, "tempogram_novelty_local_mean_window": 8,
     "tempogram_novelty_smooth_window": 3},
    {"min_bpm": 60.0, "max_bpm": 180.0, "bpm_resolution": 0.5},
    {"tempogram_band_low_max_hz": 150.0, "tempogram_band_mid_max_hz": 1500.0, "tempogram_band_high_max_hz": 6000.0, "tempogram_mel_n_mels": 24},
    {"min_amplitude_db": -30.0},
    {"key_min_tonalness": 0.1, "key_tonalness_power": 1.5, "key_energy_power": 0.8},
])
def test_accepted_config_switches_match_the_oracle(cfg):
    # every switch the ABI accepts selects a reference branch: check each against the oracle, not just the defaults
    xs = [synth.render(synth.TrackParams(76.0, 4, 1, 0.4, 0.08, SR, 24 * SR)), synth.render(synth.c2_params(99, 18 * SR, SR))]
    quiet = np.concatenate([np.zeros(SR, np.float32), xs[1][: 10 * SR] * np.float32(0.5), np.zeros(2 * SR, np.float32)])
    res = S.analyze_batch(xs + [quiet], SR, S.AnalysisConfig(**cfg))
    for i, (x, g) in enumerate(zip(xs + [quiet], res)):
        assert_parity(g, O.analyze(x, SR, cfg, fast=True), f"{cfg} track {i}")


@pytest.mark.parametrize("cfg", [{"emit_tempogram_candidates": 1}, {"emit_tempogram_candidates": 1, "enable_tempogram_multi_resolution": 0},
                                 {"emit_tempogram_candidates": 1, "tempogram_candidates_top_n": 40, "tempogram_multi_res_top_k": 12}])
def test_tempogram_candidates_metadata(cfg):
    # metadata.tempogram_candidates (analysis/result.rs:170-181, lib.rs:684-697, 740-752)
    for bpm in (74.0, 121.0):
        x = synth.render(synth.TrackParams(bpm, 2, 0, 0.25, 0.1, SR, 20 * SR))
        o = O.analyze(x, SR, cfg, fast=True)
        g = S.analyze_audio(x, SR, S.AnalysisConfig(**cfg))
        assert_parity(g, o, f"{cfg} bpm={bpm}")
        exp = o.farray("result.candidates").reshape(-1, 5)
        got = g.metadata.tempogram_candidates
        assert got is not None and len(got) == len(exp)
        for (b, sc, fn, an, sel), e in zip(got, exp):
            assert b == e[0] and abs(sc - e[1]) <= 1e-3 * max(abs(e[1]), 1e-6) + 1e-6 and bool(e[4]) == sel
            assert abs(fn - e[2]) < 1e-3 and abs(an - e[3]) < 1e-3
    assert S.analyze_audio(x, SR).metadata.tempogram_candidates is None


def test_concurrent_calls_on_one_device_match_serial():
    # the reference's usage is paths.par_iter().map(analyze_audio) (examples/analyze_batch.rs:260-326): many host threads on one
    # device.  Calls serialise on the device context; each must still get its own track's result.
    import threading
    xs = [synth.render(synth.c2_params(200 + i, (8 + 3 * i) * SR, SR)) for i in range(6)]
    serial = [S.analyze_audio(x, SR) for x in xs]
    got = [None] * len(xs)
    errs = []

    def run(i):
        try:
            for _ in range(2):
                got[i] = S.analyze_audio(xs[i], SR)
        except Exception as e:  # noqa: BLE001
            errs.append((i, e))

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(xs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i, (a, b) in enumerate(zip(got, serial)):
        assert a.bpm == b.bpm and a.key == b.key and a.key_confidence == b.key_confidence and a.grid_stability == b.grid_stability, i
        assert np.array_equal(a.onsets, b.onsets) and np.array_equal(a.hmm_beat_frames, b.hmm_beat_frames), i
        assert np.array_equal(a.beat_grid.beats, b.beat_grid.beats), i


def test_repeated_device_id_is_one_shard():
    xs = [synth.render(synth.c2_params(210 + i, 10 * SR, SR)) for i in range(3)]
    a = S.analyze_batch(xs, SR, devices=[0, 0])
    b = S.analyze_batch(xs, SR)
    for x, y in zip(a, b):
        assert x.bpm == y.bpm and x.key == y.key and np.array_equal(x.onsets, y.onsets)


@pytest.mark.parametrize("sr", [22050, 32000, 8000, 96000])
def test_other_sample_rates_default_key_path(sr):
    # the reference accepts any sample rate; below 39.2 kHz the 100..5000 Hz HPCP band spans more than 1024 key-STFT bins
    p = synth.c2_params(220, 14 * sr, sr)
    p.sample_rate = sr
    x = synth.render(p)
    assert_parity(S.analyze_audio(x, sr), O.analyze(x, sr), f"sr={sr}")


def test_unsupported_track_fails_alone():
    # a track the configured key path cannot take (fixed-size band tables) fails with its own NotImplemented error;
    # the rest of the batch is analysed (examples/analyze_batch.rs:293-322: ItemOut.error)
    lo = synth.c2_params(230, 10 * 22050, 22050)
    lo.sample_rate = 22050
    xs = [synth.render(synth.c2_params(231, 10 * SR, SR)), synth.render(lo), synth.render(synth.c2_params(232, 12 * SR, SR))]
    cfg = S.AnalysisConfig(enable_key_hpcp=False)
    res = S.analyze_batch(xs, [SR, 22050, SR], cfg)
    assert res[1].error is not None and res[1].error.kind == "NotImplemented" and "22050" in res[1].error.message
    for i in (0, 2):
        assert res[i].error is None
        assert_parity(res[i], O.analyze(xs[i], SR, {"enable_key_hpcp": 0}, fast=True), f"track {i}")


def test_fast_divisions_match_ieee():
    # csrc/common.cuh: the per-element divisions of the mask / spectral-flux kernels skip the generic division's range check;
    # inside their operand ranges they must return the IEEE quotient bit for bit (2^28 random tuples per form)
    import ctypes as C
    bad = (C.c_uint64 * 3)()
    L = S.lib()
    L.stratum_b200_debug_check_divisions.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
    assert L.stratum_b200_debug_check_divisions(1 << 28, 12345, bad) == 0, S.last_error()
    assert list(bad) == [0, 0, 0], list(bad)


def test_pcm_formats_match_reference_decoder_arithmetic(tmp_path):
    # every sample format of the reference's decoder loop (examples/analyze_batch.rs:70-165), converted and mixed
    # down on the device from the file's own bytes; mixed formats, channel counts and sample rates in one call
    import wavgen
    cases = [("s24", 2, SR, False), ("f32", 1, 48000, True), ("s32", 6, SR, True), ("u8", 1, SR, False), ("f64", 2, 48000, False), ("s16", 3, SR, True)]
    tracks, want = [], []
    for i, (fmt, ch, sr, ext) in enumerate(cases):
        p = synth.c2_params(300 + i, 11 * sr, sr)
        p.sample_rate = sr
        mono = synth.render(p).astype(np.float64) * 0.8
        x = mono if ch == 1 else np.stack([mono * (1.0 - 0.1 * c) for c in range(ch)], axis=1)
        path = tmp_path / f"{fmt}_{ch}.wav"
        path.write_bytes(wavgen.wav_bytes(x, sr, fmt, extensible=ext, extra_chunk=bool(i & 1)))
        t = S.read_wav(path)
        tracks.append(t)
        want.append((wavgen.decode_reference(t.data.tobytes(), fmt, ch), sr, f"{fmt} x{ch}"))
    res = S.analyze_batch_pcm(tracks)
    for g, (x, sr, label) in zip(res, want):
        assert_parity(g, O.analyze(x, sr), label)
    # the CPU-decoded samples through the f32 entry give the same bits as the device-side conversion
    same = S.analyze_audio(want[0][0], want[0][1])
    assert same.bpm == res[0].bpm and np.array_equal(same.onsets, res[0].onsets) and same.key_clarity == res[0].key_clarity
    # a track that is not a whole number of frames is an argument error
    bad = S.PcmTrack(np.zeros(7, np.uint8), S.PCM_S24, 2, SR)
    with pytest.raises(S.AnalysisError):
        S.analyze_batch_pcm([bad])
