#!/usr/bin/env python3
"""analyze_file — the reference's single-file CLI (examples/analyze_file.rs:185-797) on the B200 path.

    python examples/analyze_file.py <audio.wav> [--json] [flags...]

Every flag of the reference's tool is accepted with the same meaning (examples/analyze_file.rs:191-690: the ~130 switches
the validation harness drives, `validation/tools/run_validation.py`), the `--json` output has the same keys and number
formatting (analyze_file.rs:722-773) and the text output the same lines (:775-793), so the harness can point at this
script instead of the Rust binary.  Decoding is the caller's half of the path: this tool reads RIFF/WAVE
PCM (8/16/24/32-bit) with the standard library and applies the reference decoder's sample arithmetic and mono mixdown
(analyze_file.rs:70-165) on the host; `--debug*` flags are accepted, the Rust logger output they enable has no counterpart.
"""
from __future__ import annotations

import re
import sys
import wave
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stratum_dsp_b200 as S  # noqa: E402

# flag -> [(field, value)] for the switches without an argument (analyze_file.rs:191-219, 254-336, 400-420, 445-468, 510-530, 600-625)
BOOL_FLAGS = {
    "--no-preprocess": [("enable_normalization", 0), ("enable_silence_trimming", 0)],
    "--no-normalize": [("enable_normalization", 0)],
    "--no-trim": [("enable_silence_trimming", 0)],
    "--no-onset-consensus": [("enable_onset_consensus", 0)],
    "--force-legacy-bpm": [("force_legacy_bpm", 1)],
    "--bpm-fusion": [("enable_bpm_fusion", 1)],
    "--bpm-candidates": [("emit_tempogram_candidates", 1)],
    "--no-key-harmonic-mask": [("enable_key_harmonic_mask", 0)],
    "--key-hpss": [("enable_key_hpss_harmonic", 1)],
    "--no-key-hpss": [("enable_key_hpss_harmonic", 0)],
    "--no-key-stft-override": [("enable_key_stft_override", 0)],
    "--key-stft-override": [("enable_key_stft_override", 1)],
    "--no-key-log-freq": [("enable_key_log_frequency", 0)],
    "--key-log-freq": [("enable_key_log_frequency", 1)],
    "--no-key-beat-sync": [("enable_key_beat_synchronous", 0)],
    "--key-beat-sync": [("enable_key_beat_synchronous", 1)],
    "--no-key-multi-scale": [("enable_key_multi_scale", 0)],
    "--key-multi-scale": [("enable_key_multi_scale", 1)],
    "--key-template-temperley": [("key_template_set", 1)],
    "--key-template-kk": [("key_template_set", 0)],
    "--no-key-ensemble": [("enable_key_ensemble", 0)],
    "--key-ensemble": [("enable_key_ensemble", 1)],
    "--no-key-median": [("enable_key_median", 0)],
    "--key-median": [("enable_key_median", 1)],
    "--no-key-tuning": [("enable_key_tuning_compensation", 0)],
    "--no-key-edge-trim": [("enable_key_edge_trim", 0)],
    "--no-key-segment-voting": [("enable_key_segment_voting", 0)],
    "--no-key-mode-heuristic": [("enable_key_mode_heuristic", 0)],
    "--key-mode-heuristic": [("enable_key_mode_heuristic", 1)],
    "--key-hpcp": [("enable_key_hpcp", 1)],
    "--key-hpcp-whitening": [("enable_key_hpcp", 1), ("enable_key_hpcp_whitening", 1)],
    "--no-key-minor-harmonic-bonus": [("enable_key_minor_harmonic_bonus", 0)],
    "--key-minor-harmonic-bonus": [("enable_key_minor_harmonic_bonus", 1)],
    "--no-key-hpcp-bass": [("enable_key_hpcp_bass_blend", 0)],
    "--no-key-spec-smooth": [("enable_key_spectrogram_time_smoothing", 0)],
    "--no-key-frame-weighting": [("enable_key_frame_weighting", 0)],
    "--no-tempogram-multi-res": [("enable_tempogram_multi_resolution", 0)],
    "--multi-res-human-prior": [("enable_tempogram_multi_resolution", 1), ("tempogram_multi_res_use_human_prior", 1)],
    "--no-tempogram-percussive": [("enable_tempogram_percussive_fallback", 0)],
    "--no-tempogram-band-fusion": [("enable_tempogram_band_fusion", 0)],
    "--band-score-fusion": [("tempogram_band_seed_only", 0)],
    "--no-tempogram-mel-novelty": [("enable_tempogram_mel_novelty", 0)],
}
MR = [("enable_tempogram_multi_resolution", 1)]
HPSS = [("enable_key_hpss_harmonic", 1)]
OVR = [("enable_key_stft_override", 1)]
HPCP = [("enable_key_hpcp", 1)]
BASS = [("enable_key_hpcp_bass_blend", 1)]
# flag -> (kind, field, transform, implied switches) for the flags with a value
VALUE_FLAGS = {
    "--bpm-candidates-top": ("usize", "tempogram_candidates_top_n", None, [("emit_tempogram_candidates", 1)]),
    "--key-harmonic-mask-power": ("f32", "key_harmonic_mask_power", None, []),
    "--key-hpss-frame-step": ("usize", "key_hpss_frame_step", lambda n: max(n, 1), HPSS),
    "--key-hpss-time-margin": ("usize", "key_hpss_time_margin", None, HPSS),
    "--key-hpss-freq-margin": ("usize", "key_hpss_freq_margin", None, HPSS),
    "--key-hpss-mask-power": ("f32", "key_hpss_mask_power", None, HPSS),
    "--key-stft-frame-size": ("usize", "key_stft_frame_size", lambda n: max(n, 256), OVR),
    "--key-stft-hop-size": ("usize", "key_stft_hop_size", lambda n: max(n, 1), OVR),
    "--key-multi-scale-lengths": ("usize_list", "key_multi_scale_lengths", None, [("enable_key_multi_scale", 1)]),
    "--key-multi-scale-hop": ("usize", "key_multi_scale_hop", lambda n: max(n, 1), [("enable_key_multi_scale", 1)]),
    "--key-multi-scale-min-clarity": ("f32", "key_multi_scale_min_clarity", lambda x: min(max(x, 0.0), 1.0), [("enable_key_multi_scale", 1)]),
    "--key-multi-scale-weights": ("f32_list", "key_multi_scale_weights", None, [("enable_key_multi_scale", 1)]),
    "--key-ensemble-kk-weight": ("f32", "key_ensemble_kk_weight", lambda x: max(x, 0.0), [("enable_key_ensemble", 1)]),
    "--key-ensemble-temperley-weight": ("f32", "key_ensemble_temperley_weight", lambda x: max(x, 0.0), [("enable_key_ensemble", 1)]),
    "--key-median-segment-length-frames": ("usize", "key_median_segment_length_frames", lambda n: max(n, 120), [("enable_key_median", 1)]),
    "--key-median-segment-hop-frames": ("usize", "key_median_segment_hop_frames", lambda n: max(n, 1), [("enable_key_median", 1)]),
    "--key-median-min-segments": ("usize", "key_median_min_segments", lambda n: max(n, 1), [("enable_key_median", 1)]),
    "--key-tuning-max-semitones": ("f32", "key_tuning_max_abs_semitones", None, []),
    "--key-tuning-frame-step": ("usize", "key_tuning_frame_step", None, []),
    "--key-tuning-peak-rel-threshold": ("f32", "key_tuning_peak_rel_threshold", None, []),
    "--key-edge-trim-fraction": ("f32", "key_edge_trim_fraction", None, []),
    "--key-segment-len-frames": ("usize", "key_segment_len_frames", None, []),
    "--key-segment-hop-frames": ("usize", "key_segment_hop_frames", None, []),
    "--key-segment-min-clarity": ("f32", "key_segment_min_clarity", None, []),
    "--key-mode-third-margin": ("f32", "key_mode_third_ratio_margin", None, [("enable_key_mode_heuristic", 1)]),
    "--key-mode-flip-min-score-ratio": ("f32", "key_mode_flip_min_score_ratio", None, [("enable_key_mode_heuristic", 1)]),
    "--key-hpcp-peaks": ("usize", "key_hpcp_peaks_per_frame", None, HPCP),
    "--key-hpcp-harmonics": ("usize", "key_hpcp_num_harmonics", None, HPCP),
    "--key-hpcp-harmonic-decay": ("f32", "key_hpcp_harmonic_decay", None, HPCP),
    "--key-hpcp-mag-power": ("f32", "key_hpcp_mag_power", None, HPCP),
    "--key-hpcp-whitening-smooth-bins": ("usize", "key_hpcp_whitening_smooth_bins", lambda n: max(n, 3), HPCP + [("enable_key_hpcp_whitening", 1)]),
    "--key-minor-leading-tone-bonus-weight": ("f32", "key_minor_leading_tone_bonus_weight", None, [("enable_key_minor_harmonic_bonus", 1)]),
    "--key-hpcp-bass-fmin-hz": ("f32", "key_hpcp_bass_fmin_hz", None, BASS),
    "--key-hpcp-bass-fmax-hz": ("f32", "key_hpcp_bass_fmax_hz", None, BASS),
    "--key-hpcp-bass-weight": ("f32", "key_hpcp_bass_weight", None, BASS),
    "--key-spec-smooth-margin": ("usize", "key_spectrogram_smooth_margin", None, []),
    "--key-min-tonalness": ("f32", "key_min_tonalness", None, []),
    "--key-tonalness-power": ("f32", "key_tonalness_power", None, []),
    "--key-energy-power": ("f32", "key_energy_power", None, []),
    "--multi-res-top-k": ("usize", "tempogram_multi_res_top_k", None, MR),
    "--multi-res-w512": ("f32", "tempogram_multi_res_w512", None, MR),
    "--multi-res-w256": ("f32", "tempogram_multi_res_w256", None, MR),
    "--multi-res-w1024": ("f32", "tempogram_multi_res_w1024", None, MR),
    "--multi-res-structural-discount": ("f32", "tempogram_multi_res_structural_discount", None, MR),
    "--multi-res-double-time-512-factor": ("f32", "tempogram_multi_res_double_time_512_factor", None, MR),
    "--multi-res-margin-threshold": ("f32", "tempogram_multi_res_margin_threshold", None, MR),
    "--band-low-max-hz": ("f32", "tempogram_band_low_max_hz", None, []),
    "--band-mid-max-hz": ("f32", "tempogram_band_mid_max_hz", None, []),
    "--band-high-max-hz": ("f32", "tempogram_band_high_max_hz", None, []),
    "--band-w-full": ("f32", "tempogram_band_w_full", None, []),
    "--band-w-low": ("f32", "tempogram_band_w_low", None, []),
    "--band-w-mid": ("f32", "tempogram_band_w_mid", None, []),
    "--band-w-high": ("f32", "tempogram_band_w_high", None, []),
    "--superflux-max-filter-bins": ("usize", "tempogram_superflux_max_filter_bins", None, []),
    "--band-support-threshold": ("f32", "tempogram_band_support_threshold", None, []),
    "--band-consensus-bonus": ("f32", "tempogram_band_consensus_bonus", None, []),
    "--mel-n-mels": ("usize", "tempogram_mel_n_mels", None, []),
    "--mel-fmin-hz": ("f32", "tempogram_mel_fmin_hz", None, []),
    "--mel-fmax-hz": ("f32", "tempogram_mel_fmax_hz", None, []),
    "--mel-max-filter-bins": ("usize", "tempogram_mel_max_filter_bins", None, []),
    "--mel-weight": ("f32", "tempogram_mel_weight", None, []),
    "--novelty-w-spectral": ("f32", "tempogram_novelty_w_spectral", None, []),
    "--novelty-w-energy": ("f32", "tempogram_novelty_w_energy", None, []),
    "--novelty-w-hfc": ("f32", "tempogram_novelty_w_hfc", None, []),
    "--novelty-local-mean-window": ("usize", "tempogram_novelty_local_mean_window", None, []),
    "--novelty-smooth-window": ("usize", "tempogram_novelty_smooth_window", None, []),
    "--legacy-preferred-min": ("f32", "legacy_bpm_preferred_min", None, []),
    "--legacy-preferred-max": ("f32", "legacy_bpm_preferred_max", None, []),
    "--legacy-soft-min": ("f32", "legacy_bpm_soft_min", None, []),
    "--legacy-soft-max": ("f32", "legacy_bpm_soft_max", None, []),
    "--legacy-mul-preferred": ("f32", "legacy_bpm_conf_mul_preferred", None, []),
    "--legacy-mul-soft": ("f32", "legacy_bpm_conf_mul_soft", None, []),
    "--legacy-mul-extreme": ("f32", "legacy_bpm_conf_mul_extreme", None, []),
}
# the reference applies its overrides in source order (a later block can undo an earlier one); this is that order
ORDER = ["--no-preprocess", "--no-normalize", "--no-trim", "--no-onset-consensus", "--force-legacy-bpm", "--bpm-fusion", "--bpm-candidates",
         "--bpm-candidates-top", "--no-key-harmonic-mask", "--key-harmonic-mask-power", "--key-hpss", "--no-key-hpss", "--key-hpss-frame-step",
         "--key-hpss-time-margin", "--key-hpss-freq-margin", "--key-hpss-mask-power", "--no-key-stft-override", "--key-stft-override",
         "--key-stft-frame-size", "--key-stft-hop-size", "--no-key-log-freq", "--key-log-freq", "--no-key-beat-sync", "--key-beat-sync",
         "--no-key-multi-scale", "--key-multi-scale", "--key-multi-scale-lengths", "--key-multi-scale-hop", "--key-multi-scale-min-clarity",
         "--key-multi-scale-weights", "--key-template-temperley", "--key-template-kk", "--no-key-ensemble", "--key-ensemble", "--key-ensemble-kk-weight",
         "--key-ensemble-temperley-weight", "--no-key-median", "--key-median", "--key-median-segment-length-frames", "--key-median-segment-hop-frames",
         "--key-median-min-segments", "--no-key-tuning", "--key-tuning-max-semitones", "--key-tuning-frame-step", "--key-tuning-peak-rel-threshold",
         "--no-key-edge-trim", "--key-edge-trim-fraction", "--no-key-segment-voting", "--key-segment-len-frames", "--key-segment-hop-frames",
         "--key-segment-min-clarity", "--no-key-mode-heuristic", "--key-mode-heuristic", "--key-mode-third-margin", "--key-mode-flip-min-score-ratio",
         "--key-hpcp", "--key-hpcp-peaks", "--key-hpcp-harmonics", "--key-hpcp-harmonic-decay", "--key-hpcp-mag-power", "--key-hpcp-whitening",
         "--key-hpcp-whitening-smooth-bins", "--no-key-minor-harmonic-bonus", "--key-minor-harmonic-bonus", "--key-minor-leading-tone-bonus-weight",
         "--no-key-hpcp-bass", "--key-hpcp-bass-fmin-hz", "--key-hpcp-bass-fmax-hz", "--key-hpcp-bass-weight", "--no-key-spec-smooth",
         "--key-spec-smooth-margin", "--no-key-frame-weighting", "--key-min-tonalness", "--key-tonalness-power", "--key-energy-power",
         "--no-tempogram-multi-res", "--multi-res-top-k", "--multi-res-w512", "--multi-res-w256", "--multi-res-w1024", "--multi-res-structural-discount",
         "--multi-res-double-time-512-factor", "--multi-res-margin-threshold", "--multi-res-human-prior", "--no-tempogram-percussive",
         "--no-tempogram-band-fusion", "--band-score-fusion", "--no-tempogram-mel-novelty", "--band-low-max-hz", "--band-mid-max-hz", "--band-high-max-hz",
         "--band-w-full", "--band-w-low", "--band-w-mid", "--band-w-high", "--superflux-max-filter-bins", "--band-support-threshold", "--band-consensus-bonus",
         "--mel-n-mels", "--mel-fmin-hz", "--mel-fmax-hz", "--mel-max-filter-bins", "--mel-weight", "--novelty-w-spectral", "--novelty-w-energy",
         "--novelty-w-hfc", "--novelty-local-mean-window", "--novelty-smooth-window", "--legacy-preferred-min", "--legacy-preferred-max", "--legacy-soft-min",
         "--legacy-soft-max", "--legacy-mul-preferred", "--legacy-mul-soft", "--legacy-mul-extreme"]
IGNORED = {"--json", "--debug"}                       # output / logger switches
IGNORED_WITH_VALUE = {"--debug-track-id", "--debug-gt-bpm"}  # stderr diagnostics of the reference only


def _arg_value(args, name):
    # the reference looks a flag up by position and takes the next argument (analyze_file.rs:221-226); first occurrence wins
    if name in args:
        i = args.index(name)
        if i + 1 < len(args):
            return args[i + 1]
    return None


def _parse(kind, text):
    try:
        if kind == "f32":
            return float(np.float32(float(text)))
        if kind == "usize":  # usize::from_str: digits with an optional leading '+'
            return int(text) if re.fullmatch(r"\+?[0-9]+", text) else None
        if kind == "usize_list":
            parts = [t.strip() for t in text.split(",")]
            return [int(t) for t in parts] if all(re.fullmatch(r"\+?[0-9]+", t) for t in parts) else None
        if kind == "f32_list":
            return [float(np.float32(float(t.strip()))) for t in text.split(",")]
    except ValueError:
        return None
    return None


def build_config(args) -> "S.AnalysisConfig":
    """AnalysisConfig::default() + the overrides of examples/analyze_file.rs:254-690, applied in the reference's order.
    Like the reference, unknown arguments are ignored and a value that does not parse leaves the default in place."""
    cfg = S.AnalysisConfig()
    for flag in ORDER:
        if flag in BOOL_FLAGS:
            if flag in args:
                for field, v in BOOL_FLAGS[flag]:
                    setattr(cfg, field, v)
            continue
        kind, field, tr, implied = VALUE_FLAGS[flag]
        text = _arg_value(args, flag)
        if text is None:
            continue
        v = _parse(kind, text)
        if v is None:
            continue
        for f2, v2 in implied:
            setattr(cfg, f2, v2)
        setattr(cfg, field, tr(v) if tr else v)
    return cfg


def decode_wav(path: str):
    """RIFF/WAVE PCM to mono f32 with the sample arithmetic of analyze_file.rs:70-165 (per-channel conversion, left-to-right
    f32 sum over the channels, division by the channel count)."""
    with wave.open(path, "rb") as w:
        ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    elif width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) / np.float32(8388608.0)
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / np.float32(2147483648.0)
    else:
        raise ValueError("Unsupported audio format")
    if ch > 1:
        x = x.reshape(-1, ch)
        acc = x[:, 0].copy()
        for c in range(1, ch):
            acc = acc + x[:, c]
        x = acc / np.float32(ch)
    return np.ascontiguousarray(x, dtype=np.float32), sr


def render_json(r, c) -> str:
    """The --json document of analyze_file.rs:722-773 (same keys, order and precision)."""
    tf = lambda v: "true" if v else "false"
    out = ["{", f'  "bpm": {r.bpm:.2f},', f'  "bpm_confidence": {c.bpm_confidence:.2f},', f'  "key": "{r.key.name()}",',
           f'  "key_confidence": {c.key_confidence:.2f},', f'  "key_clarity": {r.key_clarity:.2f},', f'  "grid_stability": {r.grid_stability:.2f},']
    m = r.metadata
    for name in ("tempogram_multi_res_triggered", "tempogram_multi_res_used", "tempogram_percussive_triggered", "tempogram_percussive_used"):
        v = getattr(m, name)
        if v is not None:
            out.append(f'  "{name}": {tf(v)},')
    if m.tempogram_candidates is not None:
        out.append('  "bpm_candidates": [')
        for i, (bpm, score, fn, an, sel) in enumerate(m.tempogram_candidates):
            comma = "" if i + 1 == len(m.tempogram_candidates) else ","
            out.append(f'    {{ "bpm": {bpm:.2f}, "score": {score:.4f}, "fft_norm": {fn:.4f}, "autocorr_norm": {an:.4f}, "selected": {tf(sel)} }}{comma}')
        out.append("  ],")
    out.append(f'  "processing_time_ms": {m.processing_time_ms:.2f}')
    out.append("}")
    return "\n".join(out)


def main(argv) -> int:
    if len(argv) < 1 or argv[0] in ("--help", "-h"):
        print(__doc__, file=sys.stderr)
        return 1
    path, args = argv[0], argv[1:]
    try:
        samples, sr = decode_wav(path)
    except Exception as e:
        print(f"Error: {e}", file=sys.stderr)
        return 1
    if samples.size == 0:
        print("ERROR: No audio samples decoded from file", file=sys.stderr)
        return 1
    cfg = build_config(args)
    if "--debug" in args:
        print("=== DEBUG MODE ===")
        print(f"Audio file: {path}")
        print(f"Samples: {samples.size}, Sample rate: {sr} Hz")
        print(f"Duration: {np.float32(samples.size) / np.float32(sr):.2f} seconds")
        print()
    try:
        r = S.analyze_audio(samples, sr, cfg)
    except S.AnalysisError as e:
        # Display of AnalysisError (src/error.rs:24-34)
        names = {"InvalidInput": "Invalid input", "DecodingError": "Decoding error", "ProcessingError": "Processing error", "NotImplemented": "Not implemented",
                 "NumericalError": "Numerical error"}
        print(f"ERROR: Analysis failed: {names.get(e.kind, e.kind)}: {e.message}", file=sys.stderr)
        return 1
    c = S.compute_confidence(r)
    if "--json" in args:
        print(render_json(r, c))
    else:
        print("Analysis Results:")
        print(f"  BPM: {r.bpm:.2f} (confidence: {c.bpm_confidence:.2f})")
        print(f"  Key: {r.key.name()} (confidence: {c.key_confidence:.2f}, clarity: {r.key_clarity:.2f})")
        print(f"  Grid stability: {r.grid_stability:.2f}")
        print(f"  Processing time: {r.metadata.processing_time_ms:.2f} ms")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
