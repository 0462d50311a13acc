"""Writes tests/golden/oracle_golden.json: results of the CPU oracle (oracle/, the restatement of the reference) on
small seeded inputs.  The Rust reference cannot run in this environment, so these are regression vectors of the
oracle, not outputs of the reference itself: they freeze the oracle's behaviour (any edit that changes a result
shows up in tests/test_golden.py) and give the GPU tests a checker that does not need the oracle at run time.
Usage: python tools/make_golden.py"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402
import synth  # noqa: E402

SR = 44100


def cases():
    """name -> (samples, sample_rate, oracle config overrides)"""
    out = {}
    out["c1_12s"] = (synth.render(synth.c1_params(12 * SR, SR)), SR, {})
    for i in (0, 1, 2, 3):
        out[f"c2_{i}_20s"] = (synth.render(synth.c2_params(i, 20 * SR, SR)), SR, {})
    p = synth.c2_params(5, 15 * 48000, 48000)
    p.sample_rate = 48000
    out["c2_5_48k_15s"] = (synth.render(p), 48000, {})
    out["trap_74bpm_25s"] = (synth.render(synth.TrackParams(74.0, 2, 0, 0.25, 0.1, SR, 25 * SR)), SR, {})
    out["fixture_120bpm"] = (synth.fixture_kick(120.0, 8.0), SR, {})
    out["fixture_cmajor"] = (synth.fixture_cmajor_scale(), SR, {})
    out["fixture_mixed_silence"] = (synth.fixture_mixed_silence(), SR, {})
    out["c2_6_rms"] = (synth.render(synth.c2_params(6, 14 * SR, SR)) * np.float32(0.3), SR, {"normalization": 1})
    out["c2_7_lufs"] = (synth.render(synth.c2_params(7, 14 * SR, SR)) * np.float32(0.3), SR, {"normalization": 2})
    # optional key-path variants on material whose modes do not tie
    prog = synth.render_progression(11, 16, SR, tonic=2, minor=True, bpm=124, detune_cents=30)
    out["prog_default"] = (prog, SR, {})
    out["prog_mode_heuristic_bonus"] = (prog, SR, {"enable_key_mode_heuristic": 1, "enable_key_minor_harmonic_bonus": 1, "enable_key_segment_voting": 0})
    out["prog_ensemble"] = (prog, SR, {"enable_key_ensemble": 1, "key_ensemble_kk_weight": 0.7, "key_ensemble_temperley_weight": 0.3})
    out["prog_multi_scale_temperley"] = (prog, SR, {"enable_key_multi_scale": 1, "key_template_set": 1, "enable_key_edge_trim": 1})
    out["prog_tuning_whiten_bass"] = (prog, SR, {"enable_key_tuning_compensation": 1, "key_tuning_max_abs_semitones": 0.5, "enable_key_hpcp_whitening": 1,
                                                 "enable_key_hpcp_bass_blend": 1})
    out["prog_log_frequency"] = (prog, SR, {"enable_key_log_frequency": 1})
    out["prog_hpss_mask_fold"] = (prog, SR, {"enable_key_hpss_harmonic": 1, "enable_key_hpcp": 0})
    out["prog_no_override_beat_sync"] = (synth.render(synth.c2_params(9, 14 * SR, SR)), SR, {"enable_key_stft_override": 0, "enable_key_beat_synchronous": 1})
    out["prog_bpm_fusion_band_scoring"] = (prog, SR, {"enable_bpm_fusion": 1, "tempogram_band_seed_only": 0})
    return out


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def summarize(o) -> dict:
    return {
        "status": int(o.status), "bpm": float(o.bpm), "bpm_confidence": float(o.bpm_confidence), "key": int(o.key),
        "key_confidence": float(o.key_confidence), "key_clarity": float(o.key_clarity), "grid_stability": float(o.grid_stability),
        "trim": [int(o.trim_start), int(o.trim_end)], "n_onsets": int(len(o.onsets)), "onsets_sha": digest(o.onsets.astype(np.int64)),
        "hmm_frames_sha": digest(o.hmm_beat_frames.astype(np.int32)), "n_beats": int(len(o.beats)), "n_downbeats": int(len(o.downbeats)),
        "beats_head": [float(b) for b in o.beats[:4]], "time_sig": int(o.time_sig_beats_per_bar), "beats_refined": int(o.beats_refined),
        "multi_res": [int(o.multi_res_triggered), int(o.multi_res_used)], "warnings": int(o.warnings), "flags": int(o.flags),
        "confidence_overall": float(o.confidence["overall"]),
    }


def main():
    gold = {name: summarize(O.analyze(x, sr, cfg or None)) for name, (x, sr, cfg) in cases().items()}
    path = ROOT / "tests" / "golden" / "oracle_golden.json"
    path.parent.mkdir(exist_ok=True)
    path.write_text(json.dumps(gold, indent=1, sort_keys=True) + "\n")
    print(f"wrote {path} ({len(gold)} cases)")


if __name__ == "__main__":
    main()
