"""Shared helpers of the GPU parity tests: run the CUDA path through the C ABI and the CPU oracle on the
same samples and compare under the north-star contract (BASELINE.json): bit-exact integer views
(trim range, onsets, HMM beat-frame indices, beat counts, key label), 1e-3 relative on floats."""
import numpy as np

import oracle_lib as O
import stratum_dsp_b200 as S

REL_TOL = 1e-3  # BASELINE.json north_star: "BPM, confidence and grid stability within 1e-3 relative"


def close(a, b, tol=REL_TOL):
    return abs(a - b) <= tol * max(abs(a), abs(b)) + 1e-6


def assert_parity(g: "S.AnalysisResult", o: "O.OracleResult", label=""):
    assert g.error is None and o.status == 0, (label, g.error, o.error)
    # integer / label views: exact
    assert (g.trim_start, g.trim_end) == (int(o.trim_start), int(o.trim_end)), label
    assert np.array_equal(g.onsets, o.onsets), label
    assert np.array_equal(g.hmm_beat_frames, o.hmm_beat_frames), label
    assert g.key.id == o.key, (label, g.key.id, o.key)
    assert g.time_sig_beats_per_bar == int(o.time_sig_beats_per_bar), label
    assert g.beats_refined == int(o.beats_refined), label
    assert len(g.beat_grid.beats) == len(o.beats) and len(g.beat_grid.downbeats) == len(o.downbeats), label
    opt = lambda v: None if v < 0 else bool(v)
    assert g.metadata.tempogram_multi_res_triggered == opt(o.multi_res_triggered), label
    assert g.metadata.tempogram_multi_res_used == opt(o.multi_res_used), label
    assert g.warnings_mask == int(o.warnings) and g.flags_mask == int(o.flags), label
    # floats: 1e-3 relative
    for name, a, b in [("bpm", g.bpm, o.bpm), ("bpm_confidence", g.bpm_confidence, o.bpm_confidence), ("key_confidence", g.key_confidence, o.key_confidence),
                       ("key_clarity", g.key_clarity, o.key_clarity), ("grid_stability", g.grid_stability, o.grid_stability),
                       ("duration", g.metadata.duration_seconds, o.duration_seconds)]:
        assert close(a, b), (label, name, a, b)
    if len(o.beats):
        assert np.allclose(g.beat_grid.beats, o.beats, rtol=0, atol=1e-4), label
        assert np.allclose(g.beat_grid.downbeats, o.downbeats, rtol=0, atol=1e-4), label
        assert np.array_equal(g.beat_grid.bars, g.beat_grid.downbeats), label
    c = S.compute_confidence(g)
    for k, v in (("bpm", c.bpm_confidence), ("key", c.key_confidence), ("grid", c.grid_stability), ("overall", c.overall_confidence)):
        assert close(v, o.confidence[k]), (label, "confidence." + k, v, o.confidence[k])
