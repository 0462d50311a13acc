"""Golden vectors (tests/golden/oracle_golden.json, written by tools/make_golden.py): the oracle must keep
reproducing them bit for bit (CPU), and the CUDA path must match them under the parity contract (GPU) without the
oracle in the loop."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import make_golden as G  # noqa: E402

GOLD = json.loads((ROOT / "tests" / "golden" / "oracle_golden.json").read_text())
CASES = G.cases()


def test_golden_covers_the_generator():
    assert set(GOLD) == set(CASES) and len(GOLD) >= 10


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden(name):
    import oracle_lib as O

    x, sr, cfg = CASES[name]
    assert G.summarize(O.analyze(x, sr, cfg or None)) == GOLD[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_matches_golden(name):
    import stratum_dsp_b200 as S

    x, sr, cfg = CASES[name]
    g = S.analyze_audio(x, sr, S.AnalysisConfig(**cfg) if cfg else None)
    e = GOLD[name]
    assert [g.trim_start, g.trim_end] == e["trim"]
    assert len(g.onsets) == e["n_onsets"] and G.digest(g.onsets.astype(np.int64)) == e["onsets_sha"]
    assert G.digest(g.hmm_beat_frames.astype(np.int32)) == e["hmm_frames_sha"]
    assert g.key.id == e["key"] and g.time_sig_beats_per_bar == e["time_sig"] and g.beats_refined == e["beats_refined"]
    assert len(g.beat_grid.beats) == e["n_beats"] and len(g.beat_grid.downbeats) == e["n_downbeats"]
    assert g.warnings_mask == e["warnings"] and g.flags_mask == e["flags"]
    opt = lambda v: -1 if v is None else int(v)
    assert [opt(g.metadata.tempogram_multi_res_triggered), opt(g.metadata.tempogram_multi_res_used)] == e["multi_res"]
    for k, v in (("bpm", g.bpm), ("bpm_confidence", g.bpm_confidence), ("key_confidence", g.key_confidence), ("key_clarity", g.key_clarity),
                 ("grid_stability", g.grid_stability), ("confidence_overall", S.compute_confidence(g).overall_confidence)):
        assert abs(v - e[k]) <= 1e-3 * max(abs(v), abs(e[k])) + 1e-6, (k, v, e[k])  # north-star tolerance: 1e-3 relative
    assert np.allclose(g.beat_grid.beats[:4], e["beats_head"], rtol=0, atol=1e-4)
