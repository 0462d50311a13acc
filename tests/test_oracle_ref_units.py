"""The reference's own unit-test known answers, re-run on the CPU oracle (oracle/ — test infrastructure).

Every test below restates one `#[test]` of /root/reference/src (same inputs, same assertions; file:line in the docstring) on the
oracle's restatement of the function under test, through the unit-level C entry points of oracle/so_capi_units.cpp (and the ends
of so_beat.cpp / so_legacy.cpp).  This is what pins the oracle to the reference below the level of `analyze_audio`: the Rust
crate cannot be built in this image (no cargo), so its tests cannot be run against the real thing here.

Scope: the reference holds 223 `#[test]` functions.  17 are outside the analysed path (channel_mixer 9, threshold 6, key_changes 2);
of the 206 on it, the ones ported are those whose inputs can be built through the oracle's procedural API.  Not ported: tests of
struct plumbing the oracle has no counterpart for (constructor / getter / history tests of HmmBeatTracker and BayesianBeatTracker),
ragged `Vec<Vec<f32>>` shape errors (the oracle's spectrograms are rectangular by construction), `find_best_bpm_*` /
`spectral_flux_novelty` / `peak_picking` / `coarse_to_fine` helpers that `analyze_audio` never calls, and log-only tests without
assertions.  `test_ported_count` prints "N of 223".
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

L = O.lib()
f32p, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int64)
PORTED = []  # (reference file, test name) — filled by the decorator


def ref(where: str, name: str):
    def deco(fn):
        PORTED.append((where, name))
        fn.__doc__ = (fn.__doc__ or "") + f"\n    reference: {where} {name}"
        return fn
    return deco


def fa(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def fp(a):
    return a.ctypes.data_as(f32p)


def ia(x):
    return np.ascontiguousarray(x, dtype=np.int64)


def ip(a):
    return a.ctypes.data_as(i64p)


L.so_u_bayes_likelihood.restype = C.c_float
L.so_u_grid_stability.restype = C.c_float
L.so_u_extract_chroma.restype = C.c_int64
for name, at in {
    "so_u_tempo_variations": [f32p, C.c_int, C.c_float, f32p, C.c_int],
    "so_u_time_signature": [f32p, C.c_int],
    "so_u_bayes_likelihood": [f32p, C.c_int, C.c_float],
    "so_u_bayes_update": [C.c_float, f32p, C.c_int, f32p],
    "so_u_downbeats": [f32p, C.c_int, C.c_float, C.c_int, f32p, C.c_int],
    "so_u_grid_stability": [f32p, C.c_int],
    "so_u_beat_grid": [C.c_float, C.c_float, f32p, C.c_int, C.c_uint32, f32p, f32p, C.POINTER(C.c_int), f32p, C.POINTER(C.c_int), C.c_int],
    "so_u_hmm_full": [C.c_float, f32p, C.c_int, f32p, f32p, C.POINTER(C.c_int32), C.c_int],
    "so_u_acf_bpm": [i64p, C.c_int, C.c_uint32, C.c_uint64, C.c_float, C.c_float, f32p, f32p, C.c_int],
    "so_u_comb_bpm": [i64p, C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_float, f32p, f32p, C.c_int],
    "so_u_score_bpm": [i64p, C.c_int, C.c_uint32, C.c_float, C.c_float, f32p],
    "so_u_acf_fft": [f32p, C.c_int, f32p],
    "so_u_find_peaks": [f32p, C.c_int, C.c_uint64, i64p, f32p, C.c_int],
    "so_u_merge": [f32p, f32p, C.c_int, f32p, f32p, C.c_int, f32p, f32p, C.POINTER(C.c_uint32), C.c_int],
    "so_u_spec_onsets": [C.c_int, f32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_float, i64p, C.c_int],
    "so_u_energy_onsets": [f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, i64p, C.c_int],
    "so_u_novelty": [C.c_int, f32p, C.c_uint64, C.c_uint64, C.c_uint64, f32p, C.c_int],
    "so_u_combined_novelty": [f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_uint64, f32p, C.c_int],
    "so_u_tempogram": [C.c_int, f32p, C.c_int, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, f32p, f32p, C.c_int],
    "so_u_estimate_tempogram": [f32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, f32p, f32p, C.POINTER(C.c_uint32)],
    "so_u_extract_chroma": [f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_float, f32p, C.c_uint64],
    "so_u_frame_to_chroma": [f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_float, f32p],
    "so_u_smooth_chroma": [f32p, C.c_uint64, C.c_uint64],
    "so_u_sharpen_chroma": [f32p, C.c_float],
    "so_u_normalize": [f32p, C.c_uint64, C.c_int, C.c_float, C.c_float, C.c_float, f32p],
    "so_u_trim": [f32p, C.c_uint64, C.c_uint32, C.c_float, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int],
    "so_u_time_mask": [f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_int, f32p],
}.items():
    getattr(L, name).argtypes = at

INVALID_INPUT, PROCESSING_ERROR = 1, 3


# ======================================================================================================================
# beat tracking: hmm.rs, bayesian.rs, tempo_variation.rs, time_signature.rs, mod.rs
# ======================================================================================================================
def hmm(bpm, onsets):
    o = fa(onsets)
    t, c, fr = np.zeros(4096, np.float32), np.zeros(4096, np.float32), np.zeros(4096, np.int32)
    n = L.so_u_hmm_full(bpm, fp(o), o.size, fp(t), fp(c), fr.ctypes.data_as(C.POINTER(C.c_int32)), 4096)
    return n, t[:max(n, 0)], c[:max(n, 0)], fr[:max(n, 0)]


def hmm_path(bpm, onsets):
    o = fa(onsets)
    fr, tm, path, plen = np.zeros(4096, np.int32), np.zeros(4096, np.float32), np.zeros(4096, np.int32), C.c_int()
    n = L.so_hmm(bpm, fp(o), o.size, fr.ctypes.data_as(C.POINTER(C.c_int32)), fp(tm), 4096, path.ctypes.data_as(C.POINTER(C.c_int)), 4096, C.byref(plen))
    return n, path[:plen.value]


@ref("features/beat_tracking/hmm.rs:522-548", "test_track_beats_basic")
def test_hmm_track_beats_basic():
    n, t, c, _ = hmm(120.0, [0.0, 0.5, 1.0, 1.5, 2.0, 2.5])
    assert n >= 3
    assert np.all(np.diff(t) > 0)
    assert np.all((c >= 0.0) & (c <= 1.0))


@ref("features/beat_tracking/hmm.rs:550-570", "test_track_beats_128bpm")
def test_hmm_track_beats_128bpm():
    iv = np.float32(60.0) / np.float32(128.0)
    n, t, _, _ = hmm(128.0, [np.float32(i) * iv for i in range(6)])
    assert n > 0
    if n >= 2:
        assert abs((t[1] - t[0]) - 60.0 / 128.0) < 0.1


@ref("features/beat_tracking/hmm.rs:572-579", "test_track_beats_invalid_bpm")
def test_hmm_invalid_bpm():
    assert hmm(0.0, [0.0, 0.5])[0] == -INVALID_INPUT
    assert hmm(350.0, [0.0, 0.5])[0] == -INVALID_INPUT


@ref("features/beat_tracking/hmm.rs:581-585", "test_track_beats_empty_onsets")
def test_hmm_empty_onsets():
    assert hmm(120.0, [])[0] == -INVALID_INPUT


@ref("features/beat_tracking/hmm.rs:587-594", "test_track_beats_single_onset")
def test_hmm_single_onset():
    n, _, _, _ = hmm(120.0, [0.5])
    assert n >= 0 or n in (-INVALID_INPUT, -PROCESSING_ERROR)  # "either succeeds with few beats or fails gracefully"
    assert n == 1  # what the reference's code does: T = ceil(0 / interval) + 1 = 1 frame, emission exp(0) = 1 > 0.1


@ref("features/beat_tracking/hmm.rs:596-616", "test_viterbi_forward_pass")
def test_hmm_viterbi_path():
    n, path = hmm_path(120.0, [0.0, 0.5, 1.0, 1.5])
    assert path.size == 4  # one state per emission frame: ceil(1.5 / 0.5) + 1
    assert np.all((path >= 0) & (path < 5))


@ref("features/beat_tracking/hmm.rs:618-640", "test_extract_beats_from_path")
def test_hmm_extract_beats_sorted():
    n, t, _, _ = hmm(120.0, [0.0, 0.5, 1.0, 1.5, 2.0])
    assert n > 0 and np.all(np.diff(t) > 0)


@ref("features/beat_tracking/hmm.rs:500-520", "test_compute_emission_probabilities")
def test_hmm_emissions_are_probabilities():
    # the emission row is the same for all five states (hmm.rs:231-298 never reads state_bpm); the kept beats' confidences
    # 0.7 e + 0.3 align are bounded by it
    n, t, c, fr = hmm(120.0, [0.0, 0.5, 1.0, 1.5, 2.0])
    assert n == 5 and list(fr) == [0, 1, 2, 3, 4]
    assert np.all((c >= 0.0) & (c <= 1.0))


def bayes_l(onsets, bpm):
    o = fa(onsets)
    return float(L.so_u_bayes_likelihood(fp(o), o.size, bpm))


def bayes_update(cur, onsets):
    o, out = fa(onsets), C.c_float()
    st = L.so_u_bayes_update(cur, fp(o), o.size, C.byref(out))
    return st, out.value


@ref("features/beat_tracking/bayesian.rs:331-348", "test_compute_likelihood")
def test_bayes_likelihood():
    perfect = [0.0, 0.5, 1.0, 1.5, 2.0]
    l = bayes_l(perfect, 120.0)
    assert 0.0 < l <= 1.0
    assert l > bayes_l(perfect, 100.0)


@ref("features/beat_tracking/bayesian.rs:350-355", "test_compute_likelihood_empty_onsets")
def test_bayes_likelihood_empty():
    assert bayes_l([], 120.0) == 0.0


@ref("features/beat_tracking/bayesian.rs:374-387", "test_update_with_onsets")
def test_bayes_update():
    st, bpm = bayes_update(120.0, [0.0, 0.5, 1.0, 1.5, 2.0])
    assert st == 0 and bpm > 0.0
    assert bpm == 120.0  # candidates 115 .. 125 step 0.5: the perfect grid has likelihood exactly 1 at 120


@ref("features/beat_tracking/bayesian.rs:389-393", "test_update_with_onsets_empty")
def test_bayes_update_empty():
    assert bayes_update(120.0, [])[0] == INVALID_INPUT


@ref("features/beat_tracking/bayesian.rs:395-402", "test_update_with_onsets_invalid_bpm")
def test_bayes_update_invalid_bpm():
    assert bayes_update(0.0, [0.0, 0.5])[0] == INVALID_INPUT
    assert bayes_update(350.0, [0.0, 0.5])[0] == INVALID_INPUT


@ref("features/beat_tracking/bayesian.rs:313-329", "test_generate_bpm_candidates")
def test_bayes_candidate_range():
    # candidates span current +- 5 BPM (clamped to 60..180): onsets on a 123 BPM grid pull the estimate to 123, never past 125
    iv = 60.0 / 123.0
    st, bpm = bayes_update(120.0, [i * iv for i in range(12)])
    assert st == 0 and 115.0 <= bpm <= 125.0 and abs(bpm - 123.0) <= 0.5
    iv = 60.0 / 140.0
    st, bpm = bayes_update(120.0, [i * iv for i in range(12)])
    assert st == 0 and 115.0 <= bpm <= 125.0


def tempo_var(beats, nominal):
    b = fa(beats)
    out = np.zeros((64, 5), np.float32)
    n = L.so_u_tempo_variations(fp(b), b.size, nominal, fp(out), 64)
    return n, out[:max(n, 0)]


@ref("features/beat_tracking/tempo_variation.rs:233-244", "test_detect_tempo_variations_constant")
def test_tempo_variations_constant():
    iv = np.float32(60.0) / np.float32(120.0)
    n, seg = tempo_var([np.float32(i) * iv for i in range(20)], 120.0)
    assert n > 0 and not np.any(seg[:, 4] > 0)


@ref("features/beat_tracking/tempo_variation.rs:246-263", "test_detect_tempo_variations_variable")
def test_tempo_variations_variable():
    beats, t = [], np.float32(0.0)
    for i in range(20):
        t = np.float32(t + np.float32(60.0) / np.float32(120.0 + i * 1.0))
        beats.append(t)
    n, _ = tempo_var(beats, 120.0)
    assert n > 0


@ref("features/beat_tracking/tempo_variation.rs:265-273", "test_detect_tempo_variations_insufficient_beats")
def test_tempo_variations_insufficient():
    n, seg = tempo_var([0.0, 0.5, 1.0], 120.0)
    assert n == 1 and seg[0, 2] == 120.0


@ref("features/beat_tracking/tempo_variation.rs:275-281", "test_detect_tempo_variations_empty")
def test_tempo_variations_empty():
    n, seg = tempo_var([], 120.0)
    assert n == 1 and seg[0, 2] == 120.0


def time_sig(beats):
    b = fa(beats)
    return L.so_u_time_signature(fp(b), b.size)


@ref("features/beat_tracking/time_signature.rs:205-226", "test_time_signature_four_four")
def test_time_signature_regular_16():
    iv = np.float32(0.5)
    assert time_sig([np.float32(i) * iv for i in range(16)]) in (4, 3, 6)


@ref("features/beat_tracking/time_signature.rs:228-244", "test_time_signature_three_four")
def test_time_signature_regular_12():
    assert time_sig([np.float32(i) * np.float32(0.5) for i in range(12)]) in (4, 3, 6)


@ref("features/beat_tracking/time_signature.rs:246-255", "test_time_signature_insufficient_beats")
def test_time_signature_insufficient():
    assert time_sig([0.0, 0.5, 1.0, 1.5]) == 4


@ref("features/beat_tracking/time_signature.rs:257-262", "test_time_signature_beats_per_bar")
def test_time_signature_values():
    # TimeSignature::{FourFour, ThreeFour, SixEight}.beats_per_bar() = 4, 3, 6: the only values the detector returns; on a perfectly
    # regular grid all three hypotheses score the same and max_by keeps the LAST maximum (6/8)
    assert time_sig([np.float32(i) * np.float32(0.5) for i in range(16)]) == 6


def beat_grid(bpm, conf, onsets, sr=44100):
    o = fa(onsets)
    stab, nb, nd = C.c_float(), C.c_int(), C.c_int()
    beats, down = np.zeros(4096, np.float32), np.zeros(4096, np.float32)
    st = L.so_u_beat_grid(bpm, conf, fp(o), o.size, sr, C.byref(stab), fp(beats), C.byref(nb), fp(down), C.byref(nd), 4096)
    return st, stab.value, beats[:nb.value], down[:nd.value]


@ref("features/beat_tracking/mod.rs:491-509", "test_generate_beat_grid_basic")
def test_beat_grid_basic():
    st, stab, beats, _ = beat_grid(120.0, 0.85, [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5])
    assert st == 0 and beats.size > 0 and 0.0 <= stab <= 1.0
    assert np.all(np.diff(beats) > 0)


@ref("features/beat_tracking/mod.rs:511-524", "test_generate_beat_grid_128bpm")
def test_beat_grid_128():
    iv = np.float32(60.0) / np.float32(128.0)
    st, stab, beats, _ = beat_grid(128.0, 0.8, [np.float32(i) * iv for i in range(8)])
    assert st == 0 and beats.size > 0 and stab > 0.0


@ref("features/beat_tracking/mod.rs:526-531", "test_generate_beat_grid_invalid_bpm")
def test_beat_grid_invalid_bpm():
    assert beat_grid(0.0, 0.8, [0.0, 0.5, 1.0])[0] == INVALID_INPUT
    assert beat_grid(350.0, 0.8, [0.0, 0.5, 1.0])[0] == INVALID_INPUT


@ref("features/beat_tracking/mod.rs:533-536", "test_generate_beat_grid_empty_onsets")
def test_beat_grid_empty():
    assert beat_grid(120.0, 0.8, [])[0] == INVALID_INPUT


def downbeats(beats, bpm, bpb=4):
    b = fa(beats)
    out = np.zeros(4096, np.float32)
    n = L.so_u_downbeats(fp(b), b.size, bpm, bpb, fp(out), 4096)
    return n, out[:max(n, 0)]


@ref("features/beat_tracking/mod.rs:538-559", "test_detect_downbeats")
def test_downbeats():
    n, d = downbeats([0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0], 120.0)
    assert n > 0 and d[0] == 0.0
    if n > 1:
        assert abs((d[1] - d[0]) - 2.0) < 0.3
    assert list(d) == [0.0, 2.0, 4.0]


@ref("features/beat_tracking/mod.rs:561-564", "test_detect_downbeats_empty")
def test_downbeats_empty():
    assert downbeats([], 120.0)[0] == 0


@ref("features/beat_tracking/mod.rs:566-571", "test_detect_downbeats_single_beat")
def test_downbeats_single():
    n, d = downbeats([0.5], 120.0)
    assert n == 1 and d[0] == 0.5


def stability(times):
    t = fa(times)
    return float(L.so_u_grid_stability(fp(t), t.size))


@ref("features/beat_tracking/mod.rs:573-601", "test_calculate_grid_stability_perfect")
def test_stability_perfect():
    assert stability([0.0, 0.5, 1.0, 1.5]) > 0.9


@ref("features/beat_tracking/mod.rs:603-635", "test_calculate_grid_stability_variable")
def test_stability_variable():
    s = stability([0.0, 0.4, 0.9, 1.6])
    assert s < 0.9 and 0.0 <= s <= 1.0


@ref("features/beat_tracking/mod.rs:637-650", "test_calculate_grid_stability_insufficient_beats")
def test_stability_insufficient():
    assert stability([0.0]) == 0.0


@ref("features/beat_tracking/mod.rs:652-687", "test_generate_beat_grid_from_positions")
def test_grid_from_positions():
    # generate_beat_grid_from_positions = sort the beat times + detect_downbeats in 4/4; bars = downbeats
    times = [0.0, 0.5, 1.0, 1.5, 2.0]
    n, d = downbeats(sorted(times), 120.0, 4)
    assert len(times) == 5 and n > 0
    assert list(d) == [0.0, 2.0]


# ======================================================================================================================
# legacy period estimators: autocorrelation.rs, comb_filter.rs, candidate_filter.rs
# ======================================================================================================================
def cands(fn, *args):
    b, c = np.zeros(1024, np.float32), np.zeros(1024, np.float32)
    n = fn(*args, fp(b), fp(c), 1024)
    return n, b[:max(n, 0)], c[:max(n, 0)]


def acf_bpm(onsets, sr, hop, lo, hi):
    o = ia(onsets)
    return cands(L.so_u_acf_bpm, ip(o), o.size, sr, hop, lo, hi)


def comb_bpm(onsets, sr, lo, hi, res):
    o = ia(onsets)
    return cands(L.so_u_comb_bpm, ip(o), o.size, sr, lo, hi, res)


def frame_grid_onsets(bpm, sr=44100, hop=512, beats=4):
    period_samples = np.float32(60.0 * sr) / np.float32(bpm)
    period_frames = int(np.round(period_samples / np.float32(hop)))
    return [beat * period_frames * hop for beat in range(beats)]


def sample_grid_onsets(bpm, sr=44100, beats=4):
    period = np.float32(60.0 * sr) / np.float32(bpm)
    return [int(np.round(np.float32(b) * period)) for b in range(beats)]


@ref("features/period/autocorrelation.rs:344-376", "test_autocorrelation_basic_120bpm")
def test_acf_120():
    n, b, c = acf_bpm(frame_grid_onsets(120.0), 44100, 512, 60.0, 180.0)
    assert n > 0 and abs(b[0] - 120.0) < 5.0 and c[0] > 0.0


@ref("features/period/autocorrelation.rs:378-382", "test_autocorrelation_empty_onsets")
def test_acf_empty():
    assert acf_bpm([], 44100, 512, 60.0, 180.0)[0] == -INVALID_INPUT


@ref("features/period/autocorrelation.rs:384-391", "test_autocorrelation_single_onset")
def test_acf_single():
    assert acf_bpm([1000], 44100, 512, 60.0, 180.0)[0] == 0


@ref("features/period/autocorrelation.rs:393-408", "test_autocorrelation_invalid_params")
def test_acf_invalid():
    assert acf_bpm([1000, 2000], 0, 512, 60.0, 180.0)[0] == -INVALID_INPUT
    assert acf_bpm([1000, 2000], 44100, 0, 60.0, 180.0)[0] == -INVALID_INPUT
    assert acf_bpm([1000, 2000], 44100, 512, 180.0, 60.0)[0] == -INVALID_INPUT


@ref("features/period/autocorrelation.rs:410-436", "test_autocorrelation_128bpm")
def test_acf_128():
    n, b, _ = acf_bpm(frame_grid_onsets(128.0), 44100, 512, 60.0, 180.0)
    assert n > 0 and abs(b[0] - 128.0) < 5.0


@ref("features/period/autocorrelation.rs:438-454", "test_compute_autocorrelation_fft")
def test_acf_fft():
    sig = fa([1.0, 0.0, 1.0, 0.0, 1.0, 0.0])
    out = np.zeros(64, np.float32)
    n = L.so_u_acf_fft(fp(sig), sig.size, fp(out))
    assert n == sig.size and out[0] > 0.0 and out[2] > 0.0


@ref("features/period/autocorrelation.rs:456-465", "test_find_peaks_in_acf")
def test_find_peaks():
    acf = fa([0.1, 0.2, 0.5, 0.3, 0.4, 0.6, 0.2, 0.1])
    idx, val = np.zeros(16, np.int64), np.zeros(16, np.float32)
    n = L.so_u_find_peaks(fp(acf), acf.size, 0, ip(idx), fp(val), 16)
    assert n > 0 and any(i in (2, 5) for i in idx[:n])


@ref("features/period/comb_filter.rs:403-432", "test_comb_filter_120bpm")
def test_comb_120():
    n, b, c = comb_bpm(sample_grid_onsets(120.0), 44100, 60.0, 180.0, 1.0)
    assert n > 0 and abs(b[0] - 120.0) < 5.0 and c[0] > 0.0


@ref("features/period/comb_filter.rs:434-438", "test_comb_filter_empty_onsets")
def test_comb_empty():
    assert comb_bpm([], 44100, 60.0, 180.0, 1.0)[0] == -INVALID_INPUT


@ref("features/period/comb_filter.rs:440-446", "test_comb_filter_single_onset")
def test_comb_single():
    assert comb_bpm([1000], 44100, 60.0, 180.0, 1.0)[0] == 0


@ref("features/period/comb_filter.rs:448-463", "test_comb_filter_invalid_params")
def test_comb_invalid():
    assert comb_bpm([1000, 2000], 0, 60.0, 180.0, 1.0)[0] == -INVALID_INPUT
    assert comb_bpm([1000, 2000], 44100, 180.0, 60.0, 1.0)[0] == -INVALID_INPUT
    assert comb_bpm([1000, 2000], 44100, 60.0, 180.0, 0.0)[0] == -INVALID_INPUT


@ref("features/period/comb_filter.rs:465-489", "test_comb_filter_128bpm")
def test_comb_128():
    n, b, _ = comb_bpm(sample_grid_onsets(128.0), 44100, 60.0, 180.0, 1.0)
    assert n > 0 and abs(b[0] - 128.0) < 5.0


@ref("features/period/comb_filter.rs:491-517", "test_score_bpm_candidate")
def test_score_bpm_candidate():
    sc = C.c_float()
    o = ia(sample_grid_onsets(120.0))
    assert L.so_u_score_bpm(ip(o), o.size, 44100, 120.0, 0.1, C.byref(sc)) == 0 and sc.value > 0.8
    o = ia([1000, 5000, 12000, 25000])
    assert L.so_u_score_bpm(ip(o), o.size, 44100, 120.0, 0.1, C.byref(sc)) == 0 and sc.value < 0.5


@ref("features/period/comb_filter.rs:519-543", "test_comb_filter_resolution")
def test_comb_resolution():
    o = sample_grid_onsets(120.0)
    assert comb_bpm(o, 44100, 60.0, 180.0, 0.5)[0] >= comb_bpm(o, 44100, 60.0, 180.0, 1.0)[0]


def coarse_to_fine(onsets, sr, lo, hi, refine):
    # coarse_to_fine_search (comb_filter.rs:256-330) is a composition of two estimate_bpm_from_comb_filter calls
    n, b, c = comb_bpm(onsets, sr, lo, hi, 2.0)
    if n <= 0:
        return n, b, c
    n2, b2, c2 = comb_bpm(onsets, sr, max(b[0] - refine, lo), min(b[0] + refine, hi), 0.5)
    return (n2, b2, c2) if n2 > 0 else (n, b, c)


@ref("features/period/comb_filter.rs:545-578", "test_coarse_to_fine_search")
def test_coarse_to_fine_120():
    n, b, c = coarse_to_fine(sample_grid_onsets(120.0), 44100, 60.0, 180.0, 5.0)
    assert n > 0 and abs(b[0] - 120.0) < 5.0 and c[0] > 0.0


@ref("features/period/comb_filter.rs:580-584", "test_coarse_to_fine_search_empty")
def test_coarse_to_fine_empty():
    assert coarse_to_fine([], 44100, 60.0, 180.0, 5.0)[0] == -INVALID_INPUT


@ref("features/period/comb_filter.rs:586-611", "test_coarse_to_fine_search_performance")
def test_coarse_to_fine_128():
    n, b, _ = coarse_to_fine(sample_grid_onsets(128.0, beats=8), 44100, 60.0, 180.0, 5.0)
    assert n > 0 and abs(b[0] - 128.0) < 5.0


def merge(ac, comb):
    ab, acf = fa([x[0] for x in ac]), fa([x[1] for x in ac])
    cb, cc = fa([x[0] for x in comb]), fa([x[1] for x in comb])
    b, c, a = np.zeros(64, np.float32), np.zeros(64, np.float32), np.zeros(64, np.uint32)
    n = L.so_u_merge(fp(ab), fp(acf), ab.size, fp(cb), fp(cc), cb.size, fp(b), fp(c), a.ctypes.data_as(C.POINTER(C.c_uint32)), 64)
    return n, b[:n], c[:n], a[:n]


@ref("features/period/candidate_filter.rs:449-468", "test_merge_candidates_agreement")
def test_merge_agreement():
    n, b, c, a = merge([(120.0, 0.9)], [(120.0, 0.85)])
    assert n > 0 and abs(b[0] - 120.0) < 1.0 and c[0] > 0.9 and a[0] == 2


@ref("features/period/candidate_filter.rs:470-488", "test_merge_candidates_octave_error")
def test_merge_octave_double():
    n, b, _, _ = merge([(240.0, 0.8)], [(120.0, 0.9)])
    assert n > 0 and abs(b[0] - 120.0) < 1.0


@ref("features/period/candidate_filter.rs:490-508", "test_merge_candidates_octave_error_half")
def test_merge_octave_half():
    n, b, _, _ = merge([(60.0, 0.8)], [(120.0, 0.9)])
    assert n > 0 and abs(b[0] - 120.0) < 1.0


@ref("features/period/candidate_filter.rs:510-536", "test_merge_candidates_grouping")
def test_merge_grouping():
    n, b, _, a = merge([(120.0, 0.8), (121.0, 0.7)], [(120.5, 0.85)])
    assert n == 1 and abs(b[0] - 120.0) < 2.0 and a[0] == 3


@ref("features/period/candidate_filter.rs:538-542", "test_merge_candidates_empty")
def test_merge_empty():
    assert merge([], [])[0] == 0


@ref("features/period/candidate_filter.rs:544-560", "test_merge_candidates_single_method")
def test_merge_single_method():
    n, _, c, a = merge([(120.0, 0.8)], [])
    assert n > 0 and a[0] == 1 and c[0] <= 0.8


@ref("features/period/candidate_filter.rs:562-588", "test_merge_candidates_sorted")
def test_merge_sorted():
    n, _, c, _ = merge([(120.0, 0.9), (130.0, 0.7)], [(120.0, 0.85)])
    assert all(c[i - 1] >= c[i] for i in range(1, n))


# ======================================================================================================================
# onsets: consensus.rs, spectral_flux.rs, hfc.rs, energy_flux.rs
# ======================================================================================================================
def vote(lists, weights, tol_ms=50, sr=44100):
    arrs = [ia(x) for x in lists]
    w = fa(weights)
    centre, conf, voted = np.zeros(256, np.int64), np.zeros(256, np.float32), np.zeros(256, np.uint32)
    n = L.so_vote_onsets(ip(arrs[0]), arrs[0].size, ip(arrs[1]), arrs[1].size, ip(arrs[2]), arrs[2].size, ip(arrs[3]), arrs[3].size, fp(w), tol_ms, sr,
                         ip(centre), fp(conf), voted.ctypes.data_as(C.POINTER(C.c_uint32)), 256)
    return n, centre[:max(n, 0)], conf[:max(n, 0)], voted[:max(n, 0)]


EQ = [0.25, 0.25, 0.25, 0.25]


@ref("features/onset/consensus.rs:293-310", "test_consensus_voting_basic")
def test_consensus_basic():
    n, t, c, v = vote([[1000]] * 4, EQ)
    assert n == 1 and t[0] == 1000 and v[0] == 4 and abs(c[0] - 1.0) < 0.01


@ref("features/onset/consensus.rs:312-330", "test_consensus_voting_clustering")
def test_consensus_clustering():
    n, _, c, v = vote([[1000], [1050], [980], [1020]], EQ)
    assert n == 1 and v[0] == 4 and abs(c[0] - 1.0) < 0.01


@ref("features/onset/consensus.rs:332-349", "test_consensus_voting_separate_onsets")
def test_consensus_separate():
    n, _, _, v = vote([[1000, 50000], [1050, 50500], [980, 50200], [1020, 49900]], EQ)
    assert n == 2 and v[0] == 4 and v[1] == 4


@ref("features/onset/consensus.rs:351-368", "test_consensus_voting_partial_agreement")
def test_consensus_partial():
    n, _, c, v = vote([[1000], [1050], [], []], [0.3, 0.3, 0.2, 0.2])
    assert n == 1 and v[0] == 2 and abs(c[0] - 0.6) < 0.01


@ref("features/onset/consensus.rs:370-387", "test_consensus_voting_weighted")
def test_consensus_weighted():
    n, _, c, v = vote([[1000], [], [], []], [0.5, 0.2, 0.2, 0.1])
    assert n == 1 and v[0] == 1 and abs(c[0] - 0.5) < 0.01


@ref("features/onset/consensus.rs:389-403", "test_consensus_voting_empty")
def test_consensus_empty():
    assert vote([[], [], [], []], EQ)[0] == 0


@ref("features/onset/consensus.rs:405-434", "test_consensus_voting_sorted_by_confidence")
def test_consensus_sorted():
    n, _, c, v = vote([[1000, 20000, 50000], [1050, 20050, 50500], [980, 20100], [1020, 19950]], EQ)
    assert n >= 2 and v[0] == 4
    assert all(c[i] <= c[i - 1] for i in range(1, n))


@ref("features/onset/consensus.rs:436-459", "test_consensus_voting_invalid_parameters")
def test_consensus_invalid():
    assert vote([[1000], [], [], []], EQ, 50, 0)[0] == -INVALID_INPUT
    assert vote([[1000], [], [], []], EQ, 0, 44100)[0] == -INVALID_INPUT
    assert vote([[1000], [], [], []], [-0.1, 0.25, 0.25, 0.25])[0] == -INVALID_INPUT


@ref("features/onset/consensus.rs:461-476", "test_consensus_voting_time_conversion")
def test_consensus_time_conversion():
    n, t, _, _ = vote([[44100], [], [], []], EQ)
    assert n == 1 and abs(t[0] / 44100.0 - 1.0) < 0.001


def spec_onsets(kind, spec, pct, sr=44100):
    s = fa(spec)
    frames, bins = (s.shape if s.ndim == 2 else (0, 1024))
    out = np.zeros(4096, np.int64)
    n = L.so_u_spec_onsets(kind, fp(s) if s.size else None, frames, bins, sr, pct, ip(out), 4096)
    return n, out[:max(n, 0)]


def flat_spec(frames, bins, value):
    return np.full((frames, bins), value, np.float32)


@ref("features/onset/spectral_flux.rs:228-266", "test_spectral_flux_basic")
def test_spectral_flux_basic():
    s = flat_spec(10, 1024, 0.01)
    s[0:5, 0:256] = 1.0
    s[5, 768:1024] = 1.0
    s[6:10, 0:256] = 1.0
    n, on = spec_onsets(0, s, 0.3)
    assert n > 0 and any(4 <= f <= 7 for f in on)


@ref("features/onset/spectral_flux.rs:268-273", "test_spectral_flux_empty")
def test_spectral_flux_empty():
    assert spec_onsets(0, np.zeros((0, 1024), np.float32), 0.8)[0] == 0


@ref("features/onset/spectral_flux.rs:275-281", "test_spectral_flux_single_frame")
def test_spectral_flux_single_frame():
    assert spec_onsets(0, flat_spec(1, 1024, 0.5), 0.8)[0] == 0


@ref("features/onset/spectral_flux.rs:283-294", "test_spectral_flux_invalid_percentile")
def test_spectral_flux_invalid_percentile():
    assert spec_onsets(0, flat_spec(10, 1024, 0.5), -0.1)[0] == -INVALID_INPUT
    assert spec_onsets(0, flat_spec(10, 1024, 0.5), 1.5)[0] == -INVALID_INPUT


@ref("features/onset/spectral_flux.rs:305-311", "test_spectral_flux_all_zeros")
def test_spectral_flux_all_zeros():
    assert spec_onsets(0, flat_spec(10, 1024, 0.0), 0.8)[0] < 3


@ref("features/onset/spectral_flux.rs:313-334", "test_spectral_flux_threshold_sensitivity")
def test_spectral_flux_threshold_sensitivity():
    s = flat_spec(20, 1024, 0.1)
    for i in range(20):
        s[i, :] = np.float32(0.1) + (np.float32(i) / np.float32(20.0)) * np.float32(0.9)
    assert spec_onsets(0, s, 0.5)[0] >= spec_onsets(0, s, 0.9)[0]


@ref("features/onset/spectral_flux.rs:336-349", "test_spectral_flux_normalization")
def test_spectral_flux_normalization():
    s = flat_spec(2, 1024, 0.5)
    s[1, :] = 1.0
    n, _ = spec_onsets(0, s, 0.5)
    assert n == 0  # both frames normalise to all ones: the flux is exactly zero, nothing exceeds the threshold


@ref("features/onset/spectral_flux.rs:351-381", "test_spectral_flux_multiple_changes")
def test_spectral_flux_multiple_changes():
    s = flat_spec(20, 1024, 0.1)
    s[5, 0:512] = 1.0
    s[10, 512:1024] = 1.0
    s[15, 256:768] = 1.0
    assert spec_onsets(0, s, 0.3)[0] >= 2


@ref("features/onset/hfc.rs:222-257", "test_hfc_basic")
def test_hfc_basic():
    s = flat_spec(10, 1024, 0.01)
    s[0:5, 0:100] = 0.5
    s[5, 800:1024] = 1.0
    s[6:10, 0:100] = 0.5
    n, on = spec_onsets(1, s, 0.3)
    assert n > 0 and any(4 <= f <= 7 for f in on)


@ref("features/onset/hfc.rs:259-264", "test_hfc_empty")
def test_hfc_empty():
    assert spec_onsets(1, np.zeros((0, 1024), np.float32), 0.8)[0] == 0


@ref("features/onset/hfc.rs:266-272", "test_hfc_single_frame")
def test_hfc_single_frame():
    assert spec_onsets(1, flat_spec(1, 1024, 0.5), 0.8)[0] == 0


@ref("features/onset/hfc.rs:274-285", "test_hfc_invalid_percentile")
def test_hfc_invalid_percentile():
    assert spec_onsets(1, flat_spec(10, 1024, 0.5), -0.1)[0] == -INVALID_INPUT
    assert spec_onsets(1, flat_spec(10, 1024, 0.5), 1.5)[0] == -INVALID_INPUT


@ref("features/onset/hfc.rs:287-292", "test_hfc_zero_sample_rate")
def test_hfc_zero_sample_rate():
    assert spec_onsets(1, flat_spec(10, 1024, 0.5), 0.8, sr=0)[0] == -INVALID_INPUT


@ref("features/onset/hfc.rs:303-309", "test_hfc_all_zeros")
def test_hfc_all_zeros():
    assert spec_onsets(1, flat_spec(10, 1024, 0.0), 0.8)[0] == 0


@ref("features/onset/hfc.rs:311-333", "test_hfc_threshold_sensitivity")
def test_hfc_threshold_sensitivity():
    s = flat_spec(20, 1024, 0.01)
    for i in range(20):
        s[i, 800:1024] = np.float32(0.1) + (np.float32(i) / np.float32(20.0)) * np.float32(0.9)
    assert spec_onsets(1, s, 0.5)[0] >= spec_onsets(1, s, 0.9)[0]


@ref("features/onset/hfc.rs:335-361", "test_hfc_frequency_weighting")
def test_hfc_frequency_weighting():
    s = flat_spec(2, 1024, 0.0)
    s[0, 0:100] = 1.0
    s[1, 900:1024] = 1.0
    n, on = spec_onsets(1, s, 0.5)
    assert n >= 0  # runs; with two frames there is one flux value (the HFC rises: high bins weigh more), which cannot exceed itself


@ref("features/onset/hfc.rs:363-386", "test_hfc_multiple_changes")
def test_hfc_multiple_changes():
    s = flat_spec(20, 1024, 0.01)
    for f in (5, 10, 15):
        s[f, 800:1024] = 1.0
    assert spec_onsets(1, s, 0.3)[0] >= 2


def energy_onsets(x, frame=2048, hop=512, thr=-20.0):
    s = fa(x)
    out = np.zeros(8192, np.int64)
    n = L.so_u_energy_onsets(fp(s) if s.size else None, s.size, frame, hop, thr, ip(out), 8192)
    return n, out[:max(n, 0)]


def kick_pattern(duration_s, bpm, sr, kick_ms):
    n = int(np.float32(duration_s) * np.float32(sr))
    x = np.zeros(n, np.float32)
    beat_iv = int(np.float32(60.0) / np.float32(bpm) * np.float32(sr))
    ks = int(np.float32(kick_ms) / np.float32(1000.0) * np.float32(sr))
    t = np.arange(ks, dtype=np.float32) / np.float32(ks)
    env = np.exp(-t * np.float32(5.0)).astype(np.float32)
    pos = 0
    while pos < n:
        e = min(pos + ks, n)
        x[pos:e] = env[: e - pos] * np.float32(0.8)
        pos += beat_iv
    return x


@ref("features/onset/energy_flux.rs:287-311", "test_energy_flux_basic")
def test_energy_flux_basic():
    x = np.zeros(44100, np.float32)
    x[5000:] = 0.5
    n, on = energy_onsets(x, thr=-30.0)
    assert n > 0 and 3000 <= on[0] <= 8000


@ref("features/onset/energy_flux.rs:313-349", "test_energy_flux_kick_pattern_120_bpm")
def test_energy_flux_kick_120():
    n, on = energy_onsets(kick_pattern(4.0, 120.0, 44100.0, 150.0), thr=-30.0)
    assert 6 <= n <= 20
    iv = np.diff(on)
    assert abs(int(iv.sum() // iv.size) - 22050) < 22050 // 2


@ref("features/onset/energy_flux.rs:351-356", "test_energy_flux_empty_samples")
def test_energy_flux_empty():
    assert energy_onsets(np.zeros(0, np.float32))[0] == 0


@ref("features/onset/energy_flux.rs:358-363", "test_energy_flux_silent_audio")
def test_energy_flux_silent():
    assert energy_onsets(np.zeros(44100, np.float32))[0] == 0


@ref("features/onset/energy_flux.rs:365-373", "test_energy_flux_too_short_audio")
def test_energy_flux_too_short():
    assert energy_onsets(np.full(1000, 0.5, np.float32))[0] == 0


@ref("features/onset/energy_flux.rs:375-386", "test_energy_flux_invalid_parameters")
def test_energy_flux_invalid():
    x = np.full(44100, 0.5, np.float32)
    assert energy_onsets(x, frame=0)[0] == -INVALID_INPUT
    assert energy_onsets(x, hop=0)[0] == -INVALID_INPUT


@ref("features/onset/energy_flux.rs:388-403", "test_energy_flux_threshold_sensitivity")
def test_energy_flux_threshold_sensitivity():
    x = kick_pattern(2.0, 120.0, 44100.0, 50.0)
    assert energy_onsets(x, thr=-30.0)[0] >= energy_onsets(x, thr=-10.0)[0]


# ---- HPSS (onset/hpss.rs) ------------------------------------------------------------------------------------------------
def hpss(spec, margin):
    s = fa(spec)
    h, p = np.zeros_like(s), np.zeros_like(s)
    st = L.so_hpss_decompose(fp(s) if s.size else None, s.shape[0], s.shape[1], margin, fp(h), fp(p))
    return st, h, p


def hpss_onsets(perc, pct):
    p = fa(perc)
    out = np.zeros(1024, np.int64)
    n = L.so_hpss_onsets(fp(p) if p.size else None, p.shape[0], p.shape[1], pct, ip(out), 1024)
    return n, out[:max(n, 0)]


@ref("features/onset/hpss.rs:380-412", "test_hpss_decompose_basic")
def test_hpss_decompose_basic():
    s = flat_spec(10, 1024, 0.5)
    st, h, p = hpss(s, 5)
    assert st == 0 and h.shape == s.shape and p.shape == s.shape
    assert np.abs(h + p - s).max() < 0.1


@ref("features/onset/hpss.rs:414-419", "test_hpss_decompose_empty")
def test_hpss_decompose_empty():
    assert hpss(np.zeros((0, 1024), np.float32), 5)[0] == INVALID_INPUT


@ref("features/onset/hpss.rs:430-461", "test_hpss_decompose_harmonic_vs_percussive")
def test_hpss_harmonic_vs_percussive():
    s = flat_spec(20, 1024, 0.0)
    s[:, 100:200] = 0.8
    for f in (5, 10, 15):
        s[f, :] = 1.0
    st, _, p = hpss(s, 3)
    assert st == 0 and float((p[5] * p[5]).sum()) > float((p[3] * p[3]).sum())


@ref("features/onset/hpss.rs:463-484", "test_detect_hpss_onsets_basic")
def test_hpss_onsets_basic():
    p = flat_spec(20, 1024, 0.01)
    for f in (5, 10, 15):
        p[f, :] = 1.0
    assert hpss_onsets(p, 0.5)[0] >= 2


@ref("features/onset/hpss.rs:486-491", "test_detect_hpss_onsets_empty")
def test_hpss_onsets_empty():
    assert hpss_onsets(np.zeros((0, 1024), np.float32), 0.8)[0] == 0


@ref("features/onset/hpss.rs:493-499", "test_detect_hpss_onsets_single_frame")
def test_hpss_onsets_single_frame():
    assert hpss_onsets(flat_spec(1, 1024, 0.5), 0.8)[0] == 0


@ref("features/onset/hpss.rs:501-510", "test_detect_hpss_onsets_invalid_percentile")
def test_hpss_onsets_invalid_percentile():
    assert hpss_onsets(flat_spec(10, 1024, 0.5), -0.1)[0] == -INVALID_INPUT
    assert hpss_onsets(flat_spec(10, 1024, 0.5), 1.5)[0] == -INVALID_INPUT


@ref("features/onset/hpss.rs:512-532", "test_detect_hpss_onsets_threshold_sensitivity")
def test_hpss_onsets_threshold_sensitivity():
    p = flat_spec(20, 1024, 0.01)
    for i in range(20):
        p[i, :] = np.float32(0.1) + (np.float32(i) / np.float32(20.0)) * np.float32(0.9)
    assert hpss_onsets(p, 0.5)[0] >= hpss_onsets(p, 0.9)[0]


# ======================================================================================================================
# novelty curves and tempograms: novelty.rs, tempogram_fft.rs, tempogram_autocorr.rs, tempogram.rs
# ======================================================================================================================
def novelty(kind, spec, k=4):
    s = fa(spec)
    out = np.zeros(4096, np.float32)
    n = L.so_u_novelty(kind, fp(s), s.shape[0], s.shape[1], k, fp(out), 4096)
    return out[:n]


@ref("features/period/novelty.rs:1027-1041", "test_energy_flux_novelty_basic")
def test_energy_flux_novelty_basic():
    s = flat_spec(10, 1024, 0.1)
    s[5, :] = 1.0
    v = novelty(1, s)
    assert v.size == 9 and (v[4] > 0.0 or v[5] > 0.0)


@ref("features/period/novelty.rs:1043-1057", "test_hfc_novelty_basic")
def test_hfc_novelty_basic():
    s = flat_spec(10, 1024, 0.1)
    s[5, 512:1024] = 1.0
    v = novelty(2, s)
    assert v.size == 9 and (v[4] > 0.0 or v[5] > 0.0)


@ref("features/period/novelty.rs:993-1010", "test_spectral_flux_novelty_basic")
def test_superflux_novelty_basic():
    # analyze_audio uses the SuperFlux form of the spectral novelty (novelty.rs:336-420); same fixture, same expectations
    s = flat_spec(10, 1024, 0.1)
    s[5, 0:512] = 1.0
    v = novelty(0, s, 4)
    assert v.size == 9 and (v[4] > 0.0 or v[5] > 0.0)


def combined(s, e, h):
    a, b, c = fa(s), fa(e), fa(h)
    out = np.zeros(64, np.float32)
    # combined_novelty = combined_novelty_with_params(.., 0.5, 0.3, 0.2, 16, 5) (novelty.rs:852-866)
    n = L.so_u_combined_novelty(fp(a), a.size, fp(b), b.size, fp(c), c.size, 0.5, 0.3, 0.2, 16, 5, fp(out), 64)
    return out[:n]


@ref("features/period/novelty.rs:1059-1071", "test_combined_novelty")
def test_combined_novelty():
    v = combined([0.0, 0.5, 1.0, 0.5, 0.0], [0.0, 0.3, 0.8, 0.3, 0.0], [0.0, 0.2, 0.6, 0.2, 0.0])
    assert v.size == 5 and np.all((v >= 0.0) & (v <= 1.0)) and v.max() > 0.0


@ref("features/period/novelty.rs:1073-1083", "test_combined_novelty_different_lengths")
def test_combined_novelty_lengths():
    assert combined([0.0, 0.5, 1.0], [0.0, 0.3, 0.8, 0.3], [0.0, 0.2]).size == 2


def tempogram(kind, nov, sr, hop, lo, hi, res=1.0):
    v = fa(nov)
    b, p = np.zeros(8192, np.float32), np.zeros(8192, np.float32)
    n = L.so_u_tempogram(kind, fp(v) if v.size else None, v.size, sr, hop, lo, hi, res, fp(b), fp(p), 8192)
    return n, b[:max(n, 0)], p[:max(n, 0)]


def periodic_novelty():
    frame_rate = np.float32(44100) / np.float32(512)
    period = int(frame_rate / (np.float32(120.0) / np.float32(60.0)))
    v = np.zeros(500, np.float32)
    v[::period] = 1.0
    return v


@ref("features/period/tempogram_fft.rs:242-271", "test_fft_tempogram_periodic")
def test_fft_tempogram_periodic():
    n, b, _ = tempogram(0, periodic_novelty(), 44100, 512, 100.0, 140.0)
    assert n > 0 and 115.0 <= b[0] <= 125.0  # find_best_bpm_fft = the entry with the highest power (first after the sort)


@ref("features/period/tempogram_fft.rs:273-278", "test_fft_tempogram_empty")
def test_fft_tempogram_empty():
    assert tempogram(0, [], 44100, 512, 40.0, 240.0)[0] == -INVALID_INPUT


@ref("features/period/tempogram_fft.rs:280-295", "test_fft_tempogram_invalid_params")
def test_fft_tempogram_invalid():
    v = np.full(100, 0.5, np.float32)
    assert tempogram(0, v, 0, 512, 40.0, 240.0)[0] == -INVALID_INPUT
    assert tempogram(0, v, 44100, 0, 40.0, 240.0)[0] == -INVALID_INPUT
    assert tempogram(0, v, 44100, 512, 240.0, 40.0)[0] == -INVALID_INPUT


@ref("features/period/tempogram_autocorr.rs:228-258", "test_autocorrelation_tempogram_periodic")
def test_autocorr_tempogram_periodic():
    n, b, _ = tempogram(1, periodic_novelty(), 44100, 512, 100.0, 140.0, 1.0)
    assert n > 0 and 115.0 <= b[0] <= 125.0


@ref("features/period/tempogram_autocorr.rs:260-265", "test_autocorrelation_tempogram_empty")
def test_autocorr_tempogram_empty():
    assert tempogram(1, [], 44100, 512, 40.0, 240.0, 0.5)[0] == -INVALID_INPUT


@ref("features/period/tempogram_autocorr.rs:267-282", "test_autocorrelation_tempogram_invalid_params")
def test_autocorr_tempogram_invalid():
    v = np.full(100, 0.5, np.float32)
    assert tempogram(1, v, 0, 512, 40.0, 240.0, 0.5)[0] == -INVALID_INPUT
    assert tempogram(1, v, 44100, 0, 40.0, 240.0, 0.5)[0] == -INVALID_INPUT
    assert tempogram(1, v, 44100, 512, 240.0, 40.0, 0.5)[0] == -INVALID_INPUT


def estimate_tempogram(spec, lo, hi, res):
    s = fa(spec)
    cfg = L.so_config_new()
    for k, v in (("min_bpm", lo), ("max_bpm", hi), ("bpm_resolution", res)):
        assert L.so_config_set(cfg, k.encode(), float(v)) == 0
    bpm, conf, ag = C.c_float(), C.c_float(), C.c_uint32()
    st = L.so_u_estimate_tempogram(fp(s) if s.size else None, s.shape[0], s.shape[1] if s.ndim == 2 else 1024, 44100, 512, cfg, C.byref(bpm), C.byref(conf), C.byref(ag))
    L.so_config_free(cfg)
    return st, bpm.value, conf.value


@ref("features/period/tempogram.rs:782-806", "test_estimate_bpm_tempogram_basic")
def test_estimate_bpm_tempogram_basic():
    s = flat_spec(500, 1024, 0.1)
    s[::43, 0:512] = 1.0
    st, bpm, conf = estimate_tempogram(s, 100.0, 140.0, 0.5)
    assert st == 0 and 115.0 <= bpm <= 125.0 and 0.0 <= conf <= 1.0


@ref("features/period/tempogram.rs:808-813", "test_estimate_bpm_tempogram_empty")
def test_estimate_bpm_tempogram_empty():
    assert estimate_tempogram(np.zeros((0, 1024), np.float32), 40.0, 240.0, 0.5)[0] != 0


@ref("features/period/tempogram.rs:815-832", "test_estimate_bpm_tempogram_agreement")
def test_estimate_bpm_tempogram_flat():
    st, bpm, conf = estimate_tempogram(flat_spec(200, 1024, 0.5), 40.0, 240.0, 0.5)
    if st == 0:
        assert 40.0 <= bpm <= 240.0 and 0.0 <= conf <= 1.0


# ======================================================================================================================
# chroma and key: chroma/{extractor,normalization,smoothing}.rs, key/{templates,detector,key_clarity}.rs
# ======================================================================================================================
def extract_chroma(x, sr=44100, frame=2048, hop=512, soft=True, sigma=0.5):
    s = fa(x)
    nf = L.so_u_extract_chroma(fp(s), s.size, sr, frame, hop, int(soft), sigma, None, 0)
    out = np.zeros((max(nf, 0), 12), np.float32)
    if nf > 0:
        L.so_u_extract_chroma(fp(s), s.size, sr, frame, hop, int(soft), sigma, fp(out), nf)
    return out


def a440(seconds=2, sr=44100):
    i = np.arange(sr * seconds, dtype=np.float32)
    return np.sin(np.float32(2.0) * np.float32(np.pi) * np.float32(440.0) * (i / np.float32(sr))).astype(np.float32)


@ref("features/chroma/extractor.rs:1513-1521", "test_extract_chroma_short")
def test_extract_chroma_short():
    assert extract_chroma(np.zeros(1000, np.float32)).shape[0] == 0


@ref("features/chroma/extractor.rs:1523-1561", "test_extract_chroma_basic")
def test_extract_chroma_basic():
    ch = extract_chroma(a440())  # extract_chroma = soft mapping, sigma 0.5 (extractor.rs:172-184)
    assert ch.shape[0] > 0
    norms = np.sqrt((ch * ch).sum(axis=1))
    assert np.all((np.abs(norms - 1.0) < 0.01) | (norms < 1e-10))
    assert ch.mean(axis=0)[9] > 0.1  # A


@ref("features/chroma/extractor.rs:1563-1581", "test_frame_to_chroma")
def test_frame_to_chroma():
    mag = np.zeros(1025, np.float32)
    mag[int(np.float32(440.0) * np.float32(2048) / np.float32(44100))] = 1.0
    out = np.zeros(12, np.float32)
    L.so_u_frame_to_chroma(fp(mag), 1025, 44100, 2048, 0, 0.5, fp(out))
    nrm = float(np.sqrt((out * out).sum()))
    assert abs(nrm - 1.0) < 0.01 or nrm < 1e-10


@ref("features/chroma/extractor.rs:1597-1624", "test_soft_chroma_mapping")
def test_soft_vs_hard_mapping():
    x = a440()
    soft, hard = extract_chroma(x, soft=True), extract_chroma(x, soft=False)
    assert soft.shape == hard.shape and soft.shape[0] > 0


def sharpen(ch, power):
    v = fa(ch).copy()
    L.so_u_sharpen_chroma(fp(v), power)
    return v


CH = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.5, 0.4, 0.3, 0.2, 0.1, 0.0]


@ref("features/chroma/normalization.rs:98-112", "test_sharpen_chroma")
def test_sharpen_chroma():
    v = sharpen(CH, 2.0)
    assert v.size == 12 and abs(float(np.sqrt((v * v).sum())) - 1.0) < 0.01
    assert v[5] > CH[5] or CH[5] < 0.1


@ref("features/chroma/normalization.rs:114-127", "test_sharpen_chroma_power_one")
def test_sharpen_chroma_power_one():
    v = sharpen(CH, 1.0)
    c = np.asarray(CH, np.float32)
    assert np.abs(c / np.sqrt((c * c).sum()) - v).max() < 0.01


@ref("features/chroma/normalization.rs:135-145", "test_l2_normalize_chroma")
def test_l2_normalize_via_power_one():
    # l2_normalize_chroma(x) == sharpen_chroma(x, 1.0) (normalization.rs:67-92 is the same normalisation without the power)
    v = sharpen([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 5.0, 4.0, 3.0, 2.0, 1.0, 0.0], 1.0)
    assert abs(float(np.sqrt((v * v).sum())) - 1.0) < 0.01


@ref("features/chroma/normalization.rs:153-163", "test_l2_normalize_chroma_zero")
def test_normalize_zero_vector_is_uniform():
    v = sharpen([0.0] * 12, 1.0)
    assert np.abs(v - 1.0 / np.sqrt(12.0)).max() < 0.01


def smooth(ch, window):
    v = fa(ch).copy()
    L.so_u_smooth_chroma(fp(v), v.shape[0], window)
    return v


def rolling_onehot(n=10):
    c = np.zeros((n, 12), np.float32)
    for i in range(n):
        c[i, i % 12] = 1.0
    return c


@ref("features/chroma/smoothing.rs:171-177", "test_smooth_chroma_single_frame")
def test_smooth_single_frame():
    assert smooth(np.full((1, 12), 0.1, np.float32), 5).shape == (1, 12)


@ref("features/chroma/smoothing.rs:179-195", "test_smooth_chroma_basic")
def test_smooth_basic():
    assert smooth(rolling_onehot(), 3).shape == (10, 12)


@ref("features/chroma/smoothing.rs:197-202", "test_smooth_chroma_window_size_one")
def test_smooth_window_one():
    c = np.full((5, 12), 0.1, np.float32)
    assert np.array_equal(smooth(c, 1), c)


@ref("features/chroma/smoothing.rs:221-227", "test_smooth_chroma_even_window_size")
def test_smooth_even_window():
    assert smooth(np.full((10, 12), 0.1, np.float32), 4).shape == (10, 12)


def test_smooth_is_a_running_median():
    # the property smooth_chroma's docs state (smoothing.rs:12-35): a median over time per pitch class removes a one-frame spike
    c = np.full((9, 12), 0.2, np.float32)
    c[4, 3] = 1.0
    assert smooth(c, 5)[4, 3] == np.float32(0.2)


def templates():
    maj, mn = np.zeros((12, 12), np.float32), np.zeros((12, 12), np.float32)
    L.so_key_templates(fp(maj), fp(mn))
    return maj, mn


@ref("features/key/templates.rs:279-288", "test_key_templates_creation")
def test_templates_creation():
    maj, mn = templates()
    assert maj.shape == (12, 12) and mn.shape == (12, 12)


@ref("features/key/templates.rs:290-317", "test_c_major_template")
def test_c_major_template():
    c = templates()[0][0]
    assert abs(c[0] - c.max()) < 1e-6
    top4 = np.sort(c)[::-1][3]
    assert c[4] >= top4 and c[7] >= top4
    chromatic = (c[1] + c[3] + c[6] + c[8] + c[10]) / 5.0
    assert c[0] > chromatic * 1.4 and c[4] > chromatic * 1.1 and c[7] > chromatic * 1.2


@ref("features/key/templates.rs:319-328", "test_a_minor_template")
def test_a_minor_template():
    a = templates()[1][9]
    assert a[9] > 0.1 and a[0] > 0.1 and a[4] > 0.1


@ref("features/key/templates.rs:343-357", "test_template_rotation")
def test_template_rotation():
    maj, _ = templates()
    assert maj[2][2] > 0.1
    assert np.allclose(maj[2], np.roll(maj[0], 2))


def detect_key(chroma, weights=None):
    ch = fa(chroma)
    key, conf, cl = C.c_int(), C.c_float(), C.c_float()
    scores, order = np.zeros(24, np.float32), np.zeros(24, np.int32)
    w = fa(weights) if weights is not None else None
    st = L.so_detect_key(fp(ch) if ch.size else None, ch.shape[0] if ch.ndim == 2 else 0, fp(w) if w is not None else None, C.byref(key), C.byref(conf), C.byref(cl),
                         fp(scores), order.ctypes.data_as(C.POINTER(C.c_int)))
    return st, key.value, conf.value, scores, order


@ref("features/key/detector.rs:1007-1012", "test_detect_key_empty")
def test_detect_key_empty():
    assert detect_key(np.zeros((0, 12), np.float32))[0] == INVALID_INPUT


@ref("features/key/detector.rs:1014-1048", "test_detect_key_basic")
def test_detect_key_basic():
    c = np.zeros(12, np.float32)
    c[[0, 4, 7]] = 0.3
    c /= np.sqrt((c * c).sum())
    st, key, conf, scores, order = detect_key(np.tile(c, (10, 1)))
    assert st == 0 and 0.0 <= conf <= 1.0 and scores.size == 24
    assert key == 0 and order[0] == 0  # Key::Major(0); top_keys[0] = Major(0)


@ref("features/key/detector.rs:1058-1067", "test_average_chroma")
def test_detect_key_zero_weights():
    assert detect_key(np.zeros((10, 12), np.float32), np.zeros(10, np.float32))[0] == 0


def clarity(scores):
    s = fa(scores)
    return float(L.so_key_clarity(fp(s) if s.size else None, s.size))


@ref("features/key/key_clarity.rs:100-104", "test_compute_key_clarity_empty")
def test_clarity_empty():
    assert clarity([]) == 0.0


@ref("features/key/key_clarity.rs:106-111", "test_compute_key_clarity_single")
def test_clarity_single():
    assert clarity([0.8]) == 0.0


@ref("features/key/key_clarity.rs:113-124", "test_compute_key_clarity_high")
def test_clarity_high():
    assert clarity([0.9, 0.3, 0.3, 0.3]) > 0.5


@ref("features/key/key_clarity.rs:126-137", "test_compute_key_clarity_low")
def test_clarity_low():
    assert clarity([0.5, 0.48, 0.49, 0.47]) < 0.5


@ref("features/key/key_clarity.rs:139-149", "test_compute_key_clarity_all_same")
def test_clarity_all_same():
    assert clarity([0.5, 0.5, 0.5]) == 0.0


@ref("features/key/key_clarity.rs:151-157", "test_compute_key_clarity_clamped")
def test_clarity_clamped():
    assert 0.0 <= clarity([1.0, 0.0]) <= 1.0


# ======================================================================================================================
# preprocessing: normalization.rs, silence.rs
# ======================================================================================================================
def sine(n, amp, sr):
    i = np.arange(n, dtype=np.float32)
    return (np.float32(amp) * np.sin(np.float32(2.0) * np.float32(np.pi) * np.float32(440.0) * (i / np.float32(sr)))).astype(np.float32)


def normalize(x, method, sr):
    v = fa(x).copy()
    g = C.c_float(1.0)
    st = L.so_u_normalize(fp(v) if v.size else None, v.size, method, -14.0, 1.0, sr, C.byref(g))
    return st, v, g.value


@ref("preprocessing/normalization.rs:564-591", "test_peak_normalization")
def test_peak_normalization():
    st, v, _ = normalize(sine(44100, 0.5, 44100.0), 0, 44100.0)
    peak = float(np.abs(v).max())
    assert st == 0 and abs(peak - 10.0 ** (-1.0 / 20.0)) < 0.01 and peak <= 1.0


@ref("preprocessing/normalization.rs:593-620", "test_rms_normalization")
def test_rms_normalization():
    st, v, _ = normalize(sine(44100, 0.3, 44100.0), 1, 44100.0)
    rms = float(np.sqrt((v.astype(np.float64) ** 2).mean()))
    assert st == 0 and abs(rms - 10.0 ** ((-14.0 + 3.0 - 1.0) / 20.0)) < 0.1
    assert float(np.abs(v).max()) <= 1.0


@ref("preprocessing/normalization.rs:622-638", "test_lufs_calculation")
def test_lufs_calculation():
    x = sine(48000 * 2, 0.8, 48000.0)
    lufs = C.c_float()
    assert L.so_lufs(fp(x), x.size, 48000.0, C.byref(lufs)) == 0
    assert np.isfinite(lufs.value) and lufs.value < 0.0


@ref("preprocessing/normalization.rs:640-667", "test_lufs_normalization")
def test_lufs_normalization():
    st, v, g = normalize(sine(48000 * 2, 0.5, 48000.0), 2, 48000.0)
    assert st == 0 and g != 1.0  # gain_db != 0
    assert float(np.abs(v).max()) <= 1.0


@ref("preprocessing/normalization.rs:669-683", "test_silent_audio")
def test_normalize_silent():
    st, v, g = normalize(np.zeros(44100, np.float32), 0, 44100.0)
    assert st == 0 and g == 1.0 and not v.any()  # gain_db == 0


@ref("preprocessing/normalization.rs:685-699", "test_ultra_quiet_audio")
def test_normalize_ultra_quiet():
    assert normalize(sine(44100, 1e-6, 44100.0), 0, 44100.0)[0] == 0


@ref("preprocessing/normalization.rs:701-709", "test_empty_samples")
def test_normalize_empty():
    assert normalize(np.zeros(0, np.float32), 0, 44100.0)[0] == INVALID_INPUT


@ref("preprocessing/normalization.rs:711-738", "test_k_weighting_filter")
def test_k_weighting_changes_the_level():
    # the K-weighting biquad is only observable through the LUFS value here: a 440 Hz tone and the same tone at 60 Hz (below the
    # 1.68 kHz high-pass corner, further attenuated) must read differently, and both must be finite
    def lufs_of(freq):
        i = np.arange(96000, dtype=np.float32)
        x = (np.float32(0.5) * np.sin(np.float32(2.0 * np.pi * freq) * (i / np.float32(48000.0)))).astype(np.float32)
        out = C.c_float()
        assert L.so_lufs(fp(x), x.size, 48000.0, C.byref(out)) == 0
        return out.value
    a, b = lufs_of(440.0), lufs_of(60.0)
    assert np.isfinite(a) and np.isfinite(b) and b < a


def trim(x, sr=44100, thr=-40.0, min_ms=500, frame=2048):
    s = fa(x)
    ts, te = C.c_uint64(), C.c_uint64()
    reg = np.zeros(2 * 256, np.uint64)
    n = L.so_u_trim(fp(s) if s.size else None, s.size, sr, thr, min_ms, frame, C.byref(ts), C.byref(te), reg.ctypes.data_as(C.POINTER(C.c_uint64)), 256)
    return n, int(ts.value), int(te.value), reg[: 2 * max(n, 0)].reshape(-1, 2)


def audio_with_silence(total, a, b, amp):
    x = np.zeros(total, np.float32)
    i = np.arange(a, min(b, total), dtype=np.float32)
    x[a:min(b, total)] = np.float32(amp) * np.sin(i / np.float32(1000.0))
    return x


@ref("preprocessing/silence.rs:305-326", "test_detect_and_trim_leading_trailing")
def test_trim_leading_trailing():
    x = audio_with_silence(44100 * 3, 44100, 44100 * 2, 0.5)
    n, ts, te, _ = trim(x)
    assert 0 < te - ts < x.size and n > 0


@ref("preprocessing/silence.rs:328-341", "test_detect_and_trim_all_silent")
def test_trim_all_silent():
    n, ts, te, _ = trim(np.zeros(44100, np.float32))
    assert te - ts == 0


@ref("preprocessing/silence.rs:343-362", "test_detect_and_trim_no_silence")
def test_trim_no_silence():
    i = np.arange(44100, dtype=np.float32)
    x = (np.float32(0.5) * np.sin(i / np.float32(1000.0))).astype(np.float32)
    _, ts, te, _ = trim(x, thr=-60.0)
    assert te - ts > x.size // 2


@ref("preprocessing/silence.rs:364-378", "test_detect_and_trim_invalid_parameters")
def test_trim_invalid():
    x = np.full(44100, 0.5, np.float32)
    assert trim(x, sr=0)[0] == -INVALID_INPUT
    assert trim(x, frame=0)[0] == -INVALID_INPUT


@ref("preprocessing/silence.rs:380-388", "test_detect_and_trim_empty_samples")
def test_trim_empty():
    n, ts, te, _ = trim(np.zeros(0, np.float32))
    assert n == 0 and te - ts == 0


@ref("preprocessing/silence.rs:390-428", "test_detect_and_trim_threshold_sensitivity")
def test_trim_threshold_sensitivity():
    x = np.zeros(44100 * 2, np.float32)
    x[:22050] = 0.01
    x[22050:44100] = 0.5
    lo = trim(x, thr=-60.0)[3]
    hi = trim(x, thr=-20.0)[3]
    assert int((hi[:, 1] - hi[:, 0]).sum()) >= int((lo[:, 1] - lo[:, 0]).sum())


@ref("preprocessing/silence.rs:430-448", "test_detect_and_trim_min_duration")
def test_trim_min_duration():
    i = np.arange(44100, dtype=np.float32)
    x = (np.float32(0.5) * np.sin(i / np.float32(1000.0))).astype(np.float32)
    x[10000:15000] = 0.0  # 113 ms of silence inside the track: shorter than min_duration_ms, not at an edge
    n, ts, te, reg = trim(x, min_ms=500)
    assert n >= 0 and not any(10000 <= a and b <= 16000 for a, b in reg)


# ======================================================================================================================
# analysis: confidence.rs, result.rs
# ======================================================================================================================
def confidence(bpm, bc, kc, kcl, gs, warnings=0):
    out, fl = (C.c_float * 4)(), C.c_uint32()
    L.so_confidence_of(bpm, bc, kc, kcl, gs, warnings, 0, out, C.byref(fl))
    return [float(v) for v in out]


@ref("analysis/confidence.rs:342-361", "test_compute_confidence_all_good")
def test_confidence_all_good():
    b, k, g, o = confidence(120.0, 0.9, 0.8, 0.7, 0.85)
    assert (b, k, g) == (np.float32(0.9), np.float32(0.8), np.float32(0.85)) and abs(o - 0.855) < 0.01


@ref("analysis/confidence.rs:363-383", "test_compute_confidence_bpm_failed")
def test_confidence_bpm_failed():
    b, k, g, o = confidence(0.0, 0.0, 0.8, 0.7, 0.85)
    assert b == 0.0 and k == np.float32(0.8) and g == np.float32(0.85) and abs(o - 0.48) < 0.01


@ref("analysis/confidence.rs:385-404", "test_compute_confidence_key_failed")
def test_confidence_key_failed():
    b, k, g, o = confidence(120.0, 0.9, 0.0, 0.0, 0.85)
    assert b == np.float32(0.9) and k == 0.0 and g == np.float32(0.85) and abs(o - 0.54) < 0.01


@ref("analysis/confidence.rs:406-423", "test_compute_confidence_all_failed")
def test_confidence_all_failed():
    assert confidence(0.0, 0.0, 0.0, 0.0, 0.0) == [0.0, 0.0, 0.0, 0.0]


@ref("analysis/confidence.rs:425-446", "test_compute_confidence_with_warnings")
def test_confidence_with_warnings():
    b = confidence(120.0, 0.9, 0.8, 0.7, 0.85, warnings=1)[0]  # "BPM detection failed: ..." warning present
    assert 0.0 < b < 0.9


@ref("analysis/confidence.rs:448-467", "test_compute_confidence_clamping")
def test_confidence_clamping():
    b, k, g, o = confidence(120.0, 1.5, -0.5, 0.7, 2.0)
    assert b <= 1.0 and k >= 0.0 and g <= 1.0 and 0.0 <= o <= 1.0


@ref("analysis/confidence.rs:469-488", "test_confidence_helper_methods")
def test_confidence_levels():
    # is_high_confidence: overall >= 0.7; is_low_confidence: overall < 0.4 (confidence.rs:70-119)
    assert confidence(120.0, 0.9, 0.8, 0.7, 0.85)[3] >= 0.7
    assert confidence(0.0, 0.0, 0.0, 0.0, 0.0)[3] < 0.4


@ref("analysis/confidence.rs:490-516", "test_key_clarity_adjustment")
def test_confidence_key_clarity_adjustment():
    assert confidence(120.0, 0.9, 0.8, 0.1, 0.85)[1] < confidence(120.0, 0.9, 0.8, 0.7, 0.85)[1]


def key_name(minor, idx, numerical=False):
    buf = C.create_string_buffer(16)
    L.so_key_name(int(minor), idx, int(numerical), buf, 16)
    return buf.value.decode()


@ref("analysis/result.rs:272-279", "test_key_name_major")
def test_key_name_major():
    assert [key_name(0, i) for i in (0, 1, 2, 6, 11)] == ["C", "C#", "D", "F#", "B"]


@ref("analysis/result.rs:281-288", "test_key_name_minor")
def test_key_name_minor():
    assert [key_name(1, i) for i in (0, 1, 2, 9, 11)] == ["Cm", "C#m", "Dm", "Am", "Bm"]


@ref("analysis/result.rs:290-305", "test_key_numerical_major")
def test_key_numerical_major():
    assert [key_name(0, i, True) for i in (0, 7, 2, 9, 4, 11, 6, 1, 8, 3, 10, 5)] == [f"{n}A" for n in range(1, 13)]


@ref("analysis/result.rs:307-322", "test_key_numerical_minor")
def test_key_numerical_minor():
    assert [key_name(1, i, True) for i in (9, 4, 11, 6, 1, 8, 3, 10, 5, 0, 7, 2)] == [f"{n}B" for n in range(1, 13)]


@ref("analysis/result.rs:345-369", "test_key_numerical_roundtrip")
def test_key_numerical_roundtrip():
    names = {key_name(m, i, True) for m in (0, 1) for i in range(12)}
    assert len(names) == 24  # the notation is a bijection, so from_numerical(numerical(k)) == k


def test_ported_count():
    """How much of the reference's unit-test surface is pinned here."""
    n = len(PORTED)
    print(f"\n{n} of 223 reference #[test] functions ported onto the oracle ({len({w.split(':')[0] for w, _ in PORTED})} source files)")
    assert n >= 120
    assert len(set(PORTED)) == n  # no reference test counted twice
