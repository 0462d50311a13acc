"""Developer aid: run tracks through the CUDA path and the CPU oracle, print a field-by-field diff
(and per-stage intermediate diffs for single tracks).  Usage: python tools/dev_parity.py [seconds ...]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402
import synth  # noqa: E402

import stratum_dsp_b200 as S  # noqa: E402


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def cmp_arrays(name, g, o, tol=1e-4):
    if g is None or o is None:
        print(f"   {name}: missing gpu={g is None} oracle={o is None}")
        return
    n = min(len(g), len(o))
    if len(g) < len(o):
        print(f"   {name}: length gpu {len(g)} < oracle {len(o)}")
    if n == 0:
        print(f"   {name}: empty")
        return
    g, o = np.asarray(g[:n], np.float64), np.asarray(o[:n], np.float64)
    d = np.abs(g - o)
    scale = max(np.abs(o).max(), 1e-30)
    exact = int((g == o).sum())
    print(f"   {name}: n={n} exact={exact}/{n} max_abs={d.max():.3e} max_rel_to_peak={d.max() / scale:.3e} argmax={int(d.argmax())}")


def compare(x, sr, label, stages=False):
    t0 = time.time()
    o = O.analyze(x, sr, dump=stages, fast=True)
    t1 = time.time()
    S.debug_enable(stages)
    try:
        g = S.analyze_audio(x, sr)
        gerr = None
    except S.AnalysisError as e:
        g, gerr = None, e
    t2 = time.time()
    print(f"== {label}: oracle {t1 - t0:.2f}s gpu {t2 - t1:.2f}s")
    if gerr or o.status:
        print("   status gpu:", gerr, "| oracle:", o.status, o.error)
        return gerr is not None and o.status == gerr.code
    ok = True
    def chk(name, a, b, exact=False, tol=1e-3):
        nonlocal ok
        good = (a == b) if exact else rel(a, b) <= tol or abs(a - b) < 1e-6
        ok &= bool(good)
        print(f"   {'ok ' if good else 'BAD'} {name}: gpu={a} oracle={b}")
    chk("bpm", g.bpm, o.bpm)
    chk("bpm_conf", g.bpm_confidence, o.bpm_confidence)
    chk("key", g.key.id, o.key, exact=True)
    chk("key_conf", g.key_confidence, o.key_confidence)
    chk("key_clarity", g.key_clarity, o.key_clarity)
    chk("grid_stability", g.grid_stability, o.grid_stability)
    chk("trim", (g.trim_start, g.trim_end), (int(o.trim_start), int(o.trim_end)), exact=True)
    chk("n_onsets", len(g.onsets), len(o.onsets), exact=True)
    chk("onsets_equal", bool(np.array_equal(g.onsets, o.onsets)), True, exact=True)
    chk("hmm_frames_equal", bool(np.array_equal(g.hmm_beat_frames, o.hmm_beat_frames)), True, exact=True)
    chk("n_beats", len(g.beat_grid.beats), len(o.beats), exact=True)
    if len(g.beat_grid.beats) == len(o.beats) and len(o.beats):
        chk("beats_maxdiff", float(np.abs(g.beat_grid.beats - o.beats).max()), 0.0, tol=0, exact=False)
    chk("n_downbeats", len(g.beat_grid.downbeats), len(o.downbeats), exact=True)
    chk("time_sig", g.time_sig_beats_per_bar, int(o.time_sig_beats_per_bar), exact=True)
    chk("refined", g.beats_refined, int(o.beats_refined), exact=True)
    chk("mr_trig", g.metadata.tempogram_multi_res_triggered, None if o.multi_res_triggered < 0 else bool(o.multi_res_triggered), exact=True)
    chk("mr_used", g.metadata.tempogram_multi_res_used, None if o.multi_res_used < 0 else bool(o.multi_res_used), exact=True)
    chk("warnings", g.metadata.confidence_warnings, o.warning_strings, exact=True)
    if stages:
        D = S.debug_array
        cmp_arrays("onset.spectral_flux", D("onset.spectral_flux"), o.farray("onset.spectral_flux"))
        for nm in ("onset.energy", "onset.spectral", "onset.hfc"):
            cmp_arrays(nm, D(nm), o.iarray(nm))
        for v in ("full", "low", "mid", "high", "mel"):
            cmp_arrays(f"base.nov.{v}", D(f"base.nov.{v}"), o.farray(f"base.nov.{v}"))
        cmp_arrays("key.hpcp_raw", D("key.hpcp_raw"), o.farray("key.hpcp_raw"))
        cmp_arrays("key.hpcp_smooth", D("key.hpcp_smooth"), o.farray("key.hpcp_smooth"))
        cmp_arrays("key.energy", D("key.energy"), o.farray("key.energy"))
        cmp_arrays("key.weights", D("key.weights"), o.farray("key.weights"))
        cmp_arrays("hmm.path", D("hmm.path"), o.iarray("hmm.path"))
        ss = D("key.seg_scores")
        if ss is not None:
            ss = ss.reshape(-1, 24)
            print("   key.segments oracle (top key, clarity):", o.farray("key.segments"))
            np.savez(ROOT / "gpurun_out" / f"dbg_{label.split()[0]}_{int(time.time()*1000)%100000}.npz", seg_scores=ss, oracle_segments=o.farray("key.segments"),
                     oracle_scores=o.farray("key.scores"), oracle_order=o.farray("key.order"), gpu_clarity=g.key_clarity, oracle_clarity=o.key_clarity)
            for i, row in enumerate(ss):
                print(f"   gpu seg_scores[{i}] top={int(row.argmax())} max={row.max():.5f} min={row.min():.5f} sum={row.sum():.5f}")
        be = D("base.est")
        print("   base.est gpu:", None if be is None else be[:11], " oracle:", o.farray("base.est"), o.farray("legacy.est"))
        gc = D("base.cands")
        if gc is not None:
            gc = gc.reshape(-1, 4)
            ob, osc = o.farray("base.cands.bpm"), o.farray("base.cands.score")
            k = min(8, len(ob))
            print("   cands gpu   :", [(round(float(a), 3), round(float(b), 4)) for a, b in gc[:k, :2]])
            print("   cands oracle:", [(round(float(a), 3), round(float(b), 4)) for a, b in zip(ob[:k], osc[:k])])
    return ok


def main():
    secs = [float(a) for a in sys.argv[1:]] or [20.0]
    sr = 44100
    allok = True
    # STFT bit-exactness
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(44100 * 2) * 0.3).astype(np.float32)
    for frame, hop in ((2048, 512), (8192, 512), (2048, 256)):
        g, o = S.stft(x, frame, hop), O.stft(x, frame, hop)
        print(f"stft {frame}/{hop}: shape {g.shape} bit-exact={np.array_equal(g, o)} maxdiff={np.abs(g - o).max():.3e}")
        allok &= bool(np.array_equal(g, o))
    for s in secs:
        p = synth.c1_params(int(s * sr), sr)
        allok &= compare(synth.render(p), sr, f"C1 128bpm {s:.0f}s", stages=True)
        for i in range(3):
            p = synth.c2_params(i, int(s * sr), sr)
            allok &= compare(synth.render(p), sr, f"C2[{i}] bpm={p.bpm} tonic={p.tonic} minor={p.minor} {s:.0f}s", stages=(i == 0))
    allok &= compare(synth.fixture_kick(120.0, 8.0), sr, "fixture 120bpm_4bar")
    allok &= compare(synth.fixture_cmajor_scale(), sr, "fixture cmajor_scale")
    allok &= compare(synth.fixture_mixed_silence(), sr, "fixture mixed_silence", stages=True)
    allok &= compare(np.zeros(44100, np.float32), sr, "all silent")
    print("ALL OK" if allok else "MISMATCHES")


if __name__ == "__main__":
    main()
