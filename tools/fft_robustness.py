#!/usr/bin/env python3
"""FFT-robustness campaign: how much of "bit-exact against the reference" survives an FFT whose roundings differ.

The reference's FFT is rustfft 6.2 (Cargo.toml:18) — absent from /root/reference and not reproducible (planner- and
SIMD-dependent butterfly order).  The oracle pins ONE arithmetic (SFFT, oracle/so_fft.cpp) and the CUDA kernels match it bit
for bit; this script re-runs the whole analysis under two other arithmetics of the same DFT (oracle/so_fft.cpp variants 1
and 2: float64 rounded once, and an f32 radix-2 DIT without fma on the full complex frame) and compares EVERY discrete output
(trim range, consensus onsets, HMM beat frames, beat / downbeat counts, key label, time signature, refinement flag,
escalation flags, warning and flag masks) plus the floats (1e-3 relative) with the SFFT run.  A discrete output that is
identical under all three arithmetics has a decision margin larger than FFT rounding noise, which is the only statement about
rustfft that can be made in this image; every flip is listed by track.

    python tools/fft_robustness.py --suite full  [--jobs 8] [--out tests/golden/fft_robustness.json]
    python tools/fft_robustness.py --suite quick                       # the subset tests/test_fft_robustness.py re-runs

Workloads (BASELINE.json configs): C1 (3-min click+chord), 64 tracks of C2, 32 of C5, C4 (60-min drifting mix).
Also counts the tracks whose key label comes out of the reference's HashMap-order vote (detector.rs:254-275).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

VARIANTS = {0: "sfft (parity arithmetic)", 1: "float64 rounded once", 2: "f32 radix-2 DIT, no fma, full complex frame"}
INT_VIEWS = ["trim_start", "trim_end", "key", "time_sig_beats_per_bar", "beats_refined", "multi_res_triggered", "multi_res_used", "warnings", "flags",
             "n_beats", "n_downbeats"]
FLOATS = ["bpm", "bpm_confidence", "key_confidence", "key_clarity", "grid_stability", "duration_seconds"]


def workload(suite: str):
    import synth

    items = []  # (label, kind, arg)
    if suite == "full":
        items.append(("C1", "c1", 7_938_000))
        items += [(f"C2[{i}]", "c2", (i, 7_938_000)) for i in range(64)]
        items += [(f"C5[{i}]", "c5", i) for i in range(32)]
        items.append(("C4", "c4", 158_760_000))
    else:  # quick: same generators, short renderings (CPU test budget)
        items.append(("C1/20s", "c1", 20 * 44100))
        items += [(f"C2[{i}]/20s", "c2", (i, 20 * 44100)) for i in range(8)]
        items += [(f"C5[{i}]/16s", "c5short", i) for i in range(4)]
        items.append(("C4/40s", "c4", 40 * 44100))
    return items


def render(kind, arg):
    import synth

    if kind == "c1":
        return synth.render(synth.c1_params(arg)), 44100
    if kind == "c2":
        return synth.render(synth.c2_params(arg[0], arg[1])), 44100
    if kind == "c5":
        p = synth.c5_params(arg)
        return synth.render(p), p.sample_rate
    if kind == "c5short":
        p = synth.c5_params(arg)
        p.n_samples = 16 * p.sample_rate
        return synth.render(p), p.sample_rate
    if kind == "c4":
        return synth.c4_mix(arg), 44100
    raise ValueError(kind)


def analyse(item):
    import oracle_lib as O

    label, kind, arg = item
    x, sr = render(kind, arg)
    L = O.lib(fast=True)
    out = {"label": label, "seconds": x.size / sr, "sr": sr, "runs": {}}
    for v in VARIANTS:
        L.so_set_fft_variant(v)
        r = O.analyze(x, sr, fast=True)
        rec = {k: int(getattr(r, k)) for k in ("trim_start", "trim_end", "time_sig_beats_per_bar", "beats_refined", "multi_res_triggered", "multi_res_used",
                                               "warnings", "flags", "key_hashmap_tie")}
        rec.update(status=int(r.status), key=int(r.key), n_beats=int(len(r.beats)), n_downbeats=int(len(r.downbeats)),
                   onsets=[int(v_) for v_ in r.onsets], hmm=[int(v_) for v_ in r.hmm_beat_frames])
        rec.update({k: float(getattr(r, k)) for k in FLOATS})
        out["runs"][v] = rec
    L.so_set_fft_variant(0)
    return out


def compare(rec):
    """Differences of variants 1, 2 against variant 0 for one track."""
    base = rec["runs"][0]
    flips, worst = [], 0.0
    for v in (1, 2):
        r = rec["runs"][v]
        if r["status"] != base["status"]:
            flips.append({"variant": v, "view": "status", "sfft": base["status"], "other": r["status"]})
            continue
        for k in INT_VIEWS:
            if r[k] != base[k]:
                flips.append({"variant": v, "view": k, "sfft": base[k], "other": r[k]})
        for k in ("onsets", "hmm"):
            if r[k] != base[k]:
                a, b = np.asarray(base[k]), np.asarray(r[k])
                n_diff = int(np.sum(a != b)) if a.shape == b.shape else -1
                flips.append({"variant": v, "view": k, "sfft_len": len(base[k]), "other_len": len(r[k]), "entries_differing": n_diff})
        for k in FLOATS:
            a, b = base[k], r[k]
            rel = abs(a - b) / max(abs(a), abs(b), 1e-6) if max(abs(a), abs(b)) > 1e-6 else abs(a - b)
            worst = max(worst, rel)
            if abs(a - b) > 1e-3 * max(abs(a), abs(b)) + 1e-6:
                flips.append({"variant": v, "view": k, "sfft": a, "other": b})
    return flips, worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--suite", default="full", choices=["full", "quick"])
    ap.add_argument("--jobs", type=int, default=max(1, mp.cpu_count()))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import oracle_lib as O

    O.build_oracle()
    items = workload(args.suite)
    # longest first so the pool drains evenly (C4 is 20x a C2 track)
    order = sorted(range(len(items)), key=lambda i: -(items[i][2] if isinstance(items[i][2], int) and items[i][2] > 1000 else 8_000_000))
    t0 = time.time()
    with mp.get_context("fork").Pool(args.jobs) as pool:
        recs = pool.map(analyse, [items[i] for i in order], chunksize=1)
    recs = [r for _, r in sorted(zip(order, recs))]
    table, n_ident, worst_all, ties = [], 0, 0.0, []
    for rec in recs:
        flips, worst = compare(rec)
        worst_all = max(worst_all, worst)
        n_ident += not flips
        if rec["runs"][0]["key_hashmap_tie"]:
            ties.append(rec["label"])
        b = rec["runs"][0]
        table.append({"track": rec["label"], "seconds": round(rec["seconds"], 1), "sr": rec["sr"], "bpm": b["bpm"], "key": b["key"], "n_onsets": len(b["onsets"]),
                      "n_hmm_frames": len(b["hmm"]), "escalated": b["multi_res_triggered"], "worst_float_rel_diff": worst, "flips": flips})
    doc = {
        "what": "whole-analysis outputs under three FFT arithmetics (oracle/so_fft.cpp variants); flips = outputs that differ from the SFFT run",
        "suite": args.suite, "variants": {str(k): v for k, v in VARIANTS.items()}, "tracks": len(recs), "tracks_identical_discrete_and_within_1e-3": n_ident,
        "worst_float_rel_diff": worst_all, "tracks_with_flips": [t["track"] for t in table if t["flips"]],
        "key_label_from_hashmap_order_vote": {"count": len(ties), "tracks": ties,
                                              "note": "tracks whose returned key comes from detect_key_weighted's own vote with two keys tied at the top "
                                                      "(detector.rs:254-275: HashMap iteration order in the reference, first-ranked = Major here); with segment "
                                                      "voting on (default) the label is taken from accumulated scores instead"},
        "wall_seconds": round(time.time() - t0, 1), "table": table,
    }
    out = Path(args.out) if args.out else ROOT / "tests" / "golden" / f"fft_robustness_{args.suite}.json"
    out.write_text(json.dumps(doc, indent=1) + "\n")
    print(json.dumps({k: v for k, v in doc.items() if k != "table"}, indent=1))
    for t in table:
        if t["flips"]:
            print(t["track"], json.dumps(t["flips"]))


if __name__ == "__main__":
    main()
