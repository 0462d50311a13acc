#!/usr/bin/env python3
"""analyze_batch — the reference's batch CLI (examples/analyze_batch.rs:196-405) on the B200 path.

    python examples/analyze_batch.py [--devices 0,1,...] [--json] <file1.wav> <file2.wav> ...

Same output as the reference: one JSON object per line with --json (same keys, same number formatting,
errors as {"file":..,"error":..}), otherwise the "[i/n] path: BPM=.. Key=.." lines; summary on stderr.
The reference decodes with symphonia; this tool reads the RIFF/WAVE container itself (PCM 8/16/24/32 bit, IEEE float 32/64 bit,
WAVE_FORMAT_EXTENSIBLE, any channel count) and hands the sample bytes to the device undecoded: the per-format conversion and the
mono mixdown of examples/analyze_batch.rs:70-165 run on the GPU (compressed formats need a decoder library on the caller's side).  `--jobs` is accepted and ignored: parallelism is across the
tracks of the batch on the device(s).
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stratum_dsp_b200 as S  # noqa: E402


def _display(e) -> str:
    """Display of AnalysisError (src/error.rs:24-34)."""
    names = {"InvalidInput": "Invalid input", "DecodingError": "Decoding error", "ProcessingError": "Processing error", "NotImplemented": "Not implemented",
             "NumericalError": "Numerical error"}
    return f"{names.get(e.kind, e.kind)}: {e.message}"


def percentile(xs, p):
    xs = sorted(xs)
    idx = int(round((len(xs) - 1) * min(max(p, 0.0), 1.0)))
    return xs[min(idx, len(xs) - 1)]


def opt(v):
    return "null" if v is None else ("true" if v else "false")


def main(argv):
    as_json, devices, paths = False, None, []
    it = iter(argv)
    for a in it:
        if a == "--json":
            as_json = True
        elif a == "--jobs":
            next(it)
        elif a == "--devices":
            devices = [int(d) for d in next(it).split(",")]
        elif a in ("--help", "-h"):
            print(__doc__, file=sys.stderr)
            return 0
        else:
            paths.append(a)
    if not paths:
        print("ERROR: Provide at least one audio file path. Use --help for usage.", file=sys.stderr)
        return 2
    print(f"Batch: {len(paths)} files, devices={devices or 'current'}", file=sys.stderr)
    t0 = time.perf_counter()
    tracks, idx, errors = [], [], {}
    for i, p in enumerate(paths):
        try:
            tracks.append(S.read_wav(p))
            idx.append(i)
        except Exception as e:  # decode failure -> ItemOut.error, the batch goes on
            errors[i] = str(e)
    results = dict(zip(idx, S.analyze_batch_pcm(tracks, devices=devices))) if tracks else {}
    times = []
    for i, p in enumerate(paths):
        r = results.get(i)
        err = ("decode failed: " + errors[i]) if i in errors else (("analysis failed: " + _display(r.error)) if r is not None and r.error is not None else None)
        if err is None:
            m = r.metadata
            c = S.compute_confidence(r)  # the reference prints compute_confidence's values (analyze_batch.rs:271-279)
            times.append(m.processing_time_ms)
            if as_json:
                print("{" + f'"file":{json.dumps(p)},"bpm":{r.bpm:.2f},"bpm_confidence":{c.bpm_confidence:.4f},"key":{json.dumps(r.key.name())},'
                      f'"key_confidence":{c.key_confidence:.4f},"processing_time_ms":{m.processing_time_ms:.2f},'
                      f'"tempogram_multi_res_triggered":{opt(m.tempogram_multi_res_triggered)},"tempogram_multi_res_used":{opt(m.tempogram_multi_res_used)},'
                      f'"tempogram_percussive_triggered":{opt(m.tempogram_percussive_triggered)},"tempogram_percussive_used":{opt(m.tempogram_percussive_used)}' + "}")
            else:
                print(f"[{i + 1}/{len(paths)}] {p}: BPM={r.bpm:.2f} (conf={c.bpm_confidence:.3f}) Key={r.key.name()} (conf={c.key_confidence:.3f}) "
                      f"time={m.processing_time_ms:.2f}ms")
        elif as_json:
            print("{" + f'"file":{json.dumps(p)},"error":{json.dumps(err)}' + "}")
        else:
            print(f"[{i + 1}/{len(paths)}] {p}: ERROR: {err}")
    wall_ms = (time.perf_counter() - t0) * 1000.0
    print(f"Done: ok={len(times)}/{len(paths)} wall={wall_ms:.0f}ms", file=sys.stderr)
    if times:
        print(f"processing_time_ms: mean={sum(times) / len(times):.2f} p50={percentile(times, 0.5):.2f} p90={percentile(times, 0.9):.2f} "
              f"min={min(times):.2f} max={max(times):.2f}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
