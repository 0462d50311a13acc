"""Deterministic synthetic inputs for the parity tests and the bench.

The same generator is implemented on the device (stratum_b200_synth_batch) so the bench can build
1024 distinct tracks without a 32 GB host buffer; tests/test_gpu_parity.py checks the two agree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

C4_HZ = 261.6255653005986  # 440 * 2**(-9/12)
CLICK_LEN_S = 0.005
CLICK_HZ = 1000.0
CLICK_TAU_S = 0.001
CLICK_AMP = 0.8


class PCG32:
    """PCG-XSH-RR 64/32 (O'Neill 2014), single stream."""

    MULT = 6364136223846793005
    INC = 1442695040888963407
    MASK = (1 << 64) - 1

    def __init__(self, seed: int):
        self.state = 0
        self.next_u32()
        self.state = (self.state + seed) & self.MASK
        self.next_u32()

    def next_u32(self) -> int:
        old = self.state
        self.state = (old * self.MULT + self.INC) & self.MASK
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def bounded(self, n: int) -> int:
        # plain modulo (bias < 2^-25 for the tiny ranges used here); kept simple so the device copy is trivial
        return self.next_u32() % n

    def unit(self) -> float:
        return self.next_u32() / 4294967296.0


@dataclass
class TrackParams:
    bpm: float
    tonic: int
    minor: int
    phase_frac: float  # click phase as a fraction of the beat period
    chord_amp: float
    sample_rate: int = 44100
    n_samples: int = 7_938_000

    def chord_freqs(self):
        f0 = C4_HZ * 2.0 ** (self.tonic / 12.0)
        third = 3 if self.minor else 4
        return (f0, f0 * 2.0 ** (third / 12.0), f0 * 2.0 ** (7 / 12.0))


def c1_params(n_samples: int = 7_938_000, sr: int = 44100) -> TrackParams:
    """C1: 128 BPM click + C-major triad (0.1 each), phase 0."""
    return TrackParams(128.0, 0, 0, 0.0, 0.1, sr, n_samples)


def c2_params(i: int, n_samples: int = 7_938_000, sr: int = 44100) -> TrackParams:
    """C2/C3: track i drawn from PCG32(0x5EED0000 + i)."""
    g = PCG32(0x5EED0000 + i)
    bpm = float(70 + g.bounded(111))
    tonic = g.bounded(12)
    minor = g.bounded(2)
    phase = g.unit()
    amp = 0.05 + 0.1 * g.unit()
    return TrackParams(bpm, tonic, minor, phase, amp, sr, n_samples)


def c5_params(i: int) -> TrackParams:
    """C5: ragged — duration log-uniform in [30 s, 600 s], sr alternates 44.1/48 kHz."""
    g = PCG32(0xC5C50000 + i)
    p = c2_params(i)
    dur = 30.0 * math.exp(g.unit() * math.log(600.0 / 30.0))
    sr = 44100 if i % 2 == 0 else 48000
    p.sample_rate = sr
    p.n_samples = int(dur * sr)
    return p


def render(p: TrackParams, bpm_end: float | None = None) -> np.ndarray:
    """Render one track to float32.  bpm_end: linear tempo drift (C4) from p.bpm to bpm_end."""
    n, sr = p.n_samples, p.sample_rate
    t = np.arange(n, dtype=np.float64) / sr
    x = np.zeros(n, dtype=np.float64)
    for f in p.chord_freqs():
        x += p.chord_amp * np.sin(2.0 * np.pi * f * t)
    clen = int(CLICK_LEN_S * sr)
    ct = np.arange(clen, dtype=np.float64) / sr
    click = CLICK_AMP * np.sin(2.0 * np.pi * CLICK_HZ * ct) * np.exp(-ct / CLICK_TAU_S)
    for s in click_starts(p, bpm_end):
        e = min(s + clen, n)
        x[s:e] += click[: e - s]
    return x.astype(np.float32)


def click_starts(p: TrackParams, bpm_end: float | None = None):
    n, sr = p.n_samples, p.sample_rate
    out = []
    if bpm_end is None:
        period = 60.0 / p.bpm * sr
        k = 0
        while True:
            s = int(round((p.phase_frac + k) * period))
            if s >= n:
                break
            out.append(s)
            k += 1
    else:
        # beat k happens when the integral of bpm(t)/60 reaches k: b0 t + (b1-b0) t^2 / (2 T) = 60 k
        T = n / sr
        a = (bpm_end - p.bpm) / (2.0 * T)
        k = 0
        while True:
            if abs(a) < 1e-12:
                tk = 60.0 * k / p.bpm
            else:
                tk = (-p.bpm + math.sqrt(p.bpm * p.bpm + 4.0 * a * 60.0 * k)) / (2.0 * a)
            s = int(round(tk * sr))
            if s >= n:
                break
            out.append(s)
            k += 1
    return out


def c4_mix(n_samples: int = 158_760_000, sr: int = 44100) -> np.ndarray:
    """C4: 60-min mix, tempo 120 -> 128 BPM, chord root up a fifth every 5 minutes."""
    p = TrackParams(120.0, 0, 0, 0.0, 0.1, sr, n_samples)
    x = np.zeros(n_samples, dtype=np.float32)
    seg = 5 * 60 * sr
    tonic = 0
    for s0 in range(0, n_samples, seg):
        s1 = min(s0 + seg, n_samples)
        q = TrackParams(120.0, tonic, 0, 0.0, 0.1, sr, n_samples)
        t = np.arange(s0, s1, dtype=np.float64) / sr
        acc = np.zeros(s1 - s0, dtype=np.float64)
        for f in q.chord_freqs():
            acc += 0.1 * np.sin(2.0 * np.pi * f * t)
        x[s0:s1] = acc.astype(np.float32)
        tonic = (tonic + 7) % 12
    clen = int(CLICK_LEN_S * sr)
    ct = np.arange(clen, dtype=np.float64) / sr
    click = (CLICK_AMP * np.sin(2.0 * np.pi * CLICK_HZ * ct) * np.exp(-ct / CLICK_TAU_S)).astype(np.float32)
    for s in click_starts(p, 128.0):
        e = min(s + clen, n_samples)
        x[s:e] += click[: e - s]
    return x


# ---- re-creations of the reference's four WAV fixtures (scripts/generate_fixtures.py describes the
# signals; soundfile writes PCM_16, hound reads i16/32768 — tests/integration_tests.rs:7-33) ----------
def _pcm16_roundtrip(x: np.ndarray) -> np.ndarray:
    q = np.clip(np.round(x.astype(np.float64) * 32768.0), -32768, 32767)  # libsndfile float->short scaling
    return (q / 32768.0).astype(np.float32)


def fixture_kick(bpm: float, duration: float, sr: int = 44100) -> np.ndarray:
    n = int(duration * sr)
    x = np.zeros(n, dtype=np.float32)
    klen = int(0.1 * sr)
    t = np.arange(klen) / sr
    kick = (np.sin(2 * np.pi * 60 * t) * 0.6 + np.sin(2 * np.pi * 120 * t) * 0.3 + np.sin(2 * np.pi * 180 * t) * 0.1) * np.exp(-t * 10)
    for bt in np.arange(0, duration, 60.0 / bpm):
        s = int(bt * sr)
        e = min(s + klen, n)
        x[s:e] += kick[: e - s].astype(np.float32)
    x = x / np.max(np.abs(x))
    return _pcm16_roundtrip(x)


def fixture_cmajor_scale(sr: int = 44100) -> np.ndarray:
    freqs = [261.63, 293.66, 329.63, 349.23, 392.00, 440.00, 493.88, 523.25]
    parts = []
    for f in freqs:
        n = int(0.5 * sr)
        t = np.arange(n) / sr
        note = np.sin(2 * np.pi * f * t)
        fade = int(0.05 * sr)
        env = np.ones(n)
        env[:fade] = np.linspace(0, 1, fade)
        env[-fade:] = np.linspace(1, 0, fade)
        parts.append(note * env)
    x = np.concatenate(parts).astype(np.float32)
    x = x / np.max(np.abs(x))
    return _pcm16_roundtrip(x)


def fixture_mixed_silence(sr: int = 44100) -> np.ndarray:
    sil = np.zeros(5 * sr, dtype=np.float32)
    t = np.arange(5 * sr) / sr
    tone = (np.sin(2 * np.pi * 440.0 * t).astype(np.float32)) * 0.5
    return _pcm16_roundtrip(np.concatenate([sil, tone, sil]))


# ---- richer tonal material for the optional key-path variants ---------------------------------------
# A chord progression with harmonics, a bass line, a scale melody, clicks and a little noise: template scores no longer
# tie between the modes, tuning offsets are measurable (detune_cents), and the bass band carries its own pitch content.
def render_progression(seed: int, duration_s: float, sr: int = 44100, tonic: int = 0, minor: bool = False, bpm: float = 120.0,
                       detune_cents: float = 0.0, noise: float = 0.003) -> np.ndarray:
    rng = np.random.RandomState(seed)
    n = int(duration_s * sr)
    t = np.arange(n, dtype=np.float64) / sr
    x = np.zeros(n, dtype=np.float64)
    scale = [0, 2, 3, 5, 7, 8, 10] if minor else [0, 2, 4, 5, 7, 9, 11]
    degrees = [0, 5, 3, 4] if not minor else [0, 5, 2, 6]  # I-vi-IV-V / i-VI-III-VII
    tune = 2.0 ** (detune_cents / 1200.0)
    beat = 60.0 / bpm
    bar = 4 * beat
    nbars = int(math.ceil(duration_s / bar))

    def hz(semi_from_c4):
        return C4_HZ * tune * 2.0 ** (semi_from_c4 / 12.0)

    for b in range(nbars):
        s0, s1 = int(b * bar * sr), min(int((b + 1) * bar * sr), n)
        if s0 >= n:
            break
        d = degrees[b % 4]
        triad = [scale[d % 7] + 12 * (d // 7), scale[(d + 2) % 7] + 12 * ((d + 2) // 7), scale[(d + 4) % 7] + 12 * ((d + 4) // 7)]
        tt = t[s0:s1]
        env = np.minimum(1.0, (tt - tt[0]) / 0.02) * np.minimum(1.0, (tt[-1] - tt) / 0.02 + 1e-3)
        for semi in triad:
            f = hz(tonic + semi)
            for h in range(1, 5):
                x[s0:s1] += env * 0.06 / h * np.sin(2.0 * np.pi * f * h * tt + 0.3 * h)
        fb = hz(tonic + triad[0] - 24)  # bass: chord root two octaves down
        for h in range(1, 4):
            x[s0:s1] += env * 0.12 / h * np.sin(2.0 * np.pi * fb * h * tt)
        for q in range(8):  # eighth-note melody from the scale, one octave up
            m0, m1 = s0 + int(q * beat / 2 * sr), min(s0 + int((q + 1) * beat / 2 * sr), n)
            if m0 >= n:
                break
            deg = rng.randint(0, 7)
            f = hz(tonic + scale[deg] + 12)
            mt = t[m0:m1] - t[m0]
            x[m0:m1] += 0.05 * np.exp(-mt / 0.15) * np.sin(2.0 * np.pi * f * mt)
    clen = int(CLICK_LEN_S * sr)
    ct = np.arange(clen, dtype=np.float64) / sr
    click = 0.5 * np.sin(2.0 * np.pi * CLICK_HZ * ct) * np.exp(-ct / CLICK_TAU_S)
    k = 0
    while True:
        s = int(round(k * beat * sr))
        if s >= n:
            break
        e = min(s + clen, n)
        x[s:e] += click[: e - s]
        k += 1
    x += noise * rng.standard_normal(n)
    return x.astype(np.float32)
