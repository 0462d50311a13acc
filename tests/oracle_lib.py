"""ctypes front-end for the CPU oracle (oracle/ — test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "-j4"], check=True)


def _load(fast: bool = False) -> C.CDLL:
    name = "libstratum_oracle_fast.so" if fast else "libstratum_oracle.so"
    p = ORACLE_DIR / "_build" / name
    if not p.exists():
        build_oracle()
    lib = C.CDLL(str(p))
    vp, cp, f32p, i64p = C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)
    lib.so_config_new.restype = vp
    lib.so_config_free.argtypes = [vp]
    lib.so_config_set.argtypes = [vp, cp, C.c_double]
    lib.so_config_get.argtypes = [vp, cp]
    lib.so_config_get.restype = C.c_double
    lib.so_analyze.argtypes = [f32p, C.c_uint64, C.c_uint32, vp, C.c_int]
    lib.so_analyze.restype = vp
    lib.so_free.argtypes = [vp]
    lib.so_status.argtypes = [vp]
    lib.so_error_message.argtypes = [vp, cp, C.c_int]
    lib.so_scalar.argtypes = [vp, cp]
    lib.so_scalar.restype = C.c_double
    lib.so_confidence.argtypes = [vp, f32p, C.POINTER(C.c_uint32)]
    lib.so_confidence_of.argtypes = [C.c_float] * 5 + [C.c_uint32, C.c_uint32, f32p, C.POINTER(C.c_uint32)]
    lib.so_warning_strings.argtypes = [vp, cp, C.c_int]
    lib.so_farray_len.argtypes = [vp, cp]
    lib.so_farray_len.restype = C.c_int64
    lib.so_farray_copy.argtypes = [vp, cp, f32p, C.c_int64]
    lib.so_farray_copy.restype = C.c_int64
    lib.so_iarray_len.argtypes = [vp, cp]
    lib.so_iarray_len.restype = C.c_int64
    lib.so_iarray_copy.argtypes = [vp, cp, i64p, C.c_int64]
    lib.so_iarray_copy.restype = C.c_int64
    lib.so_array_names.argtypes = [vp, cp, C.c_int]
    lib.so_set_fft_variant.argtypes = [C.c_int]
    lib.so_cfft.argtypes = [f32p, C.c_uint64]
    lib.so_rfft.argtypes = [f32p, C.c_uint64, f32p]
    lib.so_stft.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64, f32p, C.c_uint64]
    lib.so_stft.restype = C.c_int64
    lib.so_preprocess.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_int, f32p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.so_lufs.argtypes = [f32p, C.c_uint64, C.c_float, f32p]
    lib.so_vote_onsets.argtypes = [i64p, C.c_int, i64p, C.c_int, i64p, C.c_int, i64p, C.c_int, f32p, C.c_uint32, C.c_uint32, i64p, f32p,
                                   C.POINTER(C.c_uint32), C.c_int]
    lib.so_detect_key.argtypes = [f32p, C.c_uint64, f32p, C.POINTER(C.c_int), f32p, f32p, f32p, C.POINTER(C.c_int)]
    lib.so_key_clarity.argtypes = [f32p, C.c_int]
    lib.so_key_clarity.restype = C.c_float
    lib.so_key_templates.argtypes = [f32p, f32p]
    lib.so_key_name.argtypes = [C.c_int, C.c_uint32, C.c_int, cp, C.c_int]
    lib.so_hmm.argtypes = [C.c_float, f32p, C.c_int, C.POINTER(C.c_int32), f32p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    lib.so_beat_grid.argtypes = [C.c_float, C.c_float, f32p, C.c_int, C.c_uint32, f32p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.so_hpss_decompose.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64, f32p, f32p]
    lib.so_hpss_onsets.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_float, i64p, C.c_int]
    lib.so_batch_timed.argtypes = [f32p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, f32p, C.POINTER(C.c_int),
                                   C.POINTER(C.c_double)]
    lib.so_batch_timed.restype = C.c_double
    return lib


_LIBS: dict[bool, C.CDLL] = {}


def lib(fast: bool = False) -> C.CDLL:
    if fast not in _LIBS:
        _LIBS[fast] = _load(fast)
    return _LIBS[fast]


def f32ptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def i64ptr(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int64))


SCALARS = ["bpm", "bpm_confidence", "key_is_minor", "key_index", "key_confidence", "key_clarity", "grid_stability", "duration_seconds",
           "sample_rate", "onset_method_consensus", "warnings", "flags", "multi_res_triggered", "multi_res_used", "percussive_triggered",
           "percussive_used", "trim_start", "trim_end", "time_sig_beats_per_bar", "beats_refined", "key_hashmap_tie"]


class OracleResult:
    def __init__(self, L, h):
        self._L, self._h = L, h
        self.status = L.so_status(h)
        buf = C.create_string_buffer(512)
        L.so_error_message(h, buf, 512)
        self.error = buf.value.decode()
        for s in SCALARS:
            setattr(self, s, L.so_scalar(h, s.encode()))
        self.beats = self.farray("result.beats")
        self.downbeats = self.farray("result.downbeats")
        self.bars = self.farray("result.bars")
        self.onsets = self.iarray("result.onsets")
        self.hmm_beat_frames = self.iarray("result.hmm_beat_frames")
        out = (C.c_float * 4)()
        fl = C.c_uint32()
        L.so_confidence(h, out, C.byref(fl))
        self.confidence = dict(bpm=out[0], key=out[1], grid=out[2], overall=out[3], flags=fl.value)
        wb = C.create_string_buffer(2048)
        L.so_warning_strings(h, wb, 2048)
        self.warning_strings = [w for w in wb.value.decode().split("\n") if w]

    @property
    def key(self) -> int:
        return int(self.key_is_minor) * 12 + int(self.key_index)

    def names(self):
        buf = C.create_string_buffer(1 << 16)
        self._L.so_array_names(self._h, buf, 1 << 16)
        return [x for x in buf.value.decode().split("\n") if x]

    def farray(self, name: str):
        n = self._L.so_farray_len(self._h, name.encode())
        if n < 0:
            return None
        a = np.zeros(n, dtype=np.float32)
        if n:
            self._L.so_farray_copy(self._h, name.encode(), f32ptr(a), n)
        return a

    def iarray(self, name: str):
        n = self._L.so_iarray_len(self._h, name.encode())
        if n < 0:
            return None
        a = np.zeros(n, dtype=np.int64)
        if n:
            self._L.so_iarray_copy(self._h, name.encode(), i64ptr(a), n)
        return a

    def __del__(self):
        try:
            self._L.so_free(self._h)
        except Exception:
            pass


def analyze(samples: np.ndarray, sr: int, cfg: dict | None = None, dump: bool = False, fast: bool = False) -> OracleResult:
    L = lib(fast)
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    c = None
    if cfg:
        c = L.so_config_new()
        for k, v in cfg.items():
            if L.so_config_set(c, k.encode(), float(v)) != 0:
                raise KeyError(k)
    h = L.so_analyze(f32ptr(samples) if samples.size else None, samples.size, sr, c, 1 if dump else 0)
    if c:
        L.so_config_free(c)
    return OracleResult(L, h)


def stft(samples: np.ndarray, frame: int, hop: int) -> np.ndarray:
    L = lib()
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    if samples.size < frame:
        return np.zeros((0, frame // 2 + 1), np.float32)
    nf = (samples.size - frame) // hop + 1
    out = np.zeros((nf, frame // 2 + 1), dtype=np.float32)
    got = L.so_stft(f32ptr(samples), samples.size, frame, hop, f32ptr(out), out.size)
    assert got == nf
    return out


def rfft(x: np.ndarray) -> np.ndarray:
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((x.size // 2 + 1, 2), dtype=np.float32)
    L.so_rfft(f32ptr(x), x.size, f32ptr(out))
    return out[:, 0] + 1j * out[:, 1]


def cfft(z: np.ndarray) -> np.ndarray:
    L = lib()
    buf = np.zeros((z.size, 2), dtype=np.float32)
    buf[:, 0] = z.real
    buf[:, 1] = z.imag
    L.so_cfft(f32ptr(buf), z.size)
    return buf[:, 0] + 1j * buf[:, 1]
