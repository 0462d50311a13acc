"""Writes RIFF/WAVE files in every layout the device-side decoder takes (tests only)."""
import struct

import numpy as np


def encode(x: np.ndarray, fmt: str) -> bytes:
    """x: float64/float32 in [-1, 1), shape (frames,) or (frames, channels) -> interleaved little-endian sample bytes."""
    v = np.asarray(x, np.float64).reshape(-1)
    if fmt == "u8":
        return np.clip(np.round(v * 128.0 + 128.0), 0, 255).astype(np.uint8).tobytes()
    if fmt == "s16":
        return np.clip(np.round(v * 32767.0), -32768, 32767).astype("<i2").tobytes()
    if fmt == "s24":
        i = np.clip(np.round(v * 8388607.0), -8388608, 8388607).astype(np.int32)
        b = np.zeros((i.size, 3), np.uint8)
        b[:, 0], b[:, 1], b[:, 2] = i & 0xFF, (i >> 8) & 0xFF, (i >> 16) & 0xFF
        return b.tobytes()
    if fmt == "s32":
        return np.clip(np.round(v * 2147483647.0), -2147483648, 2147483647).astype("<i4").tobytes()
    if fmt == "f32":
        return v.astype("<f4").tobytes()
    if fmt == "f64":
        return v.astype("<f8").tobytes()
    raise ValueError(fmt)


BITS = {"u8": 8, "s16": 16, "s24": 24, "s32": 32, "f32": 32, "f64": 64}


def wav_bytes(x: np.ndarray, sr: int, fmt: str, extensible: bool = False, extra_chunk: bool = False) -> bytes:
    ch = 1 if x.ndim == 1 else x.shape[1]
    data = encode(x, fmt)
    tag = 3 if fmt in ("f32", "f64") else 1
    bits = BITS[fmt]
    align = bits // 8 * ch
    if extensible:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        fmt_chunk = struct.pack("<HHIIHHHHI", 0xFFFE, ch, sr, sr * align, align, bits, 22, bits, (1 << ch) - 1) + struct.pack("<H", tag) + guid_tail
    else:
        fmt_chunk = struct.pack("<HHIIHH", tag, ch, sr, sr * align, align, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt_chunk)) + fmt_chunk
    if extra_chunk:  # an odd-sized LIST chunk in front of the data: readers must skip it including its pad byte
        chunks += b"LIST" + struct.pack("<I", 5) + b"INFOx" + b"\x00"
    chunks += b"data" + struct.pack("<I", len(data)) + data + (b"\x00" if len(data) & 1 else b"")
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def decode_reference(sample_bytes: bytes, fmt: str, ch: int) -> np.ndarray:
    """The reference decoder's arithmetic (examples/analyze_batch.rs:70-165) in numpy: per-sample conversion to f32, channels
    summed left to right in f32 from 0.0, divided by the channel count."""
    if fmt == "u8":
        v = (np.frombuffer(sample_bytes, np.uint8).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    elif fmt == "s16":
        v = np.frombuffer(sample_bytes, "<i2").astype(np.float32) / np.float32(32768.0)
    elif fmt == "s24":
        b = np.frombuffer(sample_bytes, np.uint8).reshape(-1, 3).astype(np.int32)
        i = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        i = np.where(i >= 1 << 23, i - (1 << 24), i)
        v = i.astype(np.float32) / np.float32(8388608.0)
    elif fmt == "s32":
        v = np.frombuffer(sample_bytes, "<i4").astype(np.float32) / np.float32(2147483648.0)
    elif fmt == "f32":
        v = np.frombuffer(sample_bytes, "<f4").astype(np.float32)
    else:
        v = np.frombuffer(sample_bytes, "<f8").astype(np.float32)
    if ch == 1:
        return np.ascontiguousarray(v)
    v = v.reshape(-1, ch)
    acc = np.zeros(v.shape[0], np.float32)
    for c in range(ch):
        acc = acc + v[:, c]
    return (acc / np.float32(ch)).astype(np.float32)
