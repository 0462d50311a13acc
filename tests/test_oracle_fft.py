"""The oracle's FFT (oracle/so_fft.cpp, the documented stand-in for rustfft 6.2) against a float64 DFT."""
import numpy as np
import pytest

import oracle_lib as O


@pytest.mark.parametrize("m", [2, 4, 8, 16, 64, 512, 1024, 2048, 4096, 8192, 32768])
def test_cfft_matches_float64(m):
    rng = np.random.default_rng(m)
    z = (rng.standard_normal(m) + 1j * rng.standard_normal(m)).astype(np.complex64)
    got = O.cfft(z)
    ref = np.fft.fft(z.astype(np.complex128))
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6 * max(1.0, np.log2(m) / 4)


@pytest.mark.parametrize("n", [4, 8, 2048, 8192, 16384])
def test_rfft_matches_float64(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    got = O.rfft(x)
    ref = np.fft.rfft(x.astype(np.float64))
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-6


def test_stft_shape_and_sine_bin():
    # chroma/extractor.rs:1524-1561: a 440 Hz sine peaks at bin round(440 / (sr / N))
    sr, n = 44100, 44100
    x = (0.5 * np.sin(2 * np.pi * 440.0 * np.arange(n) / sr)).astype(np.float32)
    S = O.stft(x, 2048, 512)
    assert S.shape == ((n - 2048) // 512 + 1, 1025)
    assert int(S[10].argmax()) == round(440.0 / (sr / 2048))
    assert O.stft(x[:1000], 2048, 512).shape == (0, 1025)
