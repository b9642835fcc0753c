"""GPU parity: audio_spectrum (N2) against the oracle's double-precision DFT.  Float DSP bar of
BASELINE.json: 1e-5 relative (taken against the frame's spectral peak) / -100 dBFS residual."""
import numpy as np
import pytest

from helpers import to_dev

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


def _check(got, ref):
    assert got.shape == ref.shape
    peak = np.abs(ref).max(axis=-1, keepdims=True)
    err = np.abs(got.astype(np.complex128) - ref.astype(np.complex128))
    assert (err <= REL_TOL * np.maximum(peak, 1e-30)).all(), float((err / np.maximum(peak, 1e-30)).max())


@pytest.mark.parametrize("n", [4096, 4097, 5119, 5120, 48000])
@pytest.mark.parametrize("interleaved", [False, True])
def test_stft_matches_oracle(nd, orc, n, interleaved):
    x = orc.synth_f32(n, 2, 48000, 2)
    ref = np.stack([orc.stft(x[:, c].copy()) for c in range(2)])
    d = to_dev(x) if interleaved else to_dev(x.T.copy())
    got = nd.stft(d, interleaved).cpu().numpy()
    _check(got, ref)


def test_stft_short_and_errors(nd, orc):
    assert nd.stft_frames(4095) == 0
    x = to_dev(orc.synth_f32(1000, 1, 48000, 0).T.copy())
    assert nd.stft(x, False).shape == (1, 0, 2049)
    with pytest.raises(nd.NodeyError) as e:
        nd.stft(x, False, nfft=2048)
    assert e.value.code == -5


def test_stft_impulse_and_tone(nd, orc):
    n = 4096 * 3
    x = np.zeros((1, n), np.float32)
    x[0, 2048] = 1.0
    got = nd.stft(to_dev(x), False).cpu().numpy()[0]
    w = orc.hann(4096)
    k = np.arange(2049)
    for m in range(got.shape[0]):
        pos = 2048 - 1024 * m
        exp = (w[pos] * np.exp(-2j * np.pi * k * pos / 4096)) if 0 <= pos < 4096 else np.zeros(2049)
        assert np.abs(got[m] - exp).max() <= 1e-5
    # linearity: stft(a + b) == stft(a) + stft(b) within float rounding
    a = orc.synth_f32(n, 1, 48000, 1).T.copy(); b = orc.synth_f32(n, 1, 48000, 9).T.copy()
    sa = nd.stft(to_dev(a), False).cpu().numpy(); sb = nd.stft(to_dev(b), False).cpu().numpy()
    sab = nd.stft(to_dev((a + b).astype(np.float32)), False).cpu().numpy()
    assert np.abs(sab - (sa + sb)).max() <= 1e-5 * np.abs(sab).max()
