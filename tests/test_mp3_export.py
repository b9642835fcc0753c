"""CPU: the MP3 leg of the audio_output sink (host/src/mp3-export.cpp) against the reference's LAME call sequence
(Audio_output::do_export, src/processor/audio-io.cpp:640-841).  LAME is not in the image, so a recording test double
(tests/fake_lame/fake_lame.c, same entry points) stands in: the log shows the calls and their arguments, the "MP3"
file holds one 16-byte record per encode call (kind, samples, checksum of what the entry point would read)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import FMT_FLT, FMT_FLTP, FMT_S16, FMT_S16P, FMT_S32, FMT_S32P, make_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIND = {1: "encode_buffer_interleaved", 2: "encode_buffer", 3: "encode_buffer_interleaved_int", 4: "encode_buffer_int",
        5: "encode_buffer_interleaved_ieee_float", 6: "encode_buffer_ieee_float"}


@pytest.fixture(scope="session")
def fake_lame(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fake_lame") / "libfakelame.so")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "fake_lame", "fake_lame.c"), "-lm"], check=True)
    return so


@pytest.fixture
def lame(fake_lame, tmp_path, monkeypatch):
    log = str(tmp_path / "lame.log")
    monkeypatch.setenv("NODEY_LAME_LIB", fake_lame)
    monkeypatch.setenv("FAKE_LAME_LOG", log)
    monkeypatch.delenv("FAKE_LAME_FAIL", raising=False)
    return lambda: open(log).read().split("\n")[:-1]


def records(path):
    raw = open(path, "rb").read()
    assert len(raw) % 16 == 0
    out = []
    for k in range(0, len(raw), 16):
        magic, kind, n, s = struct.unpack("<4siif", raw[k:k + 16])
        assert magic == b"FLAM"
        out.append((kind, n, s))
    return out


def scale(x):
    return np.abs(x.astype(np.float64)) / (65536.0 if x.dtype == np.int32 else 1.0)


def test_parameters_silence_and_frames_of_a_late_float_stream(orc, lame, tmp_path):
    import engine
    assert engine.mp3_available()
    x = make_input(orc, FMT_FLT, 5000, 2, rate=44100)
    out = str(tmp_path / "late.mp3")
    end = engine.encode_mp3(out, x, FMT_FLT, 44100, frame_size=1152, pts=0.25, kbps=192)
    log = lame()
    # audio-io.cpp:805-822: in rate / channels / quality 2 / STEREO = 0 / out 48000 / vbr_off = 0 / bit rate, then init_params
    assert log[:9] == ["init", "set_in_samplerate 44100", "set_num_channels 2", "set_quality 2", "set_mode 0",
                       "set_out_samplerate 48000", "set_VBR 0", "set_brate 192", "init_params"]
    lead = int(0.25 * 44100)
    sizes = [1152] * 4 + [5000 - 4 * 1152]
    want = [f"encode_buffer_interleaved n={lead} buf={int(1.25 * lead + 7200)}"]
    want += [f"encode_buffer_interleaved_ieee_float n={n} buf={4 * n + 7200}" for n in sizes]
    assert log[9:] == want + ["close"]                 # no lame_encode_flush: the reference never calls it
    rec = records(out)
    assert [(k, n) for k, n, _ in rec] == [(1, lead)] + [(5, n) for n in sizes]
    assert rec[0][2] == 0.0
    at = 0
    for (_, n, s) in rec[1:]:
        assert s == np.float32(scale(x[at:at + n]).sum())
        at += n
    assert end == pytest.approx(0.25 + 5000 / 44100, abs=1e-12)


@pytest.mark.parametrize("fmt,nch,kind", [(FMT_S16, 2, 1), (FMT_S16, 1, 2), (FMT_S16P, 2, 2), (FMT_S16P, 1, 2),
                                          (FMT_S32, 2, 3), (FMT_S32, 1, 4), (FMT_S32P, 2, 4),
                                          (FMT_FLT, 2, 5), (FMT_FLT, 1, 6), (FMT_FLTP, 2, 6), (FMT_FLTP, 1, 6)])
def test_entry_point_per_format(orc, lame, tmp_path, fmt, nch, kind):
    """packed stereo -> the interleaved entry point of its type (audio-io.cpp:704-766); planar and mono frames -> the
    two-pointer entry point (mono hands its plane as both channels; no fall-through: DESIGN.md 7)"""
    import engine
    n = 3000
    x = make_input(orc, fmt, n, nch, rate=48000, track=3)
    out = str(tmp_path / "f.mp3")
    engine.encode_mp3(out, x, fmt, 48000, frame_size=1024, kbps=320)
    log = lame()
    assert f"set_num_channels {nch}" in log and f"set_mode {0 if nch == 2 else 3}" in log and "set_brate 320" in log
    rec = records(out)
    assert [(k, m) for k, m, _ in rec] == [(kind, 1024), (kind, 1024), (kind, n - 2048)]     # pts 0: no silence call
    at = 0
    for (_, m, s) in rec:
        part = x[:, at:at + m] if fmt >= 5 else x[at:at + m]
        assert s == np.float32(scale(part).sum()), (fmt, nch, at)
        at += m
    assert [l for l in log if l.startswith("encode")] == [f"{KIND[kind]} n={m} buf={4 * m + 7200}" for _, m, _ in rec]


def test_time_carries_over_and_early_frames_get_no_silence(orc, lame, tmp_path):
    import engine
    x = make_input(orc, FMT_FLT, 2304, 2, rate=48000)
    # `time` already past the stream's start (audio-io.cpp:829-831: negative silence is skipped)
    end = engine.encode_mp3(str(tmp_path / "a.mp3"), x, FMT_FLT, 48000, pts=0.01, time=0.5)
    assert [k for k, _, _ in records(str(tmp_path / "a.mp3"))] == [5, 5]
    assert end == pytest.approx(0.01 + 1152 / 48000 + 1152 / 48000, abs=1e-12)
    # amix-style END-time stamp (App. C4): one frame of silence in front
    engine.encode_mp3(str(tmp_path / "b.mp3"), x, FMT_FLT, 48000, pts=1152 / 48000)
    assert [(k, n) for k, n, _ in records(str(tmp_path / "b.mp3"))] == [(1, 1152), (5, 1152), (5, 1152)]


def test_errors_like_the_reference(orc, lame, tmp_path, monkeypatch):
    import engine
    x = make_input(orc, FMT_FLT, 2000, 2, rate=48000)
    with pytest.raises(engine.EngineError, match="Failed to open output file"):
        engine.encode_mp3(str(tmp_path / "no_such_dir" / "x.mp3"), x, FMT_FLT, 48000)
    monkeypatch.setenv("FAKE_LAME_FAIL", "init_params")
    with pytest.raises(engine.EngineError, match="Failed to initialize LAME parameters"):
        engine.encode_mp3(str(tmp_path / "x.mp3"), x, FMT_FLT, 48000)
    monkeypatch.setenv("FAKE_LAME_FAIL", "encode")
    with pytest.raises(engine.EngineError, match=r"Failed to encode audio frame.*LAME Error: -3"):
        engine.encode_mp3(str(tmp_path / "x.mp3"), x, FMT_FLT, 48000)
    with pytest.raises(engine.EngineError, match="Failed to encode silence"):
        engine.encode_mp3(str(tmp_path / "x.mp3"), x, FMT_FLT, 48000, pts=0.1)
    monkeypatch.delenv("FAKE_LAME_FAIL")
    with pytest.raises(engine.EngineError, match="Unsupported sample format"):
        engine.encode_mp3(str(tmp_path / "x.mp3"), np.zeros((100, 2), np.float64), 4, 48000)


def test_missing_encoder_is_a_loud_error(orc, tmp_path, monkeypatch):
    import engine
    monkeypatch.setenv("NODEY_LAME_LIB", str(tmp_path / "libmp3lame-not-here.so"))
    assert not engine.mp3_available()
    x = make_input(orc, FMT_FLT, 2000, 2, rate=48000)
    with pytest.raises(engine.EngineError, match="MP3 encoder not available"):
        engine.encode_mp3(str(tmp_path / "x.mp3"), x, FMT_FLT, 48000)
    assert not os.path.exists(str(tmp_path / "x.mp3"))
