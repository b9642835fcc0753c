"""GPU: MP3 export through the audio_output sink.  The rendered stream goes to LAME in its own
sample format and frame sizes, as the reference's do_export would have fed it (src/processor/audio-io.cpp:640-841).
LAME is not in the image: the recording test double of tests/test_mp3_export.py stands in, so what is checked is the
call sequence and the payload of every call against the oracle render."""
import json
import os
import subprocess
import wave

import numpy as np
import pytest

from helpers import FMT_FLT, FMT_S16, assert_bit_equal, make_input
from test_mp3_export import fake_lame, lame, records, scale  # noqa: F401  (fixtures + helpers)

pytestmark = pytest.mark.gpu


def close(s, want):
    """the double's checksum is a sequential double sum rounded to float; numpy sums pairwise"""
    return abs(float(s) - float(want)) <= 1e-6 * max(1.0, abs(float(want)))


def test_gain_stream_exported_as_mp3(eng_gpu, orc, lame, tmp_path):
    x = make_input(orc, FMT_FLT, 5000, 2, rate=48000)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    g = p.add("audio_volume_adjust", {"volume": 0.5})
    out = p.add("audio_output")
    p.link(src, "output_0", g, "input"); p.link(g, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_export_path(str(tmp_path / "late.mp3"))
    e.set_export_kbps(128)
    e.bind_source(0, x, FMT_FLT, 48000, pts=0.25)
    e.run()
    ref = orc.gain(x, FMT_FLT, 0.5)
    assert_bit_equal(e.output().numpy(), ref, "sink keeps the rendered stream")
    log = lame()
    assert "set_in_samplerate 48000" in log and "set_brate 128" in log and "set_quality 2" in log and log[-1] == "close"
    rec = records(str(tmp_path / "late.mp3"))
    sizes = [1152] * 4 + [5000 - 4 * 1152]
    assert [(k, n) for k, n, _ in rec] == [(1, 12000)] + [(5, n) for n in sizes]
    at = 0
    for (_, n, s) in rec[1:]:
        assert close(s, scale(ref[at:at + n]).sum())
        at += n


def test_amix_output_goes_to_the_planar_entry_point_after_one_frame_of_silence(eng_gpu, orc, lame, tmp_path):
    n = 20000
    a = make_input(orc, FMT_S16, n, 2, track=1)
    b = make_input(orc, FMT_FLT, n, 2, track=2)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    mix = p.add("audio_amix", eng_gpu.amix_info([0.5, 0.5]))
    out = p.add("audio_output")
    p.link(src, "output_0", mix, "input_1"); p.link(src, "output_1", mix, "input_2"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_export_path(str(tmp_path / "mix.mp3"))
    e.bind_source(0, a, FMT_S16, 44100)
    e.bind_source(1, b, FMT_FLT, 44100)
    e.run()
    rl, rr = orc.amix([orc.make_track(a, FMT_S16, 44100), orc.make_track(b, FMT_FLT, 44100)], [0.5, 0.5])
    got = e.output()
    assert_bit_equal(got.numpy(), np.stack([rl, rr]), "mix")
    rec = records(str(tmp_path / "mix.mp3"))
    runs = e.product_runs(mix, "output")
    frames = [l for l, c in runs for _ in range(c)]
    # amix stamps frames with their END time (App. C4): the reference's export starts with one frame of silence, and encodes
    # more silence wherever amix's frame size grows (the last short frame is followed by 1152-sample flush frames whose
    # stamps run ahead of the export's `time`, audio-io.cpp:833-839) -- tests/test_export_stamps.py has the arithmetic
    from test_export_stamps import ref_export, ref_stamps_end_us
    silence, _ = ref_export(ref_stamps_end_us(frames, 48000), frames, 48000)
    assert rec[0][:2] == (1, int(got.pts * 48000)) and rec[0][1] == frames[0] == silence[0]
    want = []
    for m, quiet in zip(frames, silence):
        want += ([(1, quiet)] if quiet else []) + [(6, m)]
    assert [(k, m) for k, m, _ in rec] == want
    at = 0
    for (k, m, s) in rec:
        if k == 1:
            continue
        assert close(s, scale(rl[at:at + m]).sum() + scale(rr[at:at + m]).sum())
        at += m
    assert "set_brate 320" in lame()                     # the editor's default bit rate


def test_mp3_export_without_lame_fails_the_run(eng_gpu, orc, tmp_path, monkeypatch):
    monkeypatch.setenv("NODEY_LAME_LIB", str(tmp_path / "libmp3lame-not-here.so"))
    x = make_input(orc, FMT_FLT, 3000, 2, rate=48000)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    out = p.add("audio_output")
    p.link(src, "output_0", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_export_path(str(tmp_path / "x.mp3"))
    e.bind_source(0, x, FMT_FLT, 48000)
    with pytest.raises(eng_gpu.EngineError, match="MP3 encoder not available"):
        e.run()
    e.set_export_path(str(tmp_path / "x.wav"))            # the WAV export does not need the encoder
    e.run()
    assert os.path.getsize(str(tmp_path / "x.wav")) == 44 + 3000 * 8


def test_nodey_render_cli_mp3(eng_gpu, orc, fake_lame, tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "nodey-audio-editor_b200", "nodey_render")
    x = make_input(orc, FMT_S16, 10000, 2, rate=48000, track=5)
    with wave.open(str(tmp_path / "in.wav"), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000); w.writeframes(x.tobytes())
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [str(tmp_path / "in.wav")]})
    g = p.add("audio_volume_adjust", {"volume": 0.25})
    out = p.add("audio_output")
    p.link(src, "output_0", g, "input"); p.link(g, "output", out, "input")
    (tmp_path / "project.json").write_text(json.dumps(p.json()))
    env = dict(os.environ, NODEY_LAME_LIB=fake_lame, FAKE_LAME_LOG=str(tmp_path / "cli.log"))
    r = subprocess.run([exe, str(tmp_path / "project.json"), str(tmp_path / "out.mp3"), "--kbps", "96"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0, r.stderr
    log = open(str(tmp_path / "cli.log")).read().split("\n")
    assert "set_brate 96" in log and "set_num_channels 2" in log
    ref = orc.gain(x, FMT_S16, 0.25)
    rec = records(str(tmp_path / "out.mp3"))
    sizes = [1024] * 9 + [10000 - 9 * 1024]              # WAV PCM packets
    assert [(k, n) for k, n, _ in rec] == [(1, n) for n in sizes]      # 16-bit packed stereo: the interleaved short entry point
    at = 0
    for (_, n, s) in rec:
        assert close(s, scale(ref[at:at + n]).sum())
        at += n
