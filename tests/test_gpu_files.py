"""GPU: the PCM / WAV edges of the render.  audio_input reads canonical RIFF/WAVE files named by
the project's file_path (PCM 16 / 24 / 32 bit and IEEE float; the reference decodes with libavformat, whose wav demuxer
hands out 4096-byte packets: 1024 sample frames for 16-bit stereo, 512 for float stereo, 682 for 24-bit stereo -- CPU side in
tests/test_wav_probe.py); audio_output's export writes a float WAV and keeps do_export's pts rule: (int)((frame stamp - time) * sample_rate)
samples of silence in front of every frame where that is positive (src/processor/audio-io.cpp:833-839), with the stamps
the producing node would have put on its frames (tests/test_export_stamps.py has the arithmetic)."""
import struct
import wave

import numpy as np
import pytest

from helpers import FMT_FLT, FMT_FLTP, FMT_S16, FMT_S32, assert_bit_equal, make_input

pytestmark = pytest.mark.gpu


def _write_float_wav(path, x, rate):
    data = np.ascontiguousarray(x, np.float32).tobytes()
    ch = x.shape[1]
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, ch, rate, rate * 4 * ch, 4 * ch, 32))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def _read_float_wav(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt "
    tag, ch, rate = struct.unpack("<HHI", raw[20:28])
    assert tag == 3 and raw[36:40] == b"data"
    n = struct.unpack("<I", raw[40:44])[0]
    return np.frombuffer(raw[44:44 + n], np.float32).reshape(-1, ch), rate


def test_wav_sources_through_gain_and_float_wav_export(eng_gpu, orc, tmp_path):
    n = 30000
    s16 = make_input(orc, FMT_S16, n, 2, track=1)
    flt = make_input(orc, FMT_FLT, n + 500, 2, track=2)
    with wave.open(str(tmp_path / "a.wav"), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100); w.writeframes(s16.tobytes())
    _write_float_wav(str(tmp_path / "b.wav"), flt, 44100)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [str(tmp_path / "a.wav"), str(tmp_path / "b.wav")]})
    ga = p.add("audio_volume_adjust"); gb = p.add("audio_volume_adjust")
    mix = p.add("audio_amix", eng_gpu.amix_info([0.5, 0.5]))
    out = p.add("audio_output")
    p.link(src, "output_0", ga, "input"); p.link(src, "output_1", gb, "input")
    p.link(ga, "output", mix, "input_1"); p.link(gb, "output", mix, "input_2"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_volume(ga, 0.8); e.set_volume(gb, 0.6)
    e.set_export_path(str(tmp_path / "out.wav"))
    e.run()
    a = orc.gain(s16, FMT_S16, 0.8); b = orc.gain(flt, FMT_FLT, 0.6)
    assert_bit_equal(e.product(ga, "output").numpy(), a, "gain of the 16-bit WAV source")
    assert e.product_runs(ga, "output")[0][0] == 1024            # WAV packets: 4096 bytes of 16-bit stereo
    assert e.product_runs(gb, "output")[0][0] == 512             # ... and of float stereo
    rl, rr = orc.amix([orc.make_track(a, FMT_S16, 44100, frame_size=1024), orc.make_track(b, FMT_FLT, 44100, frame_size=512)], [0.5, 0.5])
    got = e.output()
    assert_bit_equal(got.numpy(), np.stack([rl, rr]), "mix of the two files")
    y, rate = _read_float_wav(str(tmp_path / "out.wav"))
    assert rate == 48000
    # amix's frames are as long as the shortest front frame of its live inputs (audio-amix.cpp:190-193): 512 here, the float
    # file's packets.  It stamps them with their END time (App. C4), truncated to whole microseconds (audio-amix.cpp:199-201):
    # 512 / 48000 s = 10666.67 us -> 10666 us, and do_export's (int)((frame_begin - 0) * 48000) = (int)511.97 -> the reference's
    # export starts with 511 samples of silence, not 512
    assert got.pts == 10666 * (1 / 1000000.0)
    lead = int(got.pts * 48000)
    assert lead == 511 and not y[:lead].any()
    assert e.product_stamp(mix, "output") == (eng_gpu.STAMP_END_US, 0.0)
    # ... and more silence wherever amix's frame size grows: after the 16-bit file's last frame (304 samples) the 512-sample
    # frames resume, and after the last (short) input frames come the flush frames of 1152 samples (audio-amix.cpp:195), whose end-time stamps run ahead of the export's `time` (audio-io.cpp:833-839)
    from test_export_stamps import frames_of, ref_export, ref_stamps_end_us
    sizes = frames_of(e.product_runs(mix, "output"))
    silence, _ = ref_export(ref_stamps_end_us(sizes, 48000), sizes, 48000)
    assert silence[0] == lead
    mixed = np.ascontiguousarray(np.stack([rl, rr]).T)
    parts, at = [], 0
    for nb, n0 in zip(sizes, silence):
        parts += [np.zeros((n0, 2), np.float32), mixed[at:at + nb]]
        at += nb
    assert at == mixed.shape[0]
    assert_bit_equal(y, np.concatenate(parts), "exported WAV")


def test_24_bit_wav_source_arrives_as_s32(eng_gpu, orc, tmp_path):
    """pcm_s24le decodes to AV_SAMPLE_FMT_S32 with the sample in the upper three bytes (libavcodec/pcm.c), in frames of
    4092 / 6 = 682 sample frames; the gain node then scales 32-bit integers (audio-vol.cpp:75-100)"""
    n = 5000
    rng = np.random.default_rng(24)
    s24 = rng.integers(-(1 << 23), 1 << 23, (n, 2), dtype=np.int64)
    raw = np.zeros((n, 2, 3), np.uint8)
    for k in range(3):
        raw[:, :, k] = (s24 >> (8 * k)) & 0xFF
    data = raw.tobytes()
    with open(str(tmp_path / "s24.wav"), "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 2, 48000, 48000 * 6, 6, 24))
        f.write(b"data" + struct.pack("<I", len(data)) + data)
    assert eng_gpu.probe_wav(str(tmp_path / "s24.wav")) == (FMT_S32, 48000, 2, n, 682)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [str(tmp_path / "s24.wav")]})
    g = p.add("audio_volume_adjust")
    out = p.add("audio_output")
    p.link(src, "output_0", g, "input"); p.link(g, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_volume(g, 0.37)
    e.run()
    s32 = (s24 << 8).astype(np.int32)
    got = e.output()
    assert (got.fmt, got.rate, got.ch, got.frames) == (FMT_S32, 48000, 2, n)
    assert_bit_equal(got.numpy(), orc.gain(s32, FMT_S32, 0.37), "gain of the 24-bit samples widened to 32 bits")
    assert frames_of_runs(e.product_runs(g, "output")) == [682] * 7 + [n - 7 * 682]


def frames_of_runs(runs):
    return [int(l) for l, c in runs for _ in range(int(c))]


def test_export_of_an_amix_whose_frame_size_grows(eng_gpu, orc, tmp_path):
    """audio_amix cuts its output into frames as long as the shortest live input frame (audio-amix.cpp:190-195); when a
    1024-sample input ends before an 1152-sample one the end-time stamps jump ahead of the export's `time`, and do_export
    encodes the difference as silence IN THE MIDDLE of the stream (audio-io.cpp:833-839) -- reproduced frame by frame."""
    from test_export_stamps import frames_of, ref_export, ref_stamps_end_us
    short = make_input(orc, FMT_FLT, 48000 // 3, 2, rate=48000, track=3)
    long_ = make_input(orc, FMT_FLT, 48000, 2, rate=48000, track=4)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    mix = p.add("audio_amix", eng_gpu.amix_info([0.7, 0.4]))
    out = p.add("audio_output")
    p.link(src, "output_0", mix, "input_1"); p.link(src, "output_1", mix, "input_2"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_export_path(str(tmp_path / "grow.wav"))
    e.bind_source(0, short, FMT_FLT, 48000, frame_size=1024)
    e.bind_source(1, long_, FMT_FLT, 48000, frame_size=1152)
    e.run()
    rl, rr = orc.amix([orc.make_track(short, FMT_FLT, 48000, frame_size=1024), orc.make_track(long_, FMT_FLT, 48000, frame_size=1152)], [0.7, 0.4])
    mixed = np.ascontiguousarray(np.stack([rl, rr]).T)
    assert_bit_equal(e.output().numpy(), np.stack([rl, rr]), "mix")
    runs = e.product_runs(mix, "output")
    sizes = frames_of(runs)
    assert sizes[0] == 1024 and 1152 in sizes
    silence, _ = ref_export(ref_stamps_end_us(sizes, 48000), sizes, 48000)
    assert silence[0] == 1023 and sum(1 for n in silence if n) >= 2, "lead-in and at least the growth step"
    assert eng_gpu.export_plan(eng_gpu.STAMP_END_US, 0.0, 48000, runs)[0] == silence
    parts, at = [], 0
    for nb, n in zip(sizes, silence):
        parts.append(np.zeros((n, 2), np.float32))
        parts.append(mixed[at:at + nb])
        at += nb
    y, rate = _read_float_wav(str(tmp_path / "grow.wav"))
    assert rate == 48000
    assert_bit_equal(y, np.concatenate(parts), "exported WAV with the reference's silence")


def test_start_time_stamps_switch_exports_the_mix_and_nothing_else(eng_gpu, orc, tmp_path):
    """SURVEY.md App. C4 switch (JSON key "start_time_stamps" on audio_amix / audio_bimix): frames stamped with exact start
    times from 0 instead of the reference's truncated end times -- same samples, no silence in the export"""
    short = make_input(orc, FMT_FLT, 48000 // 3, 2, rate=48000, track=3)
    long_ = make_input(orc, FMT_FLT, 48000, 2, rate=48000, track=4)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    mix = p.add("audio_amix", dict(eng_gpu.amix_info([0.7, 0.4]), start_time_stamps=True))
    out = p.add("audio_output")
    p.link(src, "output_0", mix, "input_1"); p.link(src, "output_1", mix, "input_2"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_export_path(str(tmp_path / "plain.wav"))
    e.bind_source(0, short, FMT_FLT, 48000, frame_size=1024)
    e.bind_source(1, long_, FMT_FLT, 48000, frame_size=1152)
    e.run()
    rl, rr = orc.amix([orc.make_track(short, FMT_FLT, 48000, frame_size=1024), orc.make_track(long_, FMT_FLT, 48000, frame_size=1152)], [0.7, 0.4])
    got = e.output()
    assert got.pts == 0.0 and e.product_stamp(mix, "output") == (eng_gpu.STAMP_START, 0.0)
    assert_bit_equal(got.numpy(), np.stack([rl, rr]), "mix")
    y, rate = _read_float_wav(str(tmp_path / "plain.wav"))
    assert rate == 48000
    assert_bit_equal(y, np.ascontiguousarray(np.stack([rl, rr]).T), "exported WAV without any silence")


def test_export_pads_a_late_stream_with_silence(eng_gpu, orc, tmp_path):
    x = make_input(orc, FMT_FLT, 5000, 2, rate=48000)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    g = p.add("audio_volume_adjust")
    out = p.add("audio_output")
    p.link(src, "output_0", g, "input"); p.link(g, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_export_path(str(tmp_path / "late.wav"))
    e.bind_source(0, x, FMT_FLT, 48000, pts=0.25)
    e.run()
    y, _ = _read_float_wav(str(tmp_path / "late.wav"))
    assert y.shape[0] == 12000 + 5000 and not y[:12000].any()
    assert_bit_equal(y[12000:], x, "late stream after its silence")


def test_missing_and_malformed_files_are_runtime_errors(eng_gpu, tmp_path):
    (tmp_path / "bad.wav").write_bytes(b"RIFF....WAVEjunk")
    # a path that names no regular file fails with the reference's words (audio-io.cpp:232-239); a file that is there but is
    # not a WAV this engine reads fails with the engine's own
    for path, text in ((str(tmp_path / "nope.wav"), "Invalid file path in slot 1"), (str(tmp_path), "Invalid file path in slot 1"),
                       (str(tmp_path / "bad.wav"), "Cannot open audio file")):
        p = eng_gpu.Project()
        src = p.add("audio_input", {"file_path": [path]})
        out = p.add("audio_output")
        p.link(src, "output_0", out, "input")
        e = eng_gpu.Engine(p.json())
        with pytest.raises(eng_gpu.EngineError) as x:
            e.run()
        assert text in x.value.message


def test_nodey_render_cli(eng_gpu, orc, tmp_path):
    """the headless renderer: project file + WAV sources in, float WAV out, one JSON line on stdout"""
    import json, os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "nodey-audio-editor_b200", "nodey_render")
    assert os.path.exists(exe), "nodey_render is missing: run __graft_entry__.build()"
    x = make_input(orc, FMT_S16, 48000, 2, rate=48000, track=7)
    with wave.open(str(tmp_path / "in.wav"), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000); w.writeframes(x.tobytes())
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [str(tmp_path / "in.wav")]})
    pm = p.add("pitch_modifier", {"pitch": -2.0})
    out = p.add("audio_output")
    p.link(src, "output_0", pm, "input"); p.link(pm, "output", out, "input")
    (tmp_path / "project.json").write_text(json.dumps(p.json()))
    r = subprocess.run([exe, str(tmp_path / "project.json"), str(tmp_path / "out.wav")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["nodes"] == 3 and info["links"] == 2 and info["gpu_launches"] > 0
    y, rate = _read_float_wav(str(tmp_path / "out.wav"))
    ref, _, _ = orc.soundtouch(orc.extract_interleaved(x, FMT_S16), 48000, 1.0, orc.pitch_node_factor(-2.0), 1152)
    assert rate == 48000 and abs(info["audio_seconds"] - ref.shape[0] / 48000.0) < 1e-9
    assert_bit_equal(y, ref, "nodey_render output")
    # a project that cannot be read is an error exit with a message, not a crash
    (tmp_path / "broken.json").write_text("{ not json")
    r = subprocess.run([exe, str(tmp_path / "broken.json")], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "Invalid File" in r.stderr
