"""The sink's per-frame silence rule (Audio_output::do_export, src/processor/audio-io.cpp:826-839) over the three ways the
reference's nodes stamp their frames, as plain arithmetic on the CPU (include/nodey_engine.h: nodey_engine_export_plan)
against a line-by-line Python restatement of the reference's statements (doubles are IEEE on both sides):

  audio_amix / audio_bimix   time_seconds += nb / double(rate); pts = time_seconds * 1000000   (audio-amix.cpp:199-201)
  SoundTouch nodes           pts = (int64_t)(float)(time_seconds * 1000000); time_seconds += n / rate
                             (audio-velocity.cpp:238-249, 313-318)
  decoder stamps             exact start times
  do_export                  frame_begin = pts * av_q2d({1, 1000000}); silence = (int)((frame_begin - time) * rate);
                             time = frame_begin + nb / (double)rate
"""
import numpy as np
import pytest

import engine


def frames_of(runs):
    return [nb for nb, count in runs for _ in range(count)]


def ref_stamps_end_us(sizes, rate):
    t, out = 0.0, []
    for nb in sizes:
        t += nb / float(rate)
        out.append(int(t * 1000000))
    return out


def ref_stamps_float_us(sizes, rate, origin):
    t, out = origin, []
    for nb in sizes:
        out.append(int(np.float32(t * 1000000)))
        t += float(nb) / rate
    return out


def ref_export(stamps_us, sizes, rate, time=0.0):
    silence = []
    for pts, nb in zip(stamps_us, sizes):
        frame_begin = pts * (1 / 1000000.0)
        frame_end = frame_begin + nb / float(rate)
        n = int((frame_begin - time) * rate)
        silence.append(n if n > 0 else 0)
        time = frame_end
    return silence, time


@pytest.mark.parametrize("nb,lead", [(1024, 1023), (1152, 1152), (4096, 4095), (441, 440), (1, 0)])
def test_end_time_stamps_put_almost_one_frame_of_silence_in_front(nb, lead):
    # App. C4: the stamp is the frame's END, truncated to microseconds -- 1024 / 48000 s = 21333.33 us -> 21333 us -> 1023.98 samples
    runs = [(nb, 200)]
    sil, pts, end = engine.export_plan(engine.STAMP_END_US, 0.0, 48000, runs)
    want, want_end = ref_export(ref_stamps_end_us(frames_of(runs), 48000), frames_of(runs), 48000)
    assert sil == want and end == want_end
    assert sil[0] == lead and not any(sil[1:])


def test_end_time_stamps_add_silence_where_the_frame_size_grows():
    # audio_amix frames are as long as its shortest live input frame: when a 1024-sample input ends before an 1152-sample one
    # the stamps jump by the larger size while `time` advanced by the smaller one
    runs = [(1024, 37), (1152, 50), (576, 3), (1152, 2), (100, 1)]
    sizes = frames_of(runs)
    sil, pts, end = engine.export_plan(engine.STAMP_END_US, 0.0, 48000, runs)
    stamps = ref_stamps_end_us(sizes, 48000)
    want, want_end = ref_export(stamps, sizes, 48000)
    assert sil == want and end == want_end
    assert pts == [s * (1 / 1000000.0) for s in stamps]
    assert sil[37] in (127, 128) and sil[37 + 50 + 3] in (575, 576) and sum(1 for s in sil if s) == 3


@pytest.mark.parametrize("rate,nb,origin", [(44100, 1152, 0.0), (48000, 921, 0.021333), (44100, 2765, 0.0)])
def test_float_microsecond_stamps_of_a_long_stream(rate, nb, origin):
    # App. C8: a float holds microseconds in steps of 32 us after 268 s and 64 us after 537 s -- more than a sample: the
    # reference's export of a long SoundTouch stream inserts single samples of silence now and then
    count = int(700 * rate / nb)
    runs = [(nb, count)]
    sizes = frames_of(runs)
    sil, pts, end = engine.export_plan(engine.STAMP_START_FLOAT_US, origin, rate, runs)
    stamps = ref_stamps_float_us(sizes, rate, origin)
    want, want_end = ref_export(stamps, sizes, rate)
    assert sil == want and end == want_end
    assert pts == [s * (1 / 1000000.0) for s in stamps]
    if nb == 1152:      # 26122.45 us per frame against 64 us steps: some stamps land more than a sample late
        assert sum(sil[1:]) > 0, "the reference's quirk should show within 700 s"


def test_exact_start_stamps_and_a_carried_time():
    runs = [(1024, 10), (300, 1)]
    sil, pts, end = engine.export_plan(engine.STAMP_START, 0.25, 44100, runs)
    assert sil == [int(0.25 * 44100)] + [0] * 10
    assert pts[3] == 0.25 + 3 * 1024 / 44100.0
    # `time` already past the first frames (audio-io.cpp:829-831: negative silence is skipped)
    sil2, _, end2 = engine.export_plan(engine.STAMP_START, 0.25, 44100, runs, time=1.0)
    assert not any(sil2) and end2 == end
    # a stream cut short by `frames`
    sil3, pts3, _ = engine.export_plan(engine.STAMP_START, 0.0, 44100, runs, frames=2500)
    assert len(sil3) == 3


def test_seeded_random_runs_against_the_restated_rule():
    rng = np.random.default_rng(7)
    for _ in range(60):
        rate = int(rng.choice([8000, 22050, 44100, 48000, 96000]))
        runs = [(int(rng.integers(1, 5000)), int(rng.integers(1, 40))) for _ in range(int(rng.integers(1, 8)))]
        sizes = frames_of(runs)
        time = float(rng.choice([0.0, 0.013, 2.5]))
        origin = float(rng.choice([0.0, 0.024, 311.7]))
        sil, _, end = engine.export_plan(engine.STAMP_END_US, 0.0, rate, runs, time=time)
        assert (sil, end) == ref_export(ref_stamps_end_us(sizes, rate), sizes, rate, time)
        sil, _, end = engine.export_plan(engine.STAMP_START_FLOAT_US, origin, rate, runs, time=time)
        assert (sil, end) == ref_export(ref_stamps_float_us(sizes, rate, origin), sizes, rate, time)
