"""GPU: frame-streaming compatibility mode.  A processor written against the reference's frame
interface (Audio_stream::try_pop / try_push / set_eof) runs between device-resident nodes: the stream it pops is
the upstream node's buffer cut into the recorded frame sizes, the frames it pushes become one device buffer whose
frame runs are the pushed sizes.  Its arithmetic is the reference's change_volume<T> on the host, so the result
must equal the oracle gain bit for bit."""
import numpy as np
import pytest

from helpers import FMT_FLT, FMT_FLTP, FMT_S16, FMT_S16P, FMT_S32, assert_bit_equal, make_input

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", [FMT_S16, FMT_S32, FMT_FLT, FMT_S16P, FMT_FLTP])
@pytest.mark.parametrize("frame_size", [1152, 1000])
def test_frame_node_between_device_nodes(eng_gpu, orc, fmt, frame_size):
    eng_gpu.register_examples()
    n = 44100 + 321
    x = make_input(orc, fmt, n, 2)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    g0 = p.add("audio_volume_adjust")
    fg = p.add("frame_gain_example", {"volume": 0.5})
    g1 = p.add("audio_volume_adjust")
    out = p.add("audio_output")
    p.link(src, "output_0", g0, "input"); p.link(g0, "output", fg, "input")
    p.link(fg, "output", g1, "input"); p.link(g1, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_volume(g0, 0.9); e.set_volume(g1, 0.8)
    e.bind_source(0, x, fmt, 44100, frame_size=frame_size)
    e.run()
    a = orc.gain(x, fmt, 0.9); b = orc.gain(a, fmt, 0.5); c = orc.gain(b, fmt, 0.8)
    mid = e.product(fg, "output")
    assert (mid.fmt, mid.rate, mid.ch, mid.frames) == (fmt, 44100, 2, n)
    assert_bit_equal(mid.numpy(), b, "frame node output")
    assert_bit_equal(e.output().numpy(), c, "after the frame node")
    # the frame sizes travel through the frame node unchanged
    want = [(frame_size, n // frame_size)] + ([(n % frame_size, 1)] if n % frame_size else [])
    assert e.product_runs(fg, "output") == want


def test_frame_node_fan_out_and_resampled_input(eng_gpu, orc):
    """frames popped from an audio_amix product (FLTP planes, amix's own frame sizes) and pushed to two links"""
    eng_gpu.register_examples()
    x = make_input(orc, FMT_FLT, 30000, 2)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    mix = p.add("audio_amix", eng_gpu.amix_info([1.0]))
    fg = p.add("frame_gain_example", {"volume": 0.5})
    ga = p.add("audio_volume_adjust"); gb = p.add("audio_volume_adjust")
    out = p.add("audio_output")
    p.link(src, "output_0", mix, "input_1"); p.link(mix, "output", fg, "input")
    p.link(fg, "output", ga, "input"); p.link(fg, "output", gb, "input"); p.link(ga, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_volume(ga, 1.0); e.set_volume(gb, 0.25)
    e.bind_source(0, x, FMT_FLT, 44100)
    e.run()
    rl, rr = orc.amix([orc.make_track(x, FMT_FLT, 44100)], [1.0])
    ref = orc.gain(np.stack([rl, rr]), FMT_FLTP, 0.5)
    assert_bit_equal(e.output().numpy(), ref, "fan-out link a")
    assert_bit_equal(e.product(fg, "output").numpy(), ref, "frame node product (first link)")
    assert e.product_runs(fg, "output") == e.product_runs(mix, "output")
    # the frames travel with their own stamps: amix's end-time stamps (whole microseconds, App. C4) come out of the frame
    # node and of the gain behind it unchanged (`out_frame->pts = src_frame.pts`, audio-vol.cpp:170)
    first = int(1152 / 48000.0 * 1000000) * (1 / 1000000.0)
    assert e.product(mix, "output").pts == first and e.product(fg, "output").pts == first and e.output().pts == first
    assert e.product_stamp(mix, "output") == (eng_gpu.STAMP_END_US, 0.0)
    assert e.product_stamp(ga, "output")[0] == 3     # per-frame list kept from the pushed frames
