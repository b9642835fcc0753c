"""GPU: the sink's preview path (src/processor/audio-io.cpp:478-638 do_preview): per input frame
swr_convert to 48 kHz stereo float, clamp to [-1, 1], queue packed frames -- and no flush at the end.  Checked
against the oracle's streaming swr model fed frame by frame (values bit exact, chunk sizes equal)."""
import numpy as np
import pytest

from helpers import FMT_FLT, FMT_FLTP, FMT_S16, assert_bit_equal, make_input

pytestmark = pytest.mark.gpu


def _oracle_preview(orc, x, fmt, rate, nch, frame_size):
    planar = fmt >= 5
    n = x.shape[1] if planar else x.shape[0]
    swr = orc.Swr(rate, 48000, fmt, nch)
    parts, chunks = [], []
    for a in range(0, n, frame_size):
        fr = x[:, a:a + frame_size] if planar else x[a:a + frame_size]
        cap = int((fr.shape[1] if planar else fr.shape[0]) / rate * 48000 * 1.5) + 64
        l, r = swr.convert(np.ascontiguousarray(fr), cap)
        chunks.append(len(l))
        parts.append(np.stack([l, r], axis=1))
    y = np.concatenate(parts) if parts else np.zeros((0, 2), np.float32)
    y = np.where(y < -1.0, np.float32(-1.0), np.where(y > 1.0, np.float32(1.0), y)).astype(np.float32)
    return y, chunks


@pytest.mark.parametrize("case", [(FMT_FLT, 44100, 2, 1152), (FMT_S16, 44100, 1, 1024), (FMT_FLTP, 48000, 2, 1152), (FMT_FLT, 22050, 2, 4096)])
def test_preview_matches_streaming_swr_and_clamp(eng_gpu, orc, case):
    fmt, rate, nch, frame_size = case
    n = rate + 777
    x = make_input(orc, fmt, n, nch, rate=rate, track=5)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    g = p.add("audio_volume_adjust")
    out = p.add("audio_output")
    p.link(src, "output_0", g, "input"); p.link(g, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_volume(g, 3.0)                    # drives the float signal beyond [-1, 1]: the clamp has work to do
    e.set_preview(True)
    e.bind_source(0, x, fmt, rate, frame_size=frame_size)
    e.run()
    got, chunks = e.preview()
    ref, ref_chunks = _oracle_preview(orc, orc.gain(x, fmt, 3.0), fmt, rate, nch, frame_size)
    assert chunks == ref_chunks, "per-frame chunk sizes differ from swr_convert's return values"
    assert_bit_equal(got, ref, f"preview {case}")
    if fmt in (FMT_FLT, FMT_FLTP):
        assert np.abs(got).max() == 1.0 and (np.abs(ref) == 1.0).sum() > 100
    # the preview never flushes: the export of the same graph is longer when a rate change holds samples back
    e.set_preview(False)
    e.run()
    assert e.output().frames == n
