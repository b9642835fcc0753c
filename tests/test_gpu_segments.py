"""GPU: time-segment sharding through the C ABI: the ranks' segments, each computed from its
input slice alone, concatenate to the whole-stream result bit for bit; nodey_resampler_segment agrees with the
host-side rule in bindings/segments.py."""
import numpy as np
import pytest

from helpers import FMT_FLT, FMT_S16, assert_bit_equal, make_input, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("fmt", [FMT_FLT, FMT_S16])
def test_resample_segments_bit_exact(nd, orc, world, fmt):
    import segments
    n = 200003
    x = make_input(orc, fmt, n, 2)
    rs = nd.Resampler(44100, 48000)
    xd = to_dev(x)
    whole = rs.run(xd, fmt, flush=True).cpu().numpy()
    total = whole.shape[1]
    parts = []
    for rank in range(world):
        seg, out = segments.run_resample_segment(rs, xd, fmt, n, total, world, rank)
        c = rs.segment(n, seg["k0"], seg["k1"])
        assert (c["in0"], c["in1"], c["skip"], c["flush"]) == (seg["in0"], seg["in1"], seg["skip"], seg["flush"])
        parts.append(out.cpu().numpy())
    got = np.concatenate(parts, axis=1)
    assert_bit_equal(got, whole, "resample segments")


def test_resample_segment_abi_errors(nd):
    rs = nd.Resampler(44100, 48000)
    with pytest.raises(nd.NodeyError):
        rs.segment(100000, 100, 5000)                 # not a multiple of the phase count
    with pytest.raises(nd.NodeyError):
        nd.Resampler(44099, 48000).segment(100000, 0, 5000)


@pytest.mark.parametrize("world", [2, 4])
def test_stft_segments_bit_exact(nd, orc, world):
    import segments
    import torch
    n = 48000 * 3 + 11
    x = to_dev(np.ascontiguousarray(orc.synth_f32(n, 2, 48000, 9).T))
    whole = nd.stft(x, False)
    parts = [segments.run_stft_segment(nd, x, world, r)[1] for r in range(world)]
    got = torch.cat([p for p in parts if p is not None], dim=1)
    assert torch.equal(got.view(torch.float32), whole.view(torch.float32)), "stft segments differ from the whole stream"
