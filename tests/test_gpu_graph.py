"""GPU parity at graph level: the config-5 render graph (resample -> pitch -> tempo -> gain -> amix
tree -> spectrum) through the C ABI against the same graph composed from oracle nodes.  One GPU
sums the tracks in the reference's order, so the master bus is compared bit for bit; the
spectrum to the float tolerance (1e-5 of the frame peak)."""
import numpy as np
import pytest

from helpers import assert_bit_equal, to_dev

pytestmark = pytest.mark.gpu


def test_config5_graph_small(nd, orc):
    import pipeline
    from oracle import graph_oracle as G
    n = 44100 * 3 + 123
    tracks = [orc.synth_f32(n, 2, 44100, t) for t in range(32)]
    keep_ref = {}
    G.track_chain(tracks[0], G.track_gain(0), keep=keep_ref)
    ref_bus, ref_spec = G.render(tracks, threads=8)

    r = pipeline.Config5Renderer(n, sub_batch=32)
    x = to_dev(np.stack(tracks))
    keep = {}
    buses = r.render_groups(x, 0, keep=keep)
    assert_bit_equal(keep["amix1"].cpu().numpy(), keep_ref["amix1"], "audio_amix(1) product")
    assert_bit_equal(keep["pitch"].cpu().numpy(), keep_ref["pitch"], "pitch_modifier product")
    assert_bit_equal(keep["tempo"].cpu().numpy(), keep_ref["tempo"], "velocity_modifier product")
    bus = r.master(buses)
    assert_bit_equal(bus.cpu().numpy(), ref_bus, "master bus")
    spec = r.spectrum(bus).cpu().numpy()
    peak = np.abs(ref_spec).max(axis=-1, keepdims=True)
    assert (np.abs(spec - ref_spec) <= 1e-5 * np.maximum(peak, 1e-30)).all()
