"""GPU parity through the reference-facing plugin API: project JSON -> Graph::deserialize ->
Runner::create_and_run -> processor nodes -> C ABI kernels, against the same graphs composed from
oracle nodes.  Integer/routing and all single-GPU float paths are compared bit for bit."""
import numpy as np
import pytest

from helpers import FMT_FLT, FMT_FLTP, FMT_S16, FMT_S16P, assert_bit_equal, make_input

pytestmark = pytest.mark.gpu


def test_config1_gain_amix(eng_gpu, orc):
    """configs[0] (reference-only nodes): input -> fan-out to two gains -> amix(2) -> output, S16 source"""
    n = 44100 * 2 + 321
    x = make_input(orc, FMT_S16, n, 2)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    g0 = p.add("audio_volume_adjust"); g1 = p.add("audio_volume_adjust")
    mix = p.add("audio_amix", eng_gpu.amix_info([0.5, 0.5]))
    out = p.add("audio_output")
    p.link(src, "output_0", g0, "input"); p.link(src, "output_0", g1, "input")
    p.link(g0, "output", mix, "input_1"); p.link(g1, "output", mix, "input_2"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_volume(g0, 0.8); e.set_volume(g1, 0.5)
    e.bind_source(0, x, FMT_S16, 44100)
    e.run()
    a = orc.gain(x, FMT_S16, 0.8); b = orc.gain(x, FMT_S16, 0.5)
    assert_bit_equal(e.product(g0, "output").numpy(), a, "gain product")
    rl, rr = orc.amix([orc.make_track(a, FMT_S16, 44100), orc.make_track(b, FMT_S16, 44100)], [0.5, 0.5])
    got = e.output()
    assert (got.fmt, got.rate, got.ch) == (FMT_FLTP, 48000, 2)
    assert_bit_equal(got.numpy(), np.stack([rl, rr]), "amix output")


@pytest.mark.parametrize("rates", [(44100, 44100), (48000, 22050), (96000, 44100)])
def test_amix_mixed_rates_and_gaps(eng_gpu, orc, rates):
    """per-input resampling incl. an input above 48 kHz, whose frames leave gaps (reference behaviour)"""
    xs = [make_input(orc, FMT_FLT, 20000 + 777 * i, 2, rate=r, track=i) for i, r in enumerate(rates)]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    mix = p.add("audio_amix", eng_gpu.amix_info([0.7, 0.3]))
    out = p.add("audio_output")
    p.link(src, "output_0", mix, "input_1"); p.link(src, "output_1", mix, "input_2"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    for i, (x, r) in enumerate(zip(xs, rates)):
        e.bind_source(i, x, FMT_FLT, r)
    e.run()
    rl, rr = orc.amix([orc.make_track(x, FMT_FLT, r) for x, r in zip(xs, rates)], [0.7, 0.3])
    assert_bit_equal(e.output().numpy(), np.stack([rl, rr]), f"amix {rates}")


@pytest.mark.parametrize("bias", [0.0, -0.4])
def test_bimix(eng_gpu, orc, bias):
    xl = make_input(orc, FMT_FLT, 30000, 2, track=1); xr = make_input(orc, FMT_S16, 26000, 1, track=2)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    bm = p.add("audio_bimix", {"bias": bias})
    out = p.add("audio_output")
    p.link(src, "output_0", bm, "input_l"); p.link(src, "output_1", bm, "input_r"); p.link(bm, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.bind_source(0, xl, FMT_FLT, 44100); e.bind_source(1, xr, FMT_S16, 44100)
    e.run()
    rl, rr = orc.bimix(orc.make_track(xl, FMT_FLT, 44100), orc.make_track(xr, FMT_S16, 44100), bias)
    assert_bit_equal(e.output().numpy(), np.stack([rl, rr]), "bimix")


@pytest.mark.parametrize("pts", [(0.0, 0.0), (0.0, 0.25), (0.1, 0.0)])
def test_bimix_v2_alignment(eng_gpu, orc, pts):
    xl = make_input(orc, FMT_FLT, 40000, 2, track=3); xr = make_input(orc, FMT_FLT, 36000, 2, track=4)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    bm = p.add("audio_bimix_v2")
    out = p.add("audio_output")
    p.link(src, "output_0", bm, "input_l"); p.link(src, "output_1", bm, "input_r"); p.link(bm, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.bind_source(0, xl, FMT_FLT, 44100, pts=pts[0]); e.bind_source(1, xr, FMT_FLT, 44100, pts=pts[1])
    e.run()
    ref, ref_pts = orc.bimix_v2(orc.make_track(xl, FMT_FLT, 44100, pts0=pts[0]), orc.make_track(xr, FMT_FLT, 44100, pts0=pts[1]))
    got = e.output()
    assert (got.fmt, got.ch) == (FMT_FLT, 2)
    assert_bit_equal(got.numpy(), ref, f"bimix_v2 {pts}")
    assert got.pts == ref_pts


def test_split_gain_merge_variant(eng_gpu, orc):
    """configs[0], new-node variant: channel_split -> per-channel gain -> bimix_v2 merge"""
    n = 48000
    x = make_input(orc, FMT_S16, n, 2, rate=48000)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    sp = p.add("audio_channel_split")
    g0 = p.add("audio_volume_adjust"); g1 = p.add("audio_volume_adjust")
    bm = p.add("audio_bimix_v2")
    out = p.add("audio_output")
    p.link(src, "output_0", sp, "input")
    p.link(sp, "output_l", g0, "input"); p.link(sp, "output_r", g1, "input")
    p.link(g0, "output", bm, "input_l"); p.link(g1, "output", bm, "input_r"); p.link(bm, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.set_volume(g0, 0.8); e.set_volume(g1, 0.5)
    e.bind_source(0, x, FMT_S16, 48000)
    e.run()
    l, r = orc.split(x, FMT_S16)
    assert_bit_equal(e.product(sp, "output_l").numpy()[:, 0], l, "split L")
    gl = orc.gain(l.reshape(-1, 1), FMT_S16, 0.8); gr = orc.gain(r.reshape(-1, 1), FMT_S16, 0.5)
    ref, _ = orc.bimix_v2(orc.make_track(gl, FMT_S16, 48000), orc.make_track(gr, FMT_S16, 48000))
    assert_bit_equal(e.output().numpy(), ref, "split/gain/merge")


def test_config5_graph_through_engine(eng_gpu, orc):
    """configs[4] at small size, driven by project JSON: 32 tracks, device-resident sources"""
    from oracle import graph_oracle as G
    n = 44100 * 2 + 99
    tracks = [orc.synth_f32(n, 2, 44100, t) for t in range(32)]
    gains = [G.track_gain(t) for t in range(32)]
    p, ids = eng_gpu.config5_project(32, gains)
    e = eng_gpu.Engine(p.json())
    for t, x in enumerate(tracks):
        e.bind_source(t, x, FMT_FLT, 44100)
    e.run()
    ref_bus, ref_spec = G.render(tracks, threads=8)
    assert_bit_equal(e.output().numpy(), ref_bus, "master bus")
    spec = e.product(ids["spectrum"], "output").numpy()
    peak = np.abs(ref_spec).max(axis=-1, keepdims=True)
    assert (np.abs(spec - ref_spec) <= 1e-5 * np.maximum(peak, 1e-30)).all()
    # level-batched runner: 32 pitch nodes -> one batched SoundTouch launch set, not 32
    assert e.product_runs(ids["master"], "output")[0][0] == 1152


def test_node_errors_surface(eng_gpu, orc):
    p = eng_gpu.Project()
    g = p.add("audio_volume_adjust")
    out = p.add("audio_output")
    p.link(g, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    with pytest.raises(eng_gpu.EngineError) as x:
        e.run()
    assert x.value.code == eng_gpu.E_NODE and "has no input" in x.value.message
    # SoundTouch nodes refuse rates outside 8..48 kHz (audio-velocity.cpp:371)
    q = eng_gpu.Project()
    src = q.add("audio_input", {"file_path": [""]})
    pm = q.add("pitch_modifier", {"pitch": 2.0})
    o = q.add("audio_output")
    q.link(src, "output_0", pm, "input"); q.link(pm, "output", o, "input")
    e2 = eng_gpu.Engine(q.json())
    e2.bind_source(0, make_input(orc, FMT_FLT, 5000, 2, rate=96000), FMT_FLT, 96000)
    with pytest.raises(eng_gpu.EngineError) as x:
        e2.run()
    assert "Unsupported sample rate" in x.value.message


def test_amix_batch_of_mixed_nodes(eng_gpu, orc):
    """One graph level holding every kind of audio_amix the batch path must tell apart: three single-input
    resamplers of equal streams with different volumes (one launch), one of another length, one S16, one at
    22.05 kHz (320 phases: no pipelined kernel), one at 48 kHz (no resampling) and a two-input mixer -- each
    product must equal the oracle's amix of that node alone."""
    specs = [  # (format, rate, frames, volumes)
        (FMT_FLT, 44100, 30000, [1.0]), (FMT_FLT, 44100, 30000, [0.5]), (FMT_FLT, 44100, 30000, [0.25]),
        (FMT_FLT, 44100, 21111, [0.8]), (FMT_S16, 44100, 30000, [0.9]), (FMT_FLT, 22050, 15000, [0.7]),
        (FMT_FLT, 48000, 20000, [0.6]),
    ]
    xs = [make_input(orc, f, n, 2, rate=r, track=i) for i, (f, r, n, _) in enumerate(specs)]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""] * len(specs)})
    nodes = []
    for i, (_, _, _, vol) in enumerate(specs):
        m = p.add("audio_amix", eng_gpu.amix_info(vol))
        p.link(src, f"output_{i}", m, "input_1")
        p.link(m, "output", p.add("audio_volume_adjust"), "input")      # a product exists per link: give every node a consumer
        nodes.append(m)
    two = p.add("audio_amix", eng_gpu.amix_info([0.5, 0.5]))
    p.link(src, "output_0", two, "input_1"); p.link(src, "output_3", two, "input_2")
    out = p.add("audio_output")
    p.link(two, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    for i, (x, (f, r, _, _)) in enumerate(zip(xs, specs)):
        e.bind_source(i, x, f, r)
    e.run()
    for i, (m, (f, r, _, vol)) in enumerate(zip(nodes, specs)):
        rl, rr = orc.amix([orc.make_track(xs[i], f, r)], vol)
        got = e.product(m, "output")
        assert (got.fmt, got.rate, got.ch) == (FMT_FLTP, 48000, 2)
        assert_bit_equal(got.numpy(), np.stack([rl, rr]), f"amix node {i}")
    rl, rr = orc.amix([orc.make_track(xs[0], FMT_FLT, 44100), orc.make_track(xs[3], FMT_FLT, 44100)], [0.5, 0.5])
    assert_bit_equal(e.output().numpy(), np.stack([rl, rr]), "two-input amix in the same level")


def test_soundtouch_batch_of_mixed_formats(eng_gpu, orc):
    """pitch nodes of one level fed by FLT, FLTP (read in place) and S16 (converted) streams of equal length"""
    n = 48000 + 123
    fmts = [FMT_FLT, FMT_FLTP, FMT_S16, FMT_FLTP]
    xs = [make_input(orc, f, n, 2, rate=48000, track=10 + i) for i, f in enumerate(fmts)]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""] * len(fmts)})
    nodes = []
    for i in range(len(fmts)):
        pm = p.add("pitch_modifier", {"pitch": 3.0})
        p.link(src, f"output_{i}", pm, "input")
        p.link(pm, "output", p.add("audio_volume_adjust"), "input")
        nodes.append(pm)
    out = p.add("audio_output")
    p.link(nodes[0], "output", out, "input")
    e = eng_gpu.Engine(p.json())
    for i, (x, f) in enumerate(zip(xs, fmts)):
        e.bind_source(i, x, f, 48000)
    e.run()
    pitch = orc.pitch_node_factor(3.0)
    for i, (x, f) in enumerate(zip(xs, fmts)):
        xi = orc.extract_interleaved(x, f)
        ref, _, _ = orc.soundtouch(xi, 48000, 1.0, pitch, 1152)
        assert_bit_equal(e.product(nodes[i], "output").numpy(), ref, f"pitch node {i} (format {f})")


def test_tiny_and_ragged_streams_through_the_graph(eng_gpu, orc):
    """a 300-frame and a 5000-frame source (shorter than a WSOLA sequence, shorter than a resampler tile) through
    resample -> pitch -> gain -> two-input mix: every length and sample as the oracle's nodes"""
    xs = [make_input(orc, FMT_FLT, 300, 2, track=1), make_input(orc, FMT_S16, 5000, 2, track=2)]
    fmts = [FMT_FLT, FMT_S16]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    chain = []
    for i in range(2):
        rs = p.add("audio_amix", eng_gpu.amix_info([1.0]))
        pm = p.add("pitch_modifier", {"pitch": 3.0})
        g = p.add("audio_volume_adjust")
        p.link(src, f"output_{i}", rs, "input_1"); p.link(rs, "output", pm, "input"); p.link(pm, "output", g, "input")
        chain.append((rs, pm, g))
    mix = p.add("audio_amix", eng_gpu.amix_info([0.5, 0.5]))
    out = p.add("audio_output")
    p.link(chain[0][2], "output", mix, "input_1"); p.link(chain[1][2], "output", mix, "input_2"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    for i in range(2):
        e.set_volume(chain[i][2], 0.7)
        e.bind_source(i, xs[i], fmts[i], 44100)
    e.run()
    pitch = orc.pitch_node_factor(3.0)
    tracks = []
    for i in range(2):
        rl, rr = orc.amix([orc.make_track(xs[i], fmts[i], 44100)], [1.0])
        got = e.product(chain[i][0], "output")
        assert_bit_equal(got.numpy(), np.stack([rl, rr]), f"resampler node {i}")
        xi = orc.extract_interleaved(np.stack([rl, rr]), FMT_FLTP)
        y, _, _ = orc.soundtouch(xi, 48000, 1.0, pitch, 1152)
        got = e.product(chain[i][1], "output")
        assert got.frames == y.shape[0]
        if y.shape[0]:
            assert_bit_equal(got.numpy(), y, f"pitch node {i}")
        tracks.append(orc.make_track(orc.gain(y, FMT_FLT, 0.7), FMT_FLT, 48000) if y.shape[0] else None)
    if all(t is not None for t in tracks):
        rl, rr = orc.amix(tracks, [0.5, 0.5])
        assert_bit_equal(e.output().numpy(), np.stack([rl, rr]), "mix of the short streams")


@pytest.mark.parametrize("env", [{"NODEY_WAVE": "16"}, {"NODEY_WAVES": "16,8,8", "NODEY_COMPUTE_LANES": "3"},
                                 {"NODEY_WAVE": "16", "NODEY_COMPUTE_LANES": "1"}, {"NODEY_WAVES": "8,24"}],
                         ids=["two-waves-two-lanes", "three-waves-three-lanes", "two-waves-one-lane", "uneven-waves"])
def test_config5_graph_in_waves_on_several_lanes(eng_gpu, orc, env, monkeypatch):
    """The Runner's wave schedule (what a large render uses: blocks of source pins on rotating compute lanes, nodes that
    join several waves -- the level-1 mixes of a group cut by a wave boundary, the master mix -- waiting on the other
    lanes' events) must not change a single bit of the result.  Forced here at small size through the development
    overrides; host-bound sources so that the uploads run on the transfer lane as well."""
    from oracle import graph_oracle as G
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n = 44100 + 333
    tracks = [orc.synth_f32(n, 2, 44100, t) for t in range(32)]
    gains = [G.track_gain(t) for t in range(32)]
    p, ids = eng_gpu.config5_project(32, gains)
    e = eng_gpu.Engine(p.json())
    for t, x in enumerate(tracks):
        e.bind_source(t, x, FMT_FLT, 44100)          # numpy arrays: uploaded by audio_input
    ref_bus, ref_spec = G.render(tracks, threads=8)
    for _ in range(2):                                # twice: the second run reuses cached device blocks across lanes
        e.run()
        assert_bit_equal(e.output().numpy(), ref_bus, f"master bus with {env}")
    spec = e.product(ids["spectrum"], "output").numpy()
    peak = np.abs(ref_spec).max(axis=-1, keepdims=True)
    assert (np.abs(spec - ref_spec) <= 1e-5 * np.maximum(peak, 1e-30)).all()


@pytest.mark.parametrize("settings,waves,lanes", [
    (dict(wave_pins=16), 2, 2),
    (dict(wave_pattern=[16, 8, 8], compute_lanes=3), 3, 3),
    (dict(wave_pins=8, compute_lanes=1, stream_chunks=3, side_streams=0), 4, 1),
    (dict(wave_pattern="8,24", stream_priority=0, stream_chunks=64), 2, 2)],
    ids=["two-waves", "three-waves-three-lanes", "four-waves-one-lane-no-side-streams", "uneven-waves-64-chunks"])
def test_schedule_set_through_the_api(eng_gpu, orc, settings, waves, lanes, monkeypatch):
    """infra::Runner::Schedule through nodey_engine_set_schedule: the same knobs as the development environment
    variables, per engine; an explicit setting wins over the environment.  No schedule may change a bit of the result."""
    from oracle import graph_oracle as G
    monkeypatch.setenv("NODEY_WAVE", "1000000")       # would make it one wave: must lose against the explicit setting
    n = 44100 + 333
    tracks = [orc.synth_f32(n, 2, 44100, t) for t in range(32)]
    p, ids = eng_gpu.config5_project(32, [G.track_gain(t) for t in range(32)])
    e = eng_gpu.Engine(p.json())
    e.set_schedule(**settings)
    for t, x in enumerate(tracks):
        e.bind_source(t, x, FMT_FLT, 44100)
    ref_bus, _ = G.render(tracks, threads=8, spectrum=False)
    e.run()
    assert_bit_equal(e.output().numpy(), ref_bus, f"master bus with {settings}")
    steps = e.level_timings()
    assert len({s["wave"] for s in steps}) == waves
    assert len({s["lane"] for s in steps if s["level"] > 0}) == lanes
    with pytest.raises(eng_gpu.EngineError, match="unknown schedule key"):
        e.set_schedule(no_such_knob=1)
    with pytest.raises(eng_gpu.EngineError, match="compute_lanes"):
        e.set_schedule(compute_lanes=9)
    e.close()


def test_device_resident_render_of_128_tracks_takes_two_waves_and_matches(eng_gpu, orc, nd):
    """from 128 source pins on, a render whose sources are already in HBM runs as two half-size waves on two lanes (the
    default the bench uses at 256 tracks): same bits as the oracle graph and as the single-wave schedule"""
    import torch
    from oracle import graph_oracle as G
    n = 44100 + 17
    T = 128
    tracks = [orc.synth_f32(n, 2, 44100, t) for t in range(T)]
    gains = [G.track_gain(t) for t in range(T)]
    p, ids = eng_gpu.config5_project(T, gains)
    e = eng_gpu.Engine(p.json())
    dev = [torch.from_numpy(x).cuda() for x in tracks]
    for t in range(T):
        e.bind_source(t, dev[t], FMT_FLT, 44100)
    e.run()
    two = e.output().numpy().copy()
    ref_bus, _ = G.render(tracks, threads=8, spectrum=False)
    assert_bit_equal(two, ref_bus, "two-wave master bus vs oracle")
    import os
    os.environ["NODEY_WAVE"] = "1000000"
    try:
        e.run()
        assert_bit_equal(e.output().numpy(), two, "single-wave schedule")
    finally:
        del os.environ["NODEY_WAVE"]


# ---- the plugin-API nodes against the REAL libswresample (tests/golden/swr_real.npz) -------------------------
def _swr_gold():
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_swr_golden as G
    return G, np.load(os.path.join(here, "golden", "swr_real.npz"))


@pytest.mark.parametrize("tag", ["amix1_44100", "amix2_mixed", "amix3_rates"])
def test_amix_node_against_real_libswresample(eng_gpu, orc, tag):
    """project JSON -> Graph -> Runner -> Audio_amix on the GPU, compared with the reference's audio_amix loop run
    on real SwrContexts (fixture made by tests/golden/make_swr_golden.py): stream length and zero padding exact,
    samples within 1e-6 absolute (summation order of the FIR; the bar is 1e-5)"""
    G, gold = _swr_gold()
    spec, vols = G.AMIX_CASES[tag]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""] * len(spec)})
    mix = p.add("audio_amix", eng_gpu.amix_info(list(vols)))
    out = p.add("audio_output")
    for i in range(len(spec)):
        p.link(src, f"output_{i}", mix, f"input_{i + 1}")
    p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    for i, (r, f, c, n, fr) in enumerate(spec):
        e.bind_source(i, G.case_input(orc, r, f, c, n, 40 + i), f, r, frame_size=fr)
    e.run()
    got = e.output().numpy()
    gl, gr = gold[f"{tag}_l"], gold[f"{tag}_r"]
    assert got.shape == (2, len(gl)), "stream length (padding and flush iterations included)"
    assert np.abs(got[0] - gl).max() <= 1e-6 and np.abs(got[1] - gr).max() <= 1e-6
    assert np.array_equal(got[0] == 0, gl == 0)
    e.close()


@pytest.mark.parametrize("tag", ["bimix_mixed", "bimix_same_rate", "bimix_48k"])
def test_bimix_node_against_real_libswresample(eng_gpu, orc, tag):
    G, gold = _swr_gold()
    left, right, bias = G.BIMIX_CASES[tag]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    bm = p.add("audio_bimix", {"bias": bias})
    out = p.add("audio_output")
    p.link(src, "output_0", bm, "input_l"); p.link(src, "output_1", bm, "input_r"); p.link(bm, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.bind_source(0, G.case_input(orc, *left[:4], 50), left[1], left[0], frame_size=left[4])
    e.bind_source(1, G.case_input(orc, *right[:4], 51), right[1], right[0], frame_size=right[4])
    e.run()
    got = e.output().numpy()
    gl, gr = gold[f"{tag}_l"], gold[f"{tag}_r"]
    assert got.shape == (2, len(gl))
    tol = 0.0 if (left[0] == 48000 and right[0] == 48000) else 1e-6
    assert np.abs(got[0] - gl).max() <= tol and np.abs(got[1] - gr).max() <= tol
    e.close()


def test_amix_node_in_the_capped_surplus_regime_against_real_libswresample(eng_gpu, orc):
    """audio_amix with nb = 64 next to an 88.2 kHz input in 3561-sample frames (tests/golden/swr_real_amix_capped.npz):
    the real library hands the buffered surplus out over 160 later calls and returns 10213 samples, one MORE than a
    conversion with ample capacity -- the KERNEL has to produce that 10213th sample (audio-amix.cpp:263-290).  Stream
    length and zero positions exact, values within 1e-6; bit exact against the oracle node."""
    import os
    G, _ = _swr_gold()
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swr_real_amix_capped.npz"))
    tag = "amix4_capped_surplus"
    spec, vols = G.AMIX_CAPPED_CASES[tag]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""] * len(spec)})
    mix = p.add("audio_amix", eng_gpu.amix_info(list(vols)))
    out = p.add("audio_output")
    for i in range(len(spec)):
        p.link(src, f"output_{i}", mix, f"input_{i + 1}")
    p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    xs = [G.case_input(orc, r, f, c, n, 40 + i) for i, (r, f, c, n, fr) in enumerate(spec)]
    for i, (r, f, c, n, fr) in enumerate(spec):
        e.bind_source(i, xs[i], f, r, frame_size=fr)
    e.run()
    got = e.output().numpy()
    gl, gr = gold[f"{tag}_l"], gold[f"{tag}_r"]
    assert int(gold[f"{tag}_capped_total"]) == 10213
    assert got.shape == (2, len(gl)) and len(gl) == int(gold[f"{tag}_nb"].sum()), "stream length"
    assert np.array_equal(got[0] == 0, gl == 0) and np.array_equal(got[1] == 0, gr == 0), "zero positions (the 10213th sample)"
    assert np.abs(got[0] - gl).max() <= 1e-6 and np.abs(got[1] - gr).max() <= 1e-6
    ol, orr = orc.amix([orc.make_track(x, f, r, frame_size=fr) for x, (r, f, c, n, fr) in zip(xs, spec)], list(vols))
    assert_bit_equal(got, np.stack([ol, orr]), "capped amix vs oracle node")
    e.close()


def _bimix2_cases():
    G, _ = _swr_gold()
    return sorted(G.BIMIX2_CASES)


@pytest.mark.parametrize("tag", _bimix2_cases())
def test_bimix_v2_node_against_real_libswresample(eng_gpu, orc, tag):
    """Audio_bimix_v2 on the GPU against the reference's loop (audio-bimix.cpp:536-875) run on real SwrContexts
    (tests/golden/swr_real_bimix2.npz): unflushed per-frame conversion with capacity 2 * nb, mono fold, END-time stamps,
    the pts aligner.  Stream length, first pts and every silent position exact; values 1e-6 (bit exact at 48 kHz)."""
    import os
    G, _ = _swr_gold()
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swr_real_bimix2.npz"))
    left, right, pl, pr = G.BIMIX2_CASES[tag]
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    bm = p.add("audio_bimix_v2")
    out = p.add("audio_output")
    p.link(src, "output_0", bm, "input_l"); p.link(src, "output_1", bm, "input_r"); p.link(bm, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.bind_source(0, G.case_input(orc, *left[:4], 70), left[1], left[0], frame_size=left[4], pts=pl)
    e.bind_source(1, G.case_input(orc, *right[:4], 71), right[1], right[0], frame_size=right[4], pts=pr)
    e.run()
    got = e.output()
    ref = gold[f"{tag}_out"]
    g = got.numpy()
    assert g.shape == ref.shape == (int(gold[f"{tag}_sizes"].sum()), 2), "stream length"
    assert got.pts == float(gold[f"{tag}_pts"]), "pts of the first frame"
    assert np.array_equal(g == 0, ref == 0), "silent regions (alignment)"
    exact = left[0] == 48000 and right[0] == 48000
    assert np.abs(g - ref).max() <= (0.0 if exact else 1e-6)
    e.close()


def test_diagnostics_report_every_step_of_the_run(eng_gpu, orc):
    """the data behind the editor's overlay -- node states and the device time of every
    (wave, level) step of the Runner (the reference shows channel fill per link, app.cpp:1556-1592)."""
    n = 44100
    project, ids = eng_gpu.config5_project(16, [1.0] * 16)
    e = eng_gpu.Engine(project.json())
    with pytest.raises(eng_gpu.EngineError):
        e.diagnostics()                                   # nothing has run yet
    for t in range(16):
        e.bind_source(t, orc.synth_f32(n, 2, 44100, t), 3, 44100)
    e.run()
    steps = e.level_timings()
    levels = {lvl for _, _, lvl in e.nodes()}
    assert {s["level"] for s in steps} == levels          # every level of the graph ran as at least one step
    assert sum(s["nodes"] for s in steps) == len(e.nodes())
    assert steps[0]["level"] == 0 and steps[0]["lane"] == 0 and steps[0]["start_ms"] == 0.0      # sources on the transfer lane
    assert all(s["lane"] >= 1 for s in steps[1:])
    assert all(s["device_ms"] >= 0.0 and s["enqueue_ms"] > 0.0 and s["start_ms"] >= 0.0 for s in steps)
    heavy = max(steps, key=lambda s: s["device_ms"])
    assert heavy["device_ms"] > 0.05                      # the SoundTouch steps take measurable device time
    text = e.diagnostics().splitlines()
    assert text[0] == f"0 Running | {len(e.nodes())} Finished | 0 Errors"
    assert len(text) == 1 + len(steps) and text[1].startswith("W0 L0 audio_input x1:")
    assert any("pitch_modifier x16" in line for line in text)
    e.close()


@pytest.mark.parametrize("fmt", [FMT_FLT, FMT_FLTP])
def test_lazy_gain_products_fold_into_the_mixer_and_materialise_for_everyone_else(eng_gpu, orc, fmt):
    """audio_volume_adjust on a float stream publishes "source x gain" (Lazy_gain): the amix behind it multiplies while it
    reads ((x * g) rounded, then * volume -- the bits audio-vol.cpp:75-100 + audio-amix.cpp:296-304 give), every other
    consumer (here: the sink of a second gain node, a spectrum node, the host reading the product back) gets the
    ordinary buffer.  All bit exact against the oracle."""
    n = 48000 * 2 + 77
    xa = make_input(orc, fmt, n, 2, rate=48000, track=1); xb = make_input(orc, fmt, n - 5000, 2, rate=48000, track=2)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": ["", ""]})
    ga = p.add("audio_volume_adjust", {"volume": 0.8125}); gb = p.add("audio_volume_adjust", {"volume": 1.7})
    gc = p.add("audio_volume_adjust", {"volume": 0.3})              # gain behind a gain: the inner one materialises
    mix = p.add("audio_amix", eng_gpu.amix_info([0.5, 0.25, 1.0]))
    spec = p.add("audio_spectrum", {"fft_size": 4096, "hop": 1024, "window": "hann"})
    out = p.add("audio_output")
    p.link(src, "output_0", ga, "input"); p.link(src, "output_1", gb, "input"); p.link(ga, "output", gc, "input")
    p.link(ga, "output", mix, "input_1"); p.link(gb, "output", mix, "input_2"); p.link(gc, "output", mix, "input_3")
    p.link(gb, "output", spec, "input")
    p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.bind_source(0, xa, fmt, 48000); e.bind_source(1, xb, fmt, 48000)
    e.run()
    ya = orc.gain(xa, fmt, 0.8125); yb = orc.gain(xb, fmt, 1.7); yc = orc.gain(ya, fmt, 0.3)
    rl, rr = orc.amix([orc.make_track(ya, fmt, 48000), orc.make_track(yb, fmt, 48000), orc.make_track(yc, fmt, 48000)], [0.5, 0.25, 1.0])
    assert_bit_equal(e.output().numpy(), np.stack([rl, rr]), "amix over lazy gain products")
    assert_bit_equal(e.product(ga, "output").numpy(), ya, "gain product read back by the host")
    assert_bit_equal(e.product(gc, "output").numpy(), yc, "gain behind a gain")
    sp = e.product(spec, "output").numpy()
    yb_planar = yb if fmt == FMT_FLTP else np.ascontiguousarray(yb.T)
    ref = orc.stft(yb_planar[0].copy())
    assert np.abs(sp[0] - ref).max() <= 1e-5 * np.abs(ref).max()
    e.close()


def test_release_mode_renders_the_same_bus_with_a_smaller_footprint(nd, eng_gpu, orc):
    """Runner::release_products: links drop their products once consumed.  Same master bus bit for bit, intermediates
    gone after the run, lower high-water mark of device memory."""
    import torch
    n = 44100 * 4
    project, ids = eng_gpu.config5_project(32, [1.0 + 0.01 * t for t in range(32)])
    xs = [nd.synth(n, 2, 44100, track=t) for t in range(32)]

    def render(release):
        eng_gpu.set_release_products(release)
        try:
            e = eng_gpu.Engine(project.json())
            for t in range(32):
                e.bind_source(t, xs[t], nd.FMT_FLT, 44100)
            torch.cuda.synchronize()
            nd.check(nd.lib().nodey_trim_memory())             # start from an empty cache: the footprint of THIS policy
            nd.memory_stats(reset_peak=True)
            e.run(); e.run()
            torch.cuda.synchronize()
            _, peak = nd.memory_reserved()
            return e, e.output().numpy(), peak
        finally:
            eng_gpu.set_release_products(False)

    e0, bus0, peak0 = render(False)
    e0.close()
    e1, bus1, peak1 = render(True)
    assert_bit_equal(bus1, bus0, "bus in release mode")
    with pytest.raises(eng_gpu.EngineError):
        e1.product(ids["groups"][0], "output")
    assert e1.product(ids["spectrum"], "output").frames > 0
    assert peak1 < peak0, (peak0, peak1)
    e1.close()
