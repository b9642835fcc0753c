"""GPU parity: pitch_modifier / velocity_modifier (SoundTouch model) against the oracle.  The WSOLA
offset trace must be identical and, because every kernel reproduces the oracle's rounding sequence,
the samples are compared bit for bit (the 1e-5 float bar of BASELINE.json is then trivially met)."""
import numpy as np
import pytest

from helpers import assert_bit_equal, to_dev

pytestmark = pytest.mark.gpu

CASES = [
    # (sample_rate, ch, rate_arg, pitch_arg-or-None(semitones), seconds)
    ("pitch+3", 48000, 2, 1.0, ("st", 3.0), 4.0),
    ("pitch-3", 48000, 2, 1.0, ("st", -3.0), 4.0),
    ("tempo1.25keep", 48000, 2, 1.25, ("keep", 1.25), 4.0),
    ("tempo0.8keep", 48000, 2, 0.8, ("keep", 0.8), 3.0),
    ("velocity1.25", 48000, 2, 1.25, ("none", 0), 3.0),
    ("velocity0.7", 44100, 2, 0.7, ("none", 0), 3.0),
    ("pitch+7 44k", 44100, 2, 1.0, ("st", 7.0), 3.0),
    ("mono pitch+3", 48000, 1, 1.0, ("st", 3.0), 3.0),
    ("mono tempo 22k", 22050, 1, 1.5, ("keep", 1.5), 3.0),
    ("11k stereo", 11025, 2, 1.0, ("st", -5.0), 3.0),
    ("identity", 48000, 2, 1.0, ("none", 0), 2.0),
]


def _pitch(orc, spec, rate):
    kind, v = spec
    if kind == "st":
        return orc.pitch_node_factor(v)
    if kind == "keep":
        return orc.velocity_node_pitch(v, True)
    return 1.0


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_soundtouch_matches_oracle(nd, orc, case):
    _, sr, ch, rate, spec, secs = case
    n = int(sr * secs) + 17
    x = orc.synth_f32(n, ch, sr, 5)
    pitch = _pitch(orc, spec, rate)
    ref, ref_offs, info = orc.soundtouch(x, sr, rate, pitch, 1152)
    st = nd.SoundTouch(sr, ch, rate, pitch)
    gi = st.info()
    for k in ("overlap", "seek_window", "seek_length", "sample_req", "tdstretch_first"):
        assert gi[k] == getattr(info, k), k
    assert gi["rate"] == info.rate and gi["tempo"] == info.tempo and gi["nominal_skip"] == info.nominal_skip
    m, nseq = st.out_frames(n, 1152)
    assert m == ref.shape[0]
    assert nseq == info.n_sequences
    got, offs = st.run(to_dev(x), 1152, want_offsets=True)
    assert np.array_equal(offs.cpu().numpy(), ref_offs), "WSOLA offset trace differs"
    assert_bit_equal(got.cpu().numpy(), ref, "soundtouch samples")


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("cfg", [(48000, 2, 1.0, 3.0), (44100, 1, 1.0, -4.0), (11025, 2, 1.0, 2.0)])
def test_soundtouch_cluster_sizes(nd, orc, cluster, cfg):
    """the WSOLA search split over a thread-block cluster must give the same trace as one CTA"""
    sr, ch, rate, st_ = cfg
    n = sr * 3
    xs = np.stack([orc.synth_f32(n, ch, sr, t) for t in range(3)])
    pitch = orc.pitch_node_factor(st_)
    st = nd.SoundTouch(sr, ch, rate, pitch)
    st.set_cluster(cluster)
    got, offs = st.run(to_dev(xs), 1152, want_offsets=True)
    for t in range(3):
        ref, ro, _ = orc.soundtouch(xs[t], sr, rate, pitch, 1152)
        assert np.array_equal(offs[t].cpu().numpy(), ro), f"offset trace, track {t}"
        assert_bit_equal(got[t].cpu().numpy(), ref, f"track {t}")


@pytest.mark.parametrize("kt", [2, 4, 8, 11, 12, 13, 14, 15, 16])
def test_soundtouch_candidates_per_thread(nd, orc, kt):
    """every (candidates per thread, cluster size) variant of the WSOLA search gives the oracle's trace: the pitch node's
    912 candidates and the tempo node's 864 at 48 kHz, 44.1 kHz (810 / 793) and the run-time stride"""
    import os
    cases = [(48000, 1.0, orc.pitch_node_factor(3.0)), (48000, 1.25, orc.velocity_node_pitch(1.25, True)),
             (44100, 1.0, orc.pitch_node_factor(-5.0))]
    for sr, rate, pitch in cases:
        n = sr * 2 + 77
        xs = np.stack([orc.synth_f32(n, 2, sr, 20 + t) for t in range(2)])
        refs = [orc.soundtouch(xs[t], sr, rate, pitch, 1152) for t in range(2)]
        st = nd.SoundTouch(sr, 2, rate, pitch)
        st.set_candidates_per_thread(kt)
        for cluster in (1, 2, 4, 8):
            st.set_cluster(cluster)
            for runtime_sk in (False, True):
                if runtime_sk:
                    os.environ["NODEY_TDS_RUNTIME_SK"] = "1"
                try:
                    got, offs = st.run(to_dev(xs), 1152, want_offsets=True)
                finally:
                    os.environ.pop("NODEY_TDS_RUNTIME_SK", None)
                for t in range(2):
                    assert np.array_equal(offs[t].cpu().numpy(), refs[t][1]), f"offset trace: kt {kt} cluster {cluster} sr {sr} rate {rate} track {t}"
                    assert_bit_equal(got[t].cpu().numpy(), refs[t][0], f"kt {kt} cluster {cluster} track {t}")


@pytest.mark.parametrize("cfg", [(48000, 1.0, 3.0, None), (48000, 1.25, None, True), (44100, 1.7, None, False), (8000, 1.0, 11.0, None)])
def test_soundtouch_fused_and_unfused_tails_agree(nd, orc, cfg):
    """stereo, TDStretch-first: the fused cross-fade + FIR + cubic kernel and the three separate kernels"""
    sr, rate, st_, keep = cfg
    pitch = orc.pitch_node_factor(st_) if st_ is not None else orc.velocity_node_pitch(rate, keep)
    n = sr * 3 + 5
    x = orc.synth_f32(n, 2, sr, 8)
    ref, _, info = orc.soundtouch(x, sr, rate, pitch, 1152)
    assert info.tdstretch_first == 1
    st = nd.SoundTouch(sr, 2, rate, pitch)
    fused = st.run(to_dev(x)).cpu().numpy()
    st.set_unfused(True)
    unfused = st.run(to_dev(x)).cpu().numpy()
    assert_bit_equal(fused, ref, "fused tail")
    assert_bit_equal(unfused, ref, "unfused tail")


def test_soundtouch_batch_and_chunking(nd, orc):
    sr, n = 48000, 48000 * 2
    xs = np.stack([orc.synth_f32(n, 2, sr, t) for t in range(5)])
    pitch = orc.pitch_node_factor(3.0)
    st = nd.SoundTouch(sr, 2, 1.0, pitch)
    got, offs = st.run(to_dev(xs), 1152, want_offsets=True)
    for t in range(5):
        ref, ro, _ = orc.soundtouch(xs[t], sr, 1.0, pitch, 1152)
        assert np.array_equal(offs[t].cpu().numpy(), ro)
        assert_bit_equal(got[t].cpu().numpy(), ref, f"track {t}")
    # putSamples chunking must not change the result (FIFO-driven pipeline)
    ref4096, _, _ = orc.soundtouch(xs[0], sr, 1.0, pitch, 4096)
    got4096 = st.run(to_dev(xs[0]), 4096)
    assert_bit_equal(got4096.cpu().numpy(), ref4096, "frame_size 4096")


def test_soundtouch_short_and_errors(nd, orc):
    st = nd.SoundTouch(48000, 2, 1.0, orc.pitch_node_factor(3.0))
    for n in (1, 100, 3000, 4704, 6000):
        x = orc.synth_f32(n, 2, 48000, 1)
        ref, ro, _ = orc.soundtouch(x, 48000, 1.0, orc.pitch_node_factor(3.0), 1152)
        m, _ = st.out_frames(n, 1152)
        assert m == ref.shape[0], n
        if m:
            got = st.run(to_dev(x), 1152)
            assert_bit_equal(got.cpu().numpy(), ref, f"short {n}")
    with pytest.raises(nd.NodeyError) as e:
        nd.SoundTouch(96000, 2, 1.0, 1.0)       # audio-velocity.cpp:371 accepts 8 k .. 48 k only
    assert e.value.code == -5
    with pytest.raises(nd.NodeyError):
        nd.SoundTouch(48000, 3, 1.0, 1.0)


IN_PLACE = [
    # (sample_rate, ch, rate, semitones-or-None, planar)
    (48000, 2, 1.0, 3.0, True),        # TDStretch first: offsets + fused tail read the planes
    (48000, 2, 1.0, 3.0, False),
    (48000, 2, 0.8, None, True),       # rate <= 1: the cubic transposer reads the planes first
    (44100, 2, 1.25, None, True),      # odd length: plane pointers not 8-byte aligned relative to each other
    (48000, 1, 1.0, -2.0, False),      # mono
]


@pytest.mark.parametrize("unfused", [0, 1])
@pytest.mark.parametrize("cfg", IN_PLACE, ids=[f"{c[0]}-{c[1]}ch-r{c[2]}-st{c[3]}-{'planar' if c[4] else 'packed'}" for c in IN_PLACE])
def test_soundtouch_reads_tracks_in_place(nd, orc, cfg, unfused):
    """nodey_soundtouch_run_tracks (per-track pointers, FLT or FLTP read in place) == the staged batch == the oracle"""
    import torch
    sr, ch, rate, st_, planar = cfg
    pitch = orc.pitch_node_factor(st_) if st_ is not None else 1.0
    n = int(sr * 2.5) + 7
    xs = [orc.synth_f32(n, ch, sr, 20 + t) for t in range(3)]
    st = nd.SoundTouch(sr, ch, rate, pitch)
    st.set_unfused(unfused)
    if planar:
        devs = [(to_dev(np.ascontiguousarray(x[:, 0])), to_dev(np.ascontiguousarray(x[:, 1]))) for x in xs]
    else:
        devs = [to_dev(x) for x in xs]
    got, offs = st.run_tracks(devs, 1152, want_offsets=True)
    torch.cuda.synchronize()
    for t, x in enumerate(xs):
        ref, ref_offs, _ = orc.soundtouch(x, sr, rate, pitch, 1152)
        assert np.array_equal(offs[t].cpu().numpy()[:len(ref_offs)], ref_offs), f"track {t}: WSOLA offset trace differs"
        assert_bit_equal(got[t].cpu().numpy(), ref, f"track {t} samples")


@pytest.mark.parametrize("n", [0, 1, 100, 1152, 4000, 5615, 5616, 5617, 9000])
@pytest.mark.parametrize("cfg", [(48000, 2, 1.0, 3.0), (48000, 2, 1.25, None), (44100, 1, 0.8, None)])
def test_soundtouch_short_inputs(nd, orc, n, cfg):
    """inputs shorter than one WSOLA sequence, around the first sequence boundary and just above: flush() pads
    with silence until the expected amount exists (or nothing at all comes out) -- lengths and samples as the oracle"""
    sr, ch, rate, st_ = cfg
    pitch = orc.pitch_node_factor(st_) if st_ is not None else orc.velocity_node_pitch(rate, True)
    x = orc.synth_f32(max(n, 1), ch, sr, 31)[:n]
    ref, ref_offs, info = orc.soundtouch(x, sr, rate, pitch, 1152)
    st = nd.SoundTouch(sr, ch, rate, pitch)
    m, nseq = st.out_frames(n, 1152)
    assert m == ref.shape[0] and nseq == info.n_sequences
    if n == 0:
        return
    got, offs = st.run(to_dev(x), 1152, want_offsets=True)
    assert got.shape[0] == m
    if m:
        assert np.array_equal(offs.cpu().numpy()[:len(ref_offs)], ref_offs)
        assert_bit_equal(got.cpu().numpy(), ref, f"short input n={n}")


@pytest.mark.parametrize("nchunks", [2, 3, 8, 64])
@pytest.mark.parametrize("cfg", [(48000, 1.0, 3.0, None, 12.0), (48000, 1.25, None, True, 9.0), (44100, 1.7, None, False, 7.0)])
def test_soundtouch_chunked_render_is_bit_identical(nd, orc, nchunks, cfg):
    """A track's WSOLA chain cut into launches (nodey_soundtouch_run_chunk): same offset trace, same samples as the
    one-launch render and as the oracle; and chunk c really needs no more than in_need[c] input frames -- everything
    beyond is overwritten with NaN before the chunk runs and restored afterwards."""
    import torch
    sr, rate, st_, keep, secs = cfg
    pitch = orc.pitch_node_factor(st_) if st_ is not None else orc.velocity_node_pitch(rate, keep)
    n = int(sr * secs) + 31
    xs = np.stack([orc.synth_f32(n, 2, sr, t) for t in range(3)])
    st = nd.SoundTouch(sr, 2, rate, pitch)
    x = to_dev(xs)
    whole, whole_offs = st.run(x, 1152, want_offsets=True)
    keep_x = x.clone()

    def poison(c, in_need):
        x.copy_(keep_x)
        x[:, in_need:, :] = float("nan")

    got, offs, plan = st.run_chunked(x, nchunks, 1152, poison=poison)
    torch.cuda.synchronize()
    assert 1 <= len(plan) <= nchunks
    assert plan[-1][0] == n and plan[-1][1] == whole.shape[1]
    assert all(plan[k][0] <= plan[k + 1][0] and plan[k][1] <= plan[k + 1][1] for k in range(len(plan) - 1))
    assert len(plan) == min(nchunks, max(1, (offs.shape[1]) // 32)), "at least 32 sequences per launch"
    assert torch.equal(offs, whole_offs), "offset trace of the chunked render"
    assert_bit_equal(got.cpu().numpy(), whole.cpu().numpy(), "chunked vs one launch")
    ref, ro, _ = orc.soundtouch(xs[1], sr, rate, pitch, 1152)
    assert np.array_equal(offs[1].cpu().numpy(), ro)
    assert_bit_equal(got[1].cpu().numpy(), ref, "chunked vs oracle")


def test_soundtouch_chunk_progress_is_final(nd, orc):
    """after chunk c, output frames [0, out_ready[c]) already hold their final values"""
    import torch
    sr, n = 48000, 48000 * 10
    x = to_dev(np.stack([orc.synth_f32(n, 2, sr, t) for t in range(2)]))
    st = nd.SoundTouch.pitch_node(sr, 2, 3.0)
    whole = st.run(x, 1152)
    m, nseq = st.out_frames(n, 1152)
    plan = st.chunks(n, 1152, 6)
    out = torch.full((2, m, 2), float("nan"), device="cuda")
    offs = torch.zeros((2, nseq - 1), dtype=torch.int32, device="cuda")
    import ctypes as C
    for c in range(len(plan)):
        nd.check(nd.lib().nodey_soundtouch_run_chunk(st.h, nd._dp(out), out.stride(0), nd._dp(x), x.stride(0), 2, n, 1152, m,
                                                     nd._dp(offs), offs.stride(0), c, len(plan), 0, nd._stream()))
        torch.cuda.synchronize()
        r = plan[c][1]
        assert torch.equal(out[:, :r], whole[:, :r]), f"chunk {c}: frames below out_ready are not final"
    assert plan[0][1] > 0 and plan[0][0] < n


def test_soundtouch_chunks_of_uncuttable_paths(nd, orc):
    """mono and the rate <= 1 order run as one chunk; a chunk count the plan did not return is rejected"""
    st = nd.SoundTouch(48000, 1, 1.0, orc.pitch_node_factor(3.0))
    assert len(st.chunks(48000 * 5, 1152, 8)) == 1
    st2 = nd.SoundTouch(48000, 2, 0.7, 1.0)            # rate 0.7 <= 1: transposer first, TDStretch last
    assert st2.info()["tdstretch_first"] == 0 and len(st2.chunks(48000 * 5, 1152, 8)) == 1
    st3 = nd.SoundTouch.pitch_node(48000, 2, 3.0)
    import torch
    x = torch.zeros((1, 48000, 2), device="cuda")          # 1 s: 14 sequences -> one chunk
    assert len(st3.chunks(48000, 1152, 8)) == 1
    m, nseq = st3.out_frames(48000, 1152)
    out = torch.empty((1, m, 2), device="cuda")
    offs = torch.zeros((1, max(nseq - 1, 1)), dtype=torch.int32, device="cuda")
    rc = nd.lib().nodey_soundtouch_run_chunk(st3.h, nd._dp(out), out.stride(0), nd._dp(x), x.stride(0), 1, 48000, 1152, m,
                                             nd._dp(offs), offs.stride(0), 0, 4, 0, nd._stream())
    assert rc == -5      # NODEY_E_RANGE


@pytest.mark.parametrize("cfg", [(48000, 2, 1.0, ("st", 3.0), 4.0), (48000, 2, 1.25, ("keep", 1.25), 4.0), (44100, 2, 0.7, ("none", 0), 3.0),
                                 (48000, 1, 1.0, ("st", 3.0), 3.0), (22050, 1, 1.5, ("keep", 1.5), 3.0), (48000, 2, 2.0, ("keep", 2.0), 2.5),
                                 (48000, 2, 1.0, ("st", 3.0), 0.05)])
def test_reference_schedule_matches_the_restated_node_loop(nd, orc, cfg):
    """SURVEY.md App. C7 switch: nodey_soundtouch_reference_schedule (host arithmetic on lengths) against the literal
    restatement of soundtouch_process_payload on the oracle's streaming model (orc_soundtouch_reference_loop): total
    frames, every receive size, whether flush() was reached; and the kernels' render cut to that total is the loop's output."""
    sr, ch, velocity, spec, secs = cfg
    n = int(sr * secs) + 17
    x = orc.synth_f32(n, ch, sr, 9)
    pitch = _pitch(orc, spec, velocity)
    ref, sizes, flushed = orc.soundtouch_reference_loop(x, sr, velocity, pitch, 1152)
    st = nd.SoundTouch(sr, ch, velocity, pitch)
    total, runs, fl = st.reference_schedule(n, velocity)
    assert total == ref.shape[0]
    assert [s for s, c in runs for _ in range(c)] == list(sizes)
    assert fl == flushed
    canonical = st.run(to_dev(x), 1152).cpu().numpy()
    assert total <= canonical.shape[0]
    assert_bit_equal(canonical[:total], ref, "reference-loop output = prefix of the canonical render")


def test_reference_schedule_switch_on_the_nodes(eng_gpu, orc):  # noqa: C901
    """project JSON key "reference_schedule": the pitch and the tempo node emit the reference loop's frames (sizes and
    count) instead of the canonical ones; off by default and absent from serialised projects unless set"""
    import json
    sr, n = 48000, 48000 * 3 + 5
    x = orc.synth_f32(n, 2, sr, 4)

    def render(flag):
        p = eng_gpu.Project()
        src = p.add("audio_input", {"file_path": [""]})
        pm = p.add("pitch_modifier", dict({"pitch": 3.0}, **({"reference_schedule": True} if flag else {})))
        vm = p.add("velocity_modifier", dict({"velocity": 1.25, "keep_pitch": True}, **({"reference_schedule": True} if flag else {})))
        out = p.add("audio_output")
        p.link(src, "output_0", pm, "input"); p.link(pm, "output", vm, "input"); p.link(vm, "output", out, "input")
        e = eng_gpu.Engine(p.json())
        e.bind_source(0, x, FMT_FLT_, sr, pts=0.37)
        e.run()
        text = json.loads(e.serialize())
        return e, pm, vm, text

    e, pm, vm, text = render(True)
    y1, s1, _ = orc.soundtouch_reference_loop(x, sr, 1.0, orc.pitch_node_factor(3.0), 1152)
    y2, s2, _ = orc.soundtouch_reference_loop(y1, sr, 1.25, orc.velocity_node_pitch(1.25, True), 1152)
    assert_bit_equal(e.product(pm, "output").numpy(), y1, "pitch node, reference schedule")
    assert [s for s, c in e.product_runs(pm, "output") for _ in range(c)] == list(s1)
    assert_bit_equal(e.output().numpy(), y2, "tempo node, reference schedule")
    assert [s for s, c in e.product_runs(vm, "output") for _ in range(c)] == list(s2)
    assert text["nodes"][str(pm)]["info"]["reference_schedule"] is True
    # App. C8 (same switch): frames stamped through a float of microseconds, from a running time that starts at the first
    # input frame's stamp (audio-velocity.cpp:238-249, 313-318, 388); the tempo node starts from the pitch node's first stamp
    first = int(np.float32(0.37 * 1000000)) * (1 / 1000000.0)
    assert e.product_stamp(pm, "output") == (eng_gpu.STAMP_START_FLOAT_US, 0.37) and e.product(pm, "output").pts == first
    second = int(np.float32(first * 1000000)) * (1 / 1000000.0)
    assert e.product_stamp(vm, "output") == (eng_gpu.STAMP_START_FLOAT_US, first) and e.output().pts == second
    e.close()
    e, pm, vm, text = render(False)
    assert e.product_stamp(vm, "output") == (eng_gpu.STAMP_START, 0.37) and e.output().pts == 0.37
    c1, _, _ = orc.soundtouch(x, sr, 1.0, orc.pitch_node_factor(3.0), 1152)
    c2, _, _ = orc.soundtouch(c1, sr, 1.25, orc.velocity_node_pitch(1.25, True), 1152)
    assert_bit_equal(e.output().numpy(), c2, "canonical render")
    assert len(c2) > len(y2) and "reference_schedule" not in text["nodes"][str(pm)]["info"]
    e.close()


FMT_FLT_ = 3
