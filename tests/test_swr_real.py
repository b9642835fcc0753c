"""The oracle's libswresample model against a REAL libswresample.

`tests/golden/swr_real.npz` holds outputs of the stock libswresample 6.1.100 (FFmpeg 8.0.1) found in this
image (made by tests/golden/make_swr_golden.py through oracle/real_swr.py, configured like the reference's call
sites audio-amix.cpp:212-240).  Pinned here:

  * the filter bank, the phase positions and the edge handling, BIT EXACT: the response to a unit impulse is a
    sequence of single filter coefficients, whatever order the library sums in;
  * how many samples every swr_convert call returns (frame by frame, flush calls included), exact;
  * the format conversion / rematrix front end at equal rates, bit exact;
  * resampled values within 1e-6 absolute (signals of 0.4-0.57 peak; the bar of the task is 1e-5): the library
    itself has several summation orders (C template with two partial sums, SSE / AVX / FMA3 assembly with 4 / 8
    lanes), the oracle and the CUDA kernel fix one (single accumulator, ascending taps, fused multiply-add);
  * the whole audio_amix, audio_bimix and audio_bimix_v2 loops (nb = min frame size, zero padding, flush iterations,
    sequential mix; bimix_v2's unflushed per-frame conversion, END-time stamps and pts aligner) and the preview path's
    per-frame conversion to packed float.

The live tests repeat the comparison against the library itself (C template and the assembly this CPU selects)
where the image has it, and skip elsewhere."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_swr_golden as G  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "swr_real.npz"))
VALUE_TOL = 1e-6


def oracle_frames(orc, x, fmt, rate, ch, frame):
    """feed the oracle's streaming swr like real_swr.whole: per-frame converts, then flush until 0"""
    s = orc.Swr(rate, 48000, fmt, ch)
    L, R, counts = [], [], []
    for fr in G.frames_of(x, fmt, frame):
        m = fr.shape[1] if fmt >= 5 else fr.shape[0]
        l, r = s.convert(fr, int(m * 48000 / rate) + 4096)
        L.append(l.copy()); R.append(r.copy()); counts.append(len(l))
    while True:
        l, r = s.convert(None, 1 << 16)
        counts.append(len(l))
        if len(l) == 0:
            break
        L.append(l.copy()); R.append(r.copy())
    cat = lambda a: np.concatenate(a) if a else np.zeros(0, np.float32)
    return cat(L), cat(R), counts


@pytest.mark.parametrize("rate", G.IMPULSE_RATES)
def test_impulse_response_is_the_real_filter_bank(orc, rate):
    x = np.zeros((4000, 2), np.float32)
    x[2000, 0] = 1.0
    l, r = orc.swr_whole(x, orc.FMT_FLT, rate, 48000)
    assert len(l) == int(GOLD[f"impulse_{rate}_len"])
    nz = np.flatnonzero(l)
    first = int(GOLD[f"impulse_{rate}_first"])
    taps = GOLD[f"impulse_{rate}_taps"]
    assert nz[0] == first and nz[-1] == first + len(taps) - 1
    assert np.array_equal(l[first:first + len(taps)].view(np.uint32), taps.view(np.uint32)), \
        "filter coefficients differ from libswresample's"
    assert not r.any()


@pytest.mark.parametrize("case", G.VALUE_CASES, ids=[c[0] for c in G.VALUE_CASES])
def test_values_and_counts_match_real_library(orc, case):
    tag, rate, fmt, ch, n, frame = case
    x = G.case_input(orc, rate, fmt, ch, n, 20 + G.VALUE_CASES.index(case))
    l, r, counts = oracle_frames(orc, x, fmt, rate, ch, frame)
    assert counts == GOLD[f"{tag}_counts"].tolist(), "samples returned per swr_convert call"
    gl, gr = GOLD[f"{tag}_l"], GOLD[f"{tag}_r"]
    assert l.shape == gl.shape and r.shape == gr.shape
    if rate == 48000:
        assert np.array_equal(l.view(np.uint32), gl.view(np.uint32)) and np.array_equal(r.view(np.uint32), gr.view(np.uint32))
    elif len(gl):
        assert np.abs(l - gl).max() <= VALUE_TOL and np.abs(r - gr).max() <= VALUE_TOL
        # and the whole-buffer entry point the CUDA path is compared with gives the same samples
        wl, wr = orc.swr_whole(x, fmt, rate, 48000)
        assert np.array_equal(wl, l) and np.array_equal(wr, r)


@pytest.mark.parametrize("tag", sorted(G.AMIX_CASES))
def test_amix_loop_matches_real_library(orc, tag):
    spec, vols = G.AMIX_CASES[tag]
    tracks = [orc.make_track(G.case_input(orc, r, f, c, n, 40 + i), f, r, frame_size=fr)
              for i, (r, f, c, n, fr) in enumerate(spec)]
    l, r = orc.amix(tracks, vols)
    gl, gr = GOLD[f"{tag}_l"], GOLD[f"{tag}_r"]
    assert len(l) == len(gl) == int(GOLD[f"{tag}_nb"].sum()), "stream length (zero padding and flush iterations included)"
    assert np.abs(l - gl).max() <= VALUE_TOL and np.abs(r - gr).max() <= VALUE_TOL
    assert np.array_equal(l == 0, gl == 0), "zero padding in the same places"


@pytest.mark.parametrize("case", G.CAP_CASES, ids=[c[0] for c in G.CAP_CASES])
def test_limited_output_capacity_matches_real_library(orc, case):
    """swr_convert with less room than the input yields: what every call returns, how long the flush calls keep
    returning samples, and the samples themselves"""
    tag, rate, fmt, ch, n, frame, cap = case
    x = G.case_input(orc, rate, fmt, ch, n, 80 + G.CAP_CASES.index(case))
    l, r, counts = G.capped(lambda: orc.Swr(rate, 48000, fmt, ch), x, fmt, frame, cap)
    assert counts.tolist() == GOLD[f"{tag}_counts"].tolist()
    assert np.abs(l - GOLD[f"{tag}_l"]).max() <= VALUE_TOL and np.abs(r - GOLD[f"{tag}_r"]).max() <= VALUE_TOL


@pytest.mark.parametrize("tag", sorted(G.BIMIX_CASES))
def test_bimix_loop_matches_real_library(orc, tag):
    """orc_bimix against an independent restatement of audio-bimix.cpp:136-320 run on real SwrContexts"""
    left, right, bias = G.BIMIX_CASES[tag]
    tl = orc.make_track(G.case_input(orc, *left[:4], 50), left[1], left[0], frame_size=left[4])
    tr = orc.make_track(G.case_input(orc, *right[:4], 51), right[1], right[0], frame_size=right[4])
    l, r = orc.bimix(tl, tr, bias)
    gl, gr = GOLD[f"{tag}_l"], GOLD[f"{tag}_r"]
    assert len(l) == len(gl) == int(GOLD[f"{tag}_nb"].sum()), "stream length"
    exact = left[0] == 48000 and right[0] == 48000
    tol = 0.0 if exact else VALUE_TOL
    assert np.abs(l - gl).max() <= tol and np.abs(r - gr).max() <= tol


GOLD4 = np.load(os.path.join(HERE, "golden", "swr_real_short.npz"))


@pytest.mark.parametrize("rate", G.SHORT_RATES)
def test_streams_shorter_than_the_filter(orc, rate):
    """How many samples a whole stream of 1..80 frames yields (flush included), against the real library: nothing while
    the stream plus its flush reflection stays below filter_length + 1 samples (1..21 frames at 32 taps), output at the
    flush for 22..32 frames, the normal path from 33 on; longer filters (rates above 48 kHz) shift the thresholds."""
    gold = GOLD4[f"short_counts_{rate}"]
    got = [orc.swr_out_count(rate, 48000, n, True) for n in range(1, G.SHORT_MAX + 1)]
    assert got == gold.tolist()
    whole = [len(orc.swr_whole(G.case_input(orc, rate, 3, 2, n, 90), 3, rate, 48000, flush=True)[0]) for n in range(1, G.SHORT_MAX + 1)]
    assert whole == gold.tolist()
    if rate == 44100:
        assert gold[20] == 0 and gold[21] == 19 and gold[31] == 35 and gold[32] == 36      # 21 / 22 / 32 / 33 frames


@pytest.mark.parametrize("case", G.SHORT_VALUE_CASES, ids=[f"{c[0]}_{c[1]}_{c[2]}_{c[3]}" for c in G.SHORT_VALUE_CASES])
def test_values_of_streams_shorter_than_the_filter(orc, case):
    """the samples such a stream yields come from a buffer mirrored at BOTH ends (the initial mirror reads reflected
    tail samples): frame-by-frame feed (7-sample frames), values within 1e-6 of the real library"""
    rate, fmt, ch, n = case
    x = G.case_input(orc, rate, fmt, ch, n, 91)
    c = orc.Swr(rate, 48000, fmt, ch)
    L, R_ = [], []
    for fr in G.frames_of(x, fmt, 7):
        l, r = c.convert(fr, 4096)
        L.append(l); R_.append(r)
    for _ in range(4):
        l, r = c.convert(None, 4096)
        L.append(l); R_.append(r)
    l, r = np.concatenate(L), np.concatenate(R_)
    gl, gr = GOLD4[f"short_{rate}_{fmt}_{ch}_{n}_l"], GOLD4[f"short_{rate}_{fmt}_{ch}_{n}_r"]
    assert len(l) == len(gl) > 0
    tol = VALUE_TOL if rate != 47999 else 1e-5
    assert np.abs(l - gl).max() <= tol and np.abs(r - gr).max() <= tol


GOLD3 = np.load(os.path.join(HERE, "golden", "swr_real_amix_capped.npz"))


def test_amix_loop_where_capped_conversion_yields_one_more_sample(orc):
    """audio_amix with nb = 64 next to an 88.2 kHz input cut into 3561-sample frames: the real library hands the surplus
    out over 160 later calls and returns 10213 samples in total, one MORE than a conversion with ample capacity (10212).
    The oracle node and the host-side plan of the CUDA engine (nodey_amix_plan) follow the loop, not the whole-stream
    count: stream length and zero positions exact, values within 1e-6."""
    import nodey
    tag = "amix4_capped_surplus"
    spec, vols = G.AMIX_CAPPED_CASES[tag]
    assert int(GOLD3[f"{tag}_capped_total"]) == int(GOLD3[f"{tag}_whole_total"]) + 1 == 10213
    xs = [G.case_input(orc, r, f, c, n, 40 + i) for i, (r, f, c, n, fr) in enumerate(spec)]
    tracks = [orc.make_track(x, f, r, frame_size=fr) for x, (r, f, c, n, fr) in zip(xs, spec)]
    l, r_ = orc.amix(tracks, vols)
    gl, gr = GOLD3[f"{tag}_l"], GOLD3[f"{tag}_r"]
    assert len(l) == len(gl) == int(GOLD3[f"{tag}_nb"].sum())
    assert np.array_equal(l == 0, gl == 0) and np.array_equal(r_ == 0, gr == 0)
    assert np.abs(l - gl).max() <= VALUE_TOL and np.abs(r_ - gr).max() <= VALUE_TOL
    # whole-stream model of that input: one sample short of what the loop places
    whole, _ = orc.swr_whole(xs[2], orc.FMT_FLT, 88200, 48000, flush=True)
    assert len(whole) == 10212
    total, segs, runs = nodey.amix_plan([s[0] for s in spec], [nodey.uniform_runs(s[3], s[4]) for s in spec])
    assert total == len(gl) and sum(a * b for a, b in runs) == total
    assert [seg for seg in segs if seg[0] == 2] == [(2, 0, 0, 10213)]


def test_amix_capped_fixture_is_what_the_library_produces_now():
    from oracle import real_swr as R
    if not R.available():
        pytest.skip("no libswresample in this image")
    live = G.generate_amix_capped()
    for k in GOLD3.files:
        assert np.array_equal(live[k], GOLD3[k]), k


GOLD2 = np.load(os.path.join(HERE, "golden", "swr_real_bimix2.npz"))


@pytest.mark.parametrize("tag", sorted(G.BIMIX2_CASES))
def test_bimix_v2_loop_matches_real_library(orc, tag):
    """orc_bimix_v2 against an independent restatement of audio-bimix.cpp:536-875 run on real SwrContexts
    (tests/golden/make_swr_golden.py: bimix_v2_real): per-frame conversion with capacity 2 * nb and no flush, mono
    down-mix, END-time stamps, the pts aligner with its round()ed sample counts, tails with zeros on the other channel.
    Stream length, first pts and the position of every zero exact; values within 1e-6 (bit exact at 48 kHz)."""
    left, right, pl, pr = G.BIMIX2_CASES[tag]
    tl = orc.make_track(G.case_input(orc, *left[:4], 70), left[1], left[0], frame_size=left[4], pts0=pl)
    tr = orc.make_track(G.case_input(orc, *right[:4], 71), right[1], right[0], frame_size=right[4], pts0=pr)
    out, pts = orc.bimix_v2(tl, tr)
    gold = GOLD2[f"{tag}_out"]
    assert out.shape == gold.shape == (int(GOLD2[f"{tag}_sizes"].sum()), 2), "stream length"
    assert pts == float(GOLD2[f"{tag}_pts"]), "pts of the first frame"
    assert np.array_equal(out == 0, gold == 0), "silent regions (alignment)"
    exact = left[0] == 48000 and right[0] == 48000
    assert np.abs(out - gold).max() <= (0.0 if exact else VALUE_TOL)


def test_bimix_v2_fixture_is_what_the_library_produces_now():
    from oracle import real_swr as R
    if not R.available():
        pytest.skip("no libswresample in this image")
    live = G.generate_bimix2()
    for k in GOLD2.files:
        assert np.array_equal(live[k], GOLD2[k]), k


@pytest.mark.parametrize("case", G.PREVIEW_CASES, ids=[c[0] for c in G.PREVIEW_CASES])
def test_preview_conversion_matches_real_library(orc, case):
    tag, rate, fmt, ch, n, frame = case
    x = G.case_input(orc, rate, fmt, ch, n, 60 + G.PREVIEW_CASES.index(case))
    s = orc.Swr(rate, 48000, fmt, ch)
    outs, counts = [], []
    for fr in G.frames_of(x, fmt, frame):
        m = fr.shape[1] if fmt >= 5 else fr.shape[0]
        cap = int(np.float32(np.float32(m) / np.float32(rate)) * 48000 * 1.5)
        l, r = s.convert(fr, cap)
        outs.append(np.stack([l, r], 1)); counts.append(len(l))
    got = np.concatenate(outs)
    assert counts == GOLD[f"{tag}_counts"].tolist()
    ref = GOLD[f"{tag}_out"]
    assert got.shape == ref.shape
    if rate == 48000:
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    else:
        assert np.abs(got - ref).max() <= VALUE_TOL


# ---- live: the library itself, where the image has it -------------------------------------------------
def _real():
    from oracle import real_swr as R
    if not R.available():
        pytest.skip("no libswresample in this image")
    return R


def test_fixture_is_what_the_library_produces_now():
    R = _real()
    now = G.generate()
    for k in GOLD.files:
        if k == "meta":
            continue
        a, b = np.asarray(now[k]), GOLD[k]
        assert a.shape == b.shape and np.array_equal(a, b), k


@pytest.mark.parametrize("c_template", [True, False], ids=["c_template", "cpu_assembly"])
def test_live_library_both_code_paths(orc, c_template):
    """the assembly the host CPU selects (SSE / AVX / FMA3) sums in another order than the C template: both stay
    within the same bound of the oracle, with identical counts.  One divergence is the library's own: on the
    linearly interpolating path (47999 Hz: no exact rational within 1024 phases) the x86 assembly and the C template
    disagree by up to 8e-6 on the first ~26 output samples and nowhere else; the oracle follows the C template
    (1.2e-7), so against the assembly that case is held to the task's 1e-5 bar."""
    R = _real()
    R.force_c_path(c_template)
    try:
        for i, (rate, fmt, ch) in enumerate([(44100, 3, 2), (22050, 1, 1), (96000, 3, 2), (8000, 2, 2), (47999, 8, 2)]):
            x = G.case_input(orc, rate, fmt, ch, 20000, 70 + i)
            l, r, counts = R.whole(x, fmt, rate, 48000, ch, frame=1152)
            ol, orr, ocounts = oracle_frames(orc, x, fmt, rate, ch, 1152)
            assert counts == ocounts
            tol = 1e-5 if (rate == 47999 and not c_template) else VALUE_TOL
            assert np.abs(l - ol).max() <= tol and np.abs(r - orr).max() <= tol
            if tol != VALUE_TOL:
                assert np.abs(l - ol)[64:].max() <= VALUE_TOL and np.abs(r - orr)[64:].max() <= VALUE_TOL
    finally:
        R.force_c_path(False)


def test_live_library_defaults_are_the_modelled_ones():
    R = _real()
    s = R.RealSwr(44100, 48000, 3, 2)
    try:
        assert (s.option("filter_size"), s.option("phase_shift"), s.option("linear_interp"), s.option("exact_rational"),
                s.option("kaiser_beta")) == (32, 10, 1, 1, 9)
        assert s.option("cutoff", "double") == 0.0      # 0 = "pick": swr_init uses 0.97 for the swr engine
    finally:
        s.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_randomised_loops_against_the_library(orc, seed):
    """seeded random audio_amix / audio_bimix / audio_bimix_v2 configurations (all six formats, mono and stereo, nine
    rates, frame sizes from 1 to 4096, inputs shorter than the filter, late streams): the oracle nodes against the
    restated loops running on the REAL library -- length, first pts and zero positions exact, values within 1e-6.
    (600 such cases were run when this test was written; it found the capped-surplus regime and a too small output
    buffer in the oracle's Python wrapper, nothing in the C restatement.)"""
    import random
    from oracle import real_swr as R
    if not R.available():
        pytest.skip("no libswresample in this image")
    rnd = random.Random(seed)
    rates = [8000, 11025, 16000, 22050, 32000, 44100, 48000, 88200, 96000]

    def spec(edge):
        n = rnd.choice([1, 20, 33, 34, 100, rnd.randint(1, 3000)]) if edge else rnd.choice([40, 500, rnd.randint(34, 9000)])
        fs = rnd.choice([1, 7, 64, rnd.randint(1, 600)]) if edge else rnd.choice([576, 1024, 1152, 4096, rnd.randint(200, 3000)])
        return (rnd.choice(rates), rnd.choice([1, 2, 3, 6, 7, 8]), rnd.choice([1, 2]), n, fs)

    R.force_c_path(True)
    try:
        for it in range(8):
            edge = it % 2 == 1
            # audio_amix
            specs = [spec(edge) for _ in range(rnd.randint(1, 4))]
            vols = [rnd.choice([1.0, 0.5, 0.25, 0.8]) for _ in specs]
            gl, gr, _ = G.amix_real(R, orc, specs, vols)
            tracks = [orc.make_track(G.case_input(orc, r, f, c, n, 40 + i), f, r, frame_size=fr) for i, (r, f, c, n, fr) in enumerate(specs)]
            l, r_ = orc.amix(tracks, vols)
            assert len(l) == len(gl), ("amix length", specs)
            if len(l):
                assert np.array_equal(l == 0, gl == 0) and max(np.abs(l - gl).max(), np.abs(r_ - gr).max()) <= VALUE_TOL, ("amix", specs, vols)
            # audio_bimix and audio_bimix_v2 on one pair of inputs
            left, right = spec(edge), spec(edge)
            bias = rnd.choice([0.0, 0.3, -0.5, 1.0])
            gl, gr, _ = G.bimix_real(R, orc, left, right, bias)
            tl = orc.make_track(G.case_input(orc, *left[:4], 50), left[1], left[0], frame_size=left[4])
            tr = orc.make_track(G.case_input(orc, *right[:4], 51), right[1], right[0], frame_size=right[4])
            l, r_ = orc.bimix(tl, tr, bias)
            assert len(l) == len(gl), ("bimix length", left, right)
            if len(l):
                assert max(np.abs(l - gl).max(), np.abs(r_ - gr).max()) <= VALUE_TOL, ("bimix", left, right, bias)
            pl, pr = rnd.choice([0.0, 0.0001, 0.013, 0.2]), rnd.choice([0.0, 0.00002, 0.05])
            try:
                gold, gpts, _ = G.bimix_v2_real(R, orc, left, right, pl, pr)
            except ValueError:          # nothing came out of either resampler: np.concatenate of no frames
                gold, gpts = np.zeros((0, 2), np.float32), None
            tl = orc.make_track(G.case_input(orc, *left[:4], 70), left[1], left[0], frame_size=left[4], pts0=pl)
            tr = orc.make_track(G.case_input(orc, *right[:4], 71), right[1], right[0], frame_size=right[4], pts0=pr)
            out, pts = orc.bimix_v2(tl, tr)
            assert out.shape == gold.shape, ("bimix_v2 length", left, right, pl, pr)
            if len(out):
                assert pts == gpts and np.array_equal(out == 0, gold == 0) and np.abs(out - gold).max() <= VALUE_TOL, ("bimix_v2", left, right, pl, pr)
    finally:
        R.force_c_path(False)


@pytest.mark.parametrize("case", [
    # (rate, format, channels, frames, frame size) per input: what audio_input publishes for WAV files (libavformat's
    # 4096-byte packets: 1024 frames of 16-bit stereo, 512 of float stereo, 682 of 24-bit stereo, 2048 of 16-bit mono,
    # 1365 of 24-bit mono) -- mixed in one audio_amix, so that the output is cut by the SHORTER frames and the inputs with
    # longer frames are drained through a capped output, surplus buffered inside the resampler
    ("s16_stereo+flt_stereo", [(44100, 1, 2, 30000, 1024), (44100, 3, 2, 30500, 512)]),
    ("s24_stereo+s16_mono", [(48000, 2, 2, 20000, 682), (44100, 1, 1, 25000, 2048)]),
    ("flt_stereo+s24_mono+s16_stereo", [(22050, 3, 2, 9000, 512), (44100, 2, 1, 21000, 1365), (96000, 1, 2, 40000, 1024)]),
], ids=lambda c: c[0] if isinstance(c, tuple) else None)
def test_live_amix_of_wav_packet_sizes_against_the_library(orc, case):
    from oracle import real_swr as R
    if not R.available():
        pytest.skip("no libswresample in this image")
    _, specs = case
    vols = [0.5, 0.8, 0.25][:len(specs)]
    R.force_c_path(True)
    try:
        gl, gr, _ = G.amix_real(R, orc, specs, vols)
    finally:
        R.force_c_path(False)
    # the same inputs as amix_real makes for itself (track seeds 40 + i)
    tracks = [orc.make_track(G.case_input(orc, r, f, c, n, 40 + i), f, r, frame_size=fr) for i, (r, f, c, n, fr) in enumerate(specs)]
    l, r_ = orc.amix(tracks, vols)
    assert len(l) == len(gl), ("amix length", specs)
    assert np.array_equal(l == 0, gl == 0) and max(np.abs(l - gl).max(), np.abs(r_ - gr).max()) <= VALUE_TOL
