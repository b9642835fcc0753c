"""GPU parity: polyphase resampler (A7) and its fusion with the N-input mix (A4) against the
oracle's libswresample model, bit exact (same tap order, fused multiply-add) through the C ABI."""
import numpy as np
import pytest

from helpers import ALL_FMTS, FMT_FLT, FMT_FLTP, FMT_S16, assert_bit_equal, make_input, to_dev

pytestmark = pytest.mark.gpu


def test_plan_matches_oracle(nd, orc):
    for in_rate, out_rate in [(44100, 48000), (48000, 44100), (22050, 48000), (96000, 48000), (8000, 48000), (44099, 48000)]:
        for quirk in (0, 1):
            r = nd.Resampler(in_rate, out_rate, quirk)
            o = orc.Swr(in_rate, out_rate, FMT_FLT, 2, quirk)
            assert r.info() == o.plan(), (in_rate, out_rate)
            assert_bit_equal(r.filter_bank(), o.filter_bank(), "filter bank")
            for n in (0, 10, 33, 34, 1000, 44100):
                for flush in (False, True):
                    assert r.out_count(n, flush) == orc.swr_out_count(in_rate, out_rate, n, flush, quirk)
            r.close()


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("fmt", ALL_FMTS)
@pytest.mark.parametrize("nch", [1, 2])
def test_resample_441_to_48_bit_exact(nd, orc, mode, fmt, nch):
    n = 30011
    x = make_input(orc, fmt, n, nch)
    rl, rr = orc.swr_whole(x, fmt, 44100, 48000, flush=True, quirk=0)
    r = nd.Resampler(44100, 48000)
    got = r.run(to_dev(x), fmt, flush=True, mode=mode).cpu().numpy()
    assert_bit_equal(got[0], rl, "resample L")
    assert_bit_equal(got[1], rr, "resample R")


@pytest.mark.parametrize("rates", [(48000, 44100), (22050, 48000), (96000, 48000), (32000, 48000), (44099, 48000), (11025, 48000)])
@pytest.mark.parametrize("flush", [False, True])
def test_resample_other_rates_bit_exact(nd, orc, rates, flush):
    n = 20000
    x = make_input(orc, FMT_FLT, n, 2, rate=rates[0])
    rl, rr = orc.swr_whole(x, FMT_FLT, rates[0], rates[1], flush=flush, quirk=0)
    r = nd.Resampler(*rates)
    got = r.run(to_dev(x), FMT_FLT, flush=flush).cpu().numpy()
    assert got.shape[1] == len(rl)
    assert_bit_equal(got[0], rl, "resample L")
    assert_bit_equal(got[1], rr, "resample R")


# 21 frames: the longest stream that never produces anything (even its flush reflection stays below filter_length + 1
# samples); 22..32 frames produce output at the flush -- that regime has its own file, tests/test_gpu_zz_short_streams.py
@pytest.mark.parametrize("n", [0, 1, 21, 33, 34, 100, 147, 4704, 4705])
def test_resample_short_inputs(nd, orc, n):
    if n == 0:
        r = nd.Resampler(44100, 48000)
        assert r.out_count(0, True) == 0
        return
    x = make_input(orc, FMT_FLT, n, 2)
    rl, rr = orc.swr_whole(x, FMT_FLT, 44100, 48000, flush=True, quirk=0)
    r = nd.Resampler(44100, 48000)
    assert r.out_count(n, True) == len(rl)
    if len(rl):
        for mode in (1, 2, 3, 4):
            got = r.run(to_dev(x), FMT_FLT, flush=True, mode=mode).cpu().numpy()
            assert_bit_equal(got[0], rl, f"short L mode {mode}")
            assert_bit_equal(got[1], rr, f"short R mode {mode}")


def test_resample_quirk_phase(nd, orc):
    x = make_input(orc, FMT_FLT, 5000, 2)
    rl, rr = orc.swr_whole(x, FMT_FLT, 44100, 48000, flush=True, quirk=1)
    r = nd.Resampler(44100, 48000, quirk=1)
    got = r.run(to_dev(x), FMT_FLT, flush=True).cpu().numpy()
    assert_bit_equal(got[0], rl, "quirk L")


def test_resample_rejects_too_many_frames(nd, orc):
    x = to_dev(make_input(orc, FMT_FLT, 1000, 2))
    r = nd.Resampler(44100, 48000)
    with pytest.raises(nd.NodeyError) as e:
        r.run(x, FMT_FLT, flush=True, out_frames=r.out_count(1000, True) + 1)
    assert e.value.code == -5


@pytest.mark.parametrize("nin", [1, 2, 16])
def test_amix_fused_matches_oracle_node(nd, orc, nin):
    """audio_amix restated frame by frame (orc_amix) == fused resample + ordered mix on the GPU.
    The oracle output is zero padded to whole 1152-frame iterations; the common prefix must be
    bit identical and everything after the longest resampled input must be zero."""
    rng = np.random.default_rng(7 + nin)
    lens = [int(rng.integers(20000, 40000)) for _ in range(nin)]
    vols = (rng.uniform(0.1, 1.0, nin)).astype(np.float32)
    xs = [make_input(orc, FMT_FLT, l, 2, track=i) for i, l in enumerate(lens)]
    tracks = [orc.make_track(x, FMT_FLT, 44100) for x in xs]
    rl, rr = orc.amix(tracks, vols, quirk=0)
    r = nd.Resampler(44100, 48000)
    got = r.resample_mix([to_dev(x) for x in xs], [FMT_FLT] * nin, vols, flush=True).cpu().numpy()
    m = got.shape[1]
    assert m <= len(rl)
    assert_bit_equal(got[0], rl[:m], "amix L")
    assert_bit_equal(got[1], rr[:m], "amix R")
    assert not rl[m:].any() and not rr[m:].any()


def test_amix_mixed_formats_unfused(nd, orc):
    """inputs at different rates/formats: per-input resample, then the ordered mix kernel."""
    specs = [(FMT_S16, 44100, 2), (FMT_FLTP, 48000, 2), (FMT_FLT, 22050, 1), (FMT_S16, 44100, 1)]
    xs = [make_input(orc, f, 15000, ch, rate=sr, track=i) for i, (f, sr, ch) in enumerate(specs)]
    vols = np.array([0.9, 0.5, 0.7, 0.2], np.float32)
    tracks = [orc.make_track(x, f, sr) for x, (f, sr, ch) in zip(xs, specs)]
    rl, rr = orc.amix(tracks, vols, quirk=0)
    outs = []
    for x, (f, sr, ch) in zip(xs, specs):
        r = nd.Resampler(sr, 48000)
        outs.append(r.run(to_dev(x), f, flush=True))
    got = nd.mix(outs, vols).cpu().numpy()
    m = got.shape[1]
    assert_bit_equal(got[0], rl[:m], "amix mixed L")
    assert_bit_equal(got[1], rr[:m], "amix mixed R")
    assert not rl[m:].any()


@pytest.mark.parametrize("fmt", [FMT_FLT, FMT_FLTP, FMT_S16])
@pytest.mark.parametrize("nch", [1, 2])
def test_resample_tracks_batch_matches_single_mixers(nd, orc, fmt, nch):
    """nodey_resample_tracks: a batch of audio_amix(1) resamplers in one launch == the oracle's amix per track"""
    n = 61013                      # several tiles per track, ragged last tile
    ntr = 5
    vols = np.array([1.0, 0.5, 0.25, 0.9, 0.1], np.float32)
    xs = [make_input(orc, fmt, n, nch, track=i) for i in range(ntr)]
    r = nd.Resampler(44100, 48000)
    got = r.resample_tracks([to_dev(x) for x in xs], fmt, vols, flush=True).cpu().numpy()
    for t in range(ntr):
        rl, rr = orc.amix([orc.make_track(xs[t], fmt, 44100)], vols[t:t + 1], quirk=0)
        m = got.shape[2]
        assert m <= len(rl)
        assert_bit_equal(got[t, 0], rl[:m], f"track {t} L")
        assert_bit_equal(got[t, 1], rr[:m], f"track {t} R")
        assert not rl[m:].any() and not rr[m:].any()


@pytest.mark.parametrize("fmt,nch,nchunks", [(FMT_FLT, 2, 2), (FMT_FLT, 2, 8), (FMT_FLTP, 2, 5), (FMT_S16, 1, 3), (FMT_FLT, 2, 64)])
def test_resample_tracks_in_chunks_is_bit_identical(nd, orc, fmt, nch, nchunks):
    """nodey_resample_tracks_chunk: the batch resampler cut into launches along time.  Same bits as the one-launch call;
    chunk c reads nothing beyond in_need[c] (everything past it is overwritten before the chunk runs and restored after)
    and frames below out_ready[c] are final after it."""
    import torch
    n = 44100 * 9 + 321
    ntr = 3
    vols = np.array([1.0, 0.5, 0.3], np.float32)
    xs = [to_dev(make_input(orc, fmt, n, nch, track=i)) for i in range(ntr)]
    r = nd.Resampler(44100, 48000)
    whole = r.resample_tracks(xs, fmt, vols, flush=True)
    keep = [x.clone() for x in xs]
    planar = fmt >= 5
    state = {"out": None}

    def poison(c, in_need):
        for x, k in zip(xs, keep):
            x.copy_(k)
            if x.dtype.is_floating_point:
                (x[..., in_need:] if planar else x[in_need:]).fill_(float("nan"))
            else:
                (x[..., in_need:] if planar else x[in_need:]).fill_(32767)

    got, plan = r.resample_tracks_chunked(xs, fmt, vols, nchunks, flush=True, poison=poison)
    torch.cuda.synchronize()
    assert 1 <= len(plan) <= nchunks and plan[-1] == (n, whole.shape[2])
    assert all(plan[k][0] <= plan[k + 1][0] and plan[k][1] < plan[k + 1][1] for k in range(len(plan) - 1))
    if nchunks <= 5:
        assert len(plan) == nchunks and plan[0][0] < n
    assert torch.equal(got.view(torch.int32), whole.view(torch.int32)), "chunked batch resampler differs from the one-launch result"


def test_resample_tracks_rejects_plans_without_pipelined_kernel(nd, orc):
    x = to_dev(make_input(orc, FMT_FLT, 5000, 2, rate=22050))
    r = nd.Resampler(22050, 48000)            # 320 phases: two groups per warp -> staging-row kernel only
    with pytest.raises(nd.NodeyError) as e:
        r.resample_tracks([x], FMT_FLT, [1.0])
    assert e.value.code == -5


# ---- against the REAL libswresample (tests/golden/swr_real.npz, see tests/test_swr_real.py) -----------------
def _swr_gold():
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_swr_golden as G
    return G, np.load(os.path.join(here, "golden", "swr_real.npz"))


def test_cuda_impulse_response_is_the_real_filter_bank(nd):
    """a unit impulse through the CUDA resampler returns libswresample's own coefficients, bit for bit"""
    G, gold = _swr_gold()
    for rate in G.IMPULSE_RATES:
        x = np.zeros((4000, 2), np.float32)
        x[2000, 0] = 1.0
        r = nd.Resampler(rate, 48000)
        got = r.run(to_dev(x), FMT_FLT, flush=True).cpu().numpy()
        r.close()
        assert got.shape[1] == int(gold[f"impulse_{rate}_len"]), rate
        first, taps = int(gold[f"impulse_{rate}_first"]), gold[f"impulse_{rate}_taps"]
        nz = np.flatnonzero(got[0])
        assert nz[0] == first and nz[-1] == first + len(taps) - 1, rate
        assert_bit_equal(got[0][first:first + len(taps)], taps, f"impulse response {rate}")
        assert not got[1].any()


def test_cuda_values_match_real_library(nd, orc):
    """same inputs as the fixture: stream length exact, samples within 1e-6 absolute of the library's C template
    (the task's bar is 1e-5; the residual is summation order), equal-rate conversions bit exact"""
    G, gold = _swr_gold()
    for i, (tag, rate, fmt, ch, n, frame) in enumerate(G.VALUE_CASES):
        x = G.case_input(orc, rate, fmt, ch, n, 20 + i)
        gl, gr = gold[f"{tag}_l"], gold[f"{tag}_r"]
        r = nd.Resampler(rate, 48000)
        assert r.out_count(n, True) == len(gl), tag
        if len(gl) == 0:
            r.close()
            continue
        got = r.run(to_dev(x), fmt, flush=True).cpu().numpy()
        r.close()
        assert got.shape == (2, len(gl)), tag
        if rate == 48000:
            assert_bit_equal(got[0], gl, tag); assert_bit_equal(got[1], gr, tag)
        else:
            assert np.abs(got[0] - gl).max() <= 1e-6 and np.abs(got[1] - gr).max() <= 1e-6, tag
