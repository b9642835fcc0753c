"""GPU parity: streaming nodes (gain, extraction, split, format conversion, mixers, merge, synth)
against the oracle, bit exact, through the C ABI."""
import numpy as np
import pytest

from helpers import ALL_FMTS, FMT_FLT, FMT_FLTP, FMT_S16, FMT_S32, assert_bit_equal, make_input, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nch", [1, 2])
@pytest.mark.parametrize("n", [0, 1, 5, 1152, 100003])
def test_synth_bit_exact(nd, orc, n, nch):
    if n == 0:
        return
    f, s = nd.synth(n, nch, 44100, track=3, frame0=12345, want_s16=True)
    ref = orc.synth_f32(n, nch, 44100, 3, 12345)
    assert_bit_equal(f.cpu().numpy(), ref, "synth f32")
    assert_bit_equal(s.cpu().numpy(), orc.f32_to_s16(ref), "synth s16")


@pytest.mark.parametrize("fmt", ALL_FMTS)
@pytest.mark.parametrize("volume", [0.0, 0.5, 0.8, 1.0, 3.7, 10.0])
@pytest.mark.parametrize("n", [1, 7, 1152, 65537])
def test_gain_bit_exact(nd, orc, fmt, volume, n):
    x = make_input(orc, fmt, n, 2)
    if x.dtype == np.int32:
        x[: min(4, x.size // 2)].flat[:4] = [2147483647, -2147483648, 1 << 30, -(1 << 30)][: min(4, x.size)]
    ref = orc.gain(x, fmt, volume)
    got = nd.gain(to_dev(x), fmt, volume).cpu().numpy()
    assert_bit_equal(got, ref, f"gain fmt={fmt} v={volume}")


def test_gain_unaligned_and_errors(nd, orc):
    x = make_input(orc, FMT_FLT, 1001, 2)
    d = to_dev(x).reshape(-1)
    got = nd.gain(d[1:], FMT_FLT, 0.3).cpu().numpy()     # 4-byte aligned only
    assert_bit_equal(got, orc.gain(x.reshape(-1)[1:], FMT_FLT, 0.3), "gain unaligned")
    with pytest.raises(nd.NodeyError) as e:
        nd.gain(d, 4, 1.0)      # AV_SAMPLE_FMT_DBL: the reference throws "format is not support"
    assert e.value.code == -2


@pytest.mark.parametrize("fmt", ALL_FMTS)
@pytest.mark.parametrize("nch", [1, 2])
@pytest.mark.parametrize("n", [1, 1152, 40001])
def test_extract_interleaved_bit_exact(nd, orc, fmt, nch, n):
    x = make_input(orc, fmt, n, nch)
    if x.dtype != np.float32:
        info = np.iinfo(x.dtype)
        x.flat[:2] = [info.max, info.min][: min(2, x.size)]
    ref = orc.extract_interleaved(x, fmt)
    got = nd.extract_interleaved(to_dev(x), fmt).cpu().numpy()
    assert_bit_equal(got, ref, f"extract fmt={fmt}")


@pytest.mark.parametrize("fmt", ALL_FMTS)
@pytest.mark.parametrize("n", [1, 3, 1152, 99999])
def test_split_bit_exact(nd, orc, fmt, n):
    x = make_input(orc, fmt, n, 2)
    rl, rr = orc.split(x, fmt)
    gl, gr = nd.split(to_dev(x), fmt)
    assert_bit_equal(gl.cpu().numpy(), rl, "split L")
    assert_bit_equal(gr.cpu().numpy(), rr, "split R")


@pytest.mark.parametrize("fmt", ALL_FMTS)
@pytest.mark.parametrize("nch", [1, 2])
def test_to_fltp_matches_swr_passthrough(nd, orc, fmt, nch):
    n = 5000
    x = make_input(orc, fmt, n, nch, rate=48000)
    rl, rr = orc.swr_whole(x, fmt, 48000, 48000)
    got = nd.to_fltp_stereo(to_dev(x), fmt).cpu().numpy()
    assert_bit_equal(got[0], rl, "fltp L")
    assert_bit_equal(got[1], rr, "fltp R")


@pytest.mark.parametrize("nin", [1, 2, 5, 16])
def test_mix_bit_exact_order_and_ragged(nd, orc, nin):
    rng = np.random.default_rng(nin)
    lens = [int(rng.integers(1, 30000)) for _ in range(nin)]
    lens[0] = 30001
    vols = rng.uniform(0, 2, nin).astype(np.float32)
    ins = [orc.synth_f32(l, 2, 48000, i).T.copy() for i, l in enumerate(lens)]
    n = max(lens)
    rl = np.zeros(n, np.float32); rr = np.zeros(n, np.float32)
    for x, v in zip(ins, vols):     # audio-amix.cpp:296-307: temp += data * volume in input order
        pl = np.zeros(n, np.float32); pr = np.zeros(n, np.float32)
        pl[: x.shape[1]] = x[0]; pr[: x.shape[1]] = x[1]
        rl = (rl + pl * v).astype(np.float32); rr = (rr + pr * v).astype(np.float32)
    got = nd.mix([to_dev(x) for x in ins], vols).cpu().numpy()
    assert_bit_equal(got[0], rl, "mix L")
    assert_bit_equal(got[1], rr, "mix R")


def test_mix_rejects_17_inputs(nd, orc):
    x = to_dev(orc.synth_f32(16, 2, 48000, 0).T.copy())
    with pytest.raises(nd.NodeyError) as e:
        nd.mix([x] * 17, [1.0] * 17)
    assert e.value.code == -5


@pytest.mark.parametrize("bias", [-1.0, -0.3, 0.0, 0.25, 1.0])
def test_bimix_bit_exact(nd, orc, bias):
    a = orc.synth_f32(20001, 2, 48000, 1).T.copy()
    b = orc.synth_f32(15000, 2, 48000, 2).T.copy()
    n = 20001
    bp = np.zeros((2, n), np.float32); bp[:, :15000] = b
    bm, bpl = np.float32(1) - np.float32(bias), np.float32(1) + np.float32(bias)
    rl = ((a[0] / np.float32(2) + a[1] / np.float32(2)) * bm).astype(np.float32)
    rr = ((bp[0] / np.float32(2) + bp[1] / np.float32(2)) * bpl).astype(np.float32)
    got = nd.bimix(to_dev(a), to_dev(b), bias).cpu().numpy()
    assert_bit_equal(got[0], rl, "bimix L")
    assert_bit_equal(got[1], rr, "bimix R")


def test_downmix_and_merge_segments(nd, orc):
    a = orc.synth_f32(9001, 2, 48000, 4).T.copy()
    ref = ((a[0] + a[1]) * np.float32(0.5)).astype(np.float32)
    got = nd.downmix_half(to_dev(a))
    assert_bit_equal(got.cpu().numpy(), ref, "downmix")
    left = ref; right = orc.synth_f32(7000, 1, 48000, 5)[:, 0].copy()
    segs = [(0, 1000, 0, -1), (1000, 6000, 1000, 0), (7000, 1000, -1, 6000), (8000, 500, 7000, -1)]
    out = nd.merge_segments(to_dev(left), to_dev(right), segs).cpu().numpy()
    exp = np.zeros((8500, 2), np.float32)
    for o, n, l, r in segs:
        if l >= 0: exp[o:o + n, 0] = left[l:l + n]
        if r >= 0: exp[o:o + n, 1] = right[r:r + n]
    assert_bit_equal(out, exp, "merge")


def test_gain_tracks_batch_ragged(nd, orc):
    """nodey_gain_tracks: ragged float streams (one empty, one unaligned) in one launch == the oracle gain per stream"""
    import ctypes as C
    import torch
    lens = [0, 1, 7, 1152, 100003, 4096]
    vols = [0.5, 2.0, 0.25, 1.0, 0.8, 0.0]
    xs = [orc.synth_f32(max(n, 1), 1, 48000, 40 + i)[:n, 0].copy() for i, n in enumerate(lens)]
    srcs = [torch.from_numpy(np.concatenate([np.zeros(1, np.float32), x])).cuda()[1:] if i == 4 else torch.from_numpy(x).cuda() for i, x in enumerate(xs)]
    dsts = [torch.empty_like(s) for s in srcs]
    k = len(lens)
    pd = (C.c_void_p * k)(*[d.data_ptr() if d.numel() else 0 for d in dsts])
    ps = (C.c_void_p * k)(*[s.data_ptr() if s.numel() else 0 for s in srcs])
    pn = (C.c_int64 * k)(*lens)
    pv = (C.c_float * k)(*vols)
    nd.check(nd.lib().nodey_gain_tracks(pd, ps, pn, pv, 3, k, None))
    torch.cuda.synchronize()
    for i in range(k):
        ref = orc.gain(xs[i].reshape(-1, 1), 3, vols[i]).reshape(-1) if lens[i] else np.zeros(0, np.float32)
        assert_bit_equal(dsts[i].cpu().numpy(), ref, f"stream {i}")
    with pytest.raises(nd.NodeyError):
        nd.check(nd.lib().nodey_gain_tracks(pd, ps, pn, pv, 1, k, None))        # integer formats go through nodey_gain
