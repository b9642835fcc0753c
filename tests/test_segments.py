"""CPU: time-segment sharding of one long stream (bindings/segments.py).  The host arithmetic that
tells a rank which input slice it needs is checked against the oracle: every rank's outputs, computed from its
slice alone, concatenate to the whole-stream result bit for bit -- resampler (one period of lead-in), spectrum
(nfft - hop halo), streaming nodes (no halo).  One case runs as two real ranks over gloo."""
import os
import sys

import numpy as np
import pytest

from helpers import FMT_FLT, FMT_S16, assert_bit_equal, make_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _resample_rank(orc, seg_mod, x, fmt, in_rate, world, rank):
    n = x.shape[0]
    plan = orc.Swr(in_rate, 48000, fmt, x.shape[1]).plan()
    total = orc.swr_out_count(in_rate, 48000, n, True)
    seg = seg_mod.resample_ranges(plan, n, total, world, rank)
    if seg["k1"] <= seg["k0"]:
        return seg, np.zeros((2, 0), np.float32)
    l, r = orc.swr_whole(x[seg["in0"]:seg["in1"]], fmt, in_rate, 48000, flush=seg["flush"])
    want = seg["skip"] + seg["k1"] - seg["k0"]
    assert len(l) >= want, (len(l), want, seg)
    return seg, np.stack([l[seg["skip"]:want], r[seg["skip"]:want]])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("case", [(44100, FMT_FLT, 2, 100003), (44100, FMT_S16, 1, 50000), (22050, FMT_FLT, 2, 40001), (32000, FMT_FLT, 2, 30000)])
def test_resampler_segments_concatenate_to_whole_stream(orc, world, case):
    import segments
    in_rate, fmt, nch, n = case
    x = make_input(orc, fmt, n, nch, rate=in_rate)
    wl, wr = orc.swr_whole(x, fmt, in_rate, 48000, flush=True)
    parts, covered = [], 0
    for rank in range(world):
        seg, out = _resample_rank(orc, segments, x, fmt, in_rate, world, rank)
        assert seg["k0"] == covered and seg["in1"] <= n
        covered = seg["k1"]
        parts.append(out)
        if world > 1 and 0 < rank < world - 1 and out.shape[1]:
            assert seg["in1"] - seg["in0"] < n, "an inner rank must not need the whole stream"
    assert covered == len(wl)
    got = np.concatenate(parts, axis=1)
    assert_bit_equal(got[0], wl, "segments L")
    assert_bit_equal(got[1], wr, "segments R")


def test_resampler_segment_rejects_bad_requests(orc):
    import segments
    plan = orc.Swr(44100, 48000, FMT_FLT, 2).plan()
    with pytest.raises(ValueError):
        segments.resample_segment(plan, 10000, 10880, 100, 5000)          # not on a period boundary
    with pytest.raises(ValueError):
        segments.resample_segment(orc.Swr(44099, 48000, FMT_FLT, 2).plan(), 10000, 10000, 0, 100)   # interpolating plan


@pytest.mark.parametrize("world", [1, 2, 5])
def test_stft_and_stream_segments(orc, world):
    import segments
    n = 48000 + 77
    x = orc.synth_f32(n, 1, 48000, 3)[:, 0]
    whole = orc.stft(x)
    parts = []
    for rank in range(world):
        seg = segments.stft_segment(n, world, rank)
        if seg["m1"] > seg["m0"]:
            assert seg["in1"] <= n
            parts.append(orc.stft(x[seg["in0"]:seg["in1"]]))
            assert parts[-1].shape[0] == seg["m1"] - seg["m0"]
    assert_bit_equal(np.concatenate(parts), whole, "stft segments")
    # streaming node: any cut
    y = np.concatenate([orc.gain(x[a:b], FMT_FLT, 0.37) for a, b in (segments.stream_segment(n, world, r) for r in range(world))])
    assert_bit_equal(y, orc.gain(x, FMT_FLT, 0.37), "gain segments")
    assert segments.stft_segment(100, world, 0)["m1"] == 0                # shorter than one frame: nothing to do


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import segments
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # config 3 in small: 4 tracks 44.1 kHz -> 48 kHz, mixed; every rank renders its time segment of the bus
    tracks = [O.synth_f32(n, 2, 44100, t) for t in range(4)]
    plan = O.Swr(44100, 48000, 3, 2).plan()
    total = O.swr_out_count(44100, 48000, n, True)
    seg = segments.resample_ranges(plan, n, total, world, rank)
    acc = np.zeros((2, seg["k1"] - seg["k0"]), np.float32)
    want = seg["skip"] + seg["k1"] - seg["k0"]
    for x in tracks:
        l, r = O.swr_whole(x[seg["in0"]:seg["in1"]], 3, 44100, 48000, flush=seg["flush"])
        acc[0] = acc[0] + l[seg["skip"]:want] * np.float32(0.25)           # audio_amix: temp += data * volume, input order
        acc[1] = acc[1] + r[seg["skip"]:want] * np.float32(0.25)
    # no data-path collective: segments are gathered only to be compared (a render writes them to disjoint file ranges)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([acc.shape[1]]))
    width = int(max(s.item() for s in sizes))
    pad = torch.zeros((2, width)); pad[:, :acc.shape[1]] = torch.from_numpy(acc)
    got = [torch.zeros((2, width)) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, got, dst=0)
    if rank == 0:
        np.save(out_path, np.concatenate([g.numpy()[:, :int(s.item())] for g, s in zip(got, sizes)], axis=1))
    dist.destroy_process_group()


def test_two_ranks_render_time_segments_of_one_mix(tmp_path, orc):
    import torch.multiprocessing as mp
    n = 66150
    out = str(tmp_path / "bus.npy")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = np.load(out)
    acc = None
    for t in range(4):
        l, r = orc.swr_whole(orc.synth_f32(n, 2, 44100, t), FMT_FLT, 44100, 48000, flush=True)
        cur = np.stack([l, r]) * np.float32(0.25)
        acc = (np.zeros_like(cur) + cur) if acc is None else acc + cur
    assert_bit_equal(got.astype(np.float32), acc, "two-rank time-segment mix")


@pytest.mark.parametrize("seed", [1, 2])
def test_random_streams_and_rank_counts_concatenate_bit_exact(orc, seed):
    """seeded random rates, formats, lengths (including streams shorter than the filter and streams of a few periods)
    cut over 1..8 ranks: every rank's slice alone gives its outputs, the concatenation is the whole-stream result.
    (A slice at the end of a tiny stream used to be shorter than the filter_length + 1 frames a conversion needs to start.)"""
    import random
    import segments
    rnd = random.Random(seed)
    for _ in range(40):
        in_rate = rnd.choice([8000, 11025, 16000, 22050, 32000, 44100, 88200, 96000])
        fmt, nch = rnd.choice([FMT_FLT, FMT_S16]), rnd.choice([1, 2])
        n = rnd.choice([1, 21, 22, 30, 33, 147, 148, 1000, rnd.randint(1, 30000)])
        world = rnd.choice([1, 2, 3, 4, 7, 8])
        x = make_input(orc, fmt, n, nch, rate=in_rate)
        wl, wr = orc.swr_whole(x, fmt, in_rate, 48000, flush=True)
        parts, covered = [], 0
        for rank in range(world):
            seg, out = _resample_rank(orc, segments, x, fmt, in_rate, world, rank)
            assert seg["k0"] == covered and seg["in1"] <= n, (in_rate, n, world, rank, seg)
            covered = seg["k1"]
            parts.append(out)
        assert covered == len(wl), (in_rate, n, world)
        got = np.concatenate(parts, axis=1)
        assert_bit_equal(got[0], wl, f"segments L {in_rate} {n} {world}")
        assert_bit_equal(got[1], wr, f"segments R {in_rate} {n} {world}")
