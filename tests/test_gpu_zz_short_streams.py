"""GPU (named to run last): streams a little shorter than the resampler's filter.

libswresample's invert_initial_buffer() waits for filter_length + 1 samples, and the samples resample_flush() reflects
count: at 32 taps a stream of 22..32 frames produces its output at the flush, from a buffer that is mirrored at both
ends (the initial mirror reads reflected tail samples).  A randomised comparison with the real library found that the
oracle and the CUDA plan both returned nothing for such streams; both were corrected (oracle/nodey_oracle.c:
swr_producible / flush, csrc/resample.cu: plan_producible / plan_reflect -- the same two lines).  The oracle and the
host-side counts were verified on the CPU against the real library (tests/test_swr_real.py, tests/golden/
swr_real_short.npz).  This file was written after the round's GPU budget was spent: it is the first check of the
KERNEL's values in that regime (the kernels index the extended signal like the oracle does: mirror about sample 0,
then reflection at the end), and it runs last so that a surprise here cannot hide any other result."""
import os
import sys

import numpy as np
import pytest

from helpers import FMT_FLT, assert_bit_equal, make_input, to_dev

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("rate,n", [(44100, 22), (44100, 27), (44100, 32), (22050, 30), (96000, 50), (96000, 66)])
def test_short_stream_resample_matches_the_oracle(nd, orc, rate, n):
    x = make_input(orc, FMT_FLT, n, 2, rate=rate)
    rl, rr = orc.swr_whole(x, FMT_FLT, rate, 48000, flush=True)
    r = nd.Resampler(rate, 48000)
    assert len(rl) > 0 and r.out_count(n, True) == len(rl)
    assert r.out_count(n, False) == 0                      # nothing before the flush
    got = r.run(to_dev(x), FMT_FLT, flush=True).cpu().numpy()
    assert got.shape[1] == len(rl)
    assert_bit_equal(got[0], rl, f"short stream L {rate} {n}")
    assert_bit_equal(got[1], rr, f"short stream R {rate} {n}")


@pytest.mark.parametrize("case", [(44100, 3, 2, 22), (44100, 3, 2, 27), (96000, 3, 2, 50)], ids=lambda c: f"{c[0]}_{c[3]}")
def test_short_stream_resample_matches_the_real_library(nd, orc, case):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_swr_golden as G
    gold = np.load(os.path.join(HERE, "golden", "swr_real_short.npz"))
    rate, fmt, ch, n = case
    x = G.case_input(orc, rate, fmt, ch, n, 91)
    got = nd.Resampler(rate, 48000).run(to_dev(x), fmt, flush=True).cpu().numpy()
    gl, gr = gold[f"short_{rate}_{fmt}_{ch}_{n}_l"], gold[f"short_{rate}_{fmt}_{ch}_{n}_r"]
    assert got.shape == (2, len(gl))
    assert np.abs(got[0] - gl).max() <= 1e-6 and np.abs(got[1] - gr).max() <= 1e-6


def test_amix_of_a_30_frame_source(eng_gpu, orc):
    """the node level: audio_amix(1) of a 30-frame 44.1 kHz stream emits its 32 resampled samples in the flush iterations"""
    x = make_input(orc, FMT_FLT, 30, 2, track=5)
    p = eng_gpu.Project()
    src = p.add("audio_input", {"file_path": [""]})
    mix = p.add("audio_amix", eng_gpu.amix_info([0.5]))
    out = p.add("audio_output")
    p.link(src, "output_0", mix, "input_1"); p.link(mix, "output", out, "input")
    e = eng_gpu.Engine(p.json())
    e.bind_source(0, x, FMT_FLT, 44100)
    e.run()
    rl, rr = orc.amix([orc.make_track(x, FMT_FLT, 44100)], [0.5])
    assert np.count_nonzero(rl) > 0
    assert_bit_equal(e.output().numpy(), np.stack([rl, rr]), "amix of a 30-frame stream")
    e.close()
