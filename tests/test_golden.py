"""Golden vectors (tests/golden/oracle_r1.json, made by tests/golden/make_golden.py): the oracle must
reproduce them on CPU, the CUDA path must reproduce the same digests on the GPU."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_r1.json")))


def test_oracle_reproduces_golden(orc):
    import make_golden
    now = make_golden.cases(orc)
    assert sorted(now) == sorted(GOLD)
    for k in GOLD:
        assert now[k]["shape"] == GOLD[k]["shape"], k
        assert now[k]["sha256"] == GOLD[k]["sha256"], k


@pytest.mark.gpu
def test_cuda_reproduces_golden(nd, orc):
    import make_golden
    from helpers import make_input, to_dev
    d = make_golden.digest
    x = orc.synth_f32(5000, 2, 44100, 7)
    f, s = nd.synth(5000, 2, 44100, track=7, want_s16=True)
    assert d(f.cpu().numpy())["sha256"] == GOLD["synth_f32_t7"]["sha256"]
    assert d(s.cpu().numpy())["sha256"] == GOLD["synth_s16_t7"]["sha256"]
    for fmt in (1, 2, 3, 6, 7, 8):
        xi = make_input(orc, fmt, 3001, 2, track=fmt)
        assert d(nd.gain(to_dev(xi), fmt, 0.8).cpu().numpy())["sha256"] == GOLD[f"gain_fmt{fmt}_v0.8"]["sha256"]
        assert d(nd.extract_interleaved(to_dev(xi), fmt).cpu().numpy())["sha256"] == GOLD[f"extract_fmt{fmt}"]["sha256"]
    r = nd.Resampler(44100, 48000)
    assert d(r.run(to_dev(x), nd.FMT_FLT).cpu().numpy())["sha256"] == GOLD["swr_44100_48000_flt"]["sha256"]
    xs = make_input(orc, nd.FMT_S16, 4000, 1, rate=22050, track=2)
    r2 = nd.Resampler(22050, 48000)
    assert d(r2.run(to_dev(xs), nd.FMT_S16).cpu().numpy())["sha256"] == GOLD["swr_22050_48000_s16_mono"]["sha256"]
    x48 = orc.synth_f32(48000, 2, 48000, 5)
    st = nd.SoundTouch(48000, 2, 1.0, orc.pitch_node_factor(3.0))
    y, offs = st.run(to_dev(x48), want_offsets=True)
    assert d(y.cpu().numpy())["sha256"] == GOLD["soundtouch_pitch3"]["sha256"]
    assert d(offs.cpu().numpy())["sha256"] == GOLD["soundtouch_pitch3_offsets"]["sha256"]
    st2 = nd.SoundTouch(48000, 2, 1.25, orc.velocity_node_pitch(1.25, True))
    y2, offs2 = st2.run(to_dev(x48), want_offsets=True)
    assert d(y2.cpu().numpy())["sha256"] == GOLD["soundtouch_tempo1.25"]["sha256"]
    st3 = nd.SoundTouch(48000, 1, 1.0, orc.pitch_node_factor(-4.0))
    assert d(st3.run(to_dev(x48[:, :1].copy())).cpu().numpy())["sha256"] == GOLD["soundtouch_mono_pitch-4"]["sha256"]
