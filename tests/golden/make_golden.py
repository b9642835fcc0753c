"""Generates tests/golden/oracle_r1.json: digests and excerpts of the oracle's outputs on seeded
synthetic inputs.  The reference ships no fixtures and cannot be built or imported here (SURVEY.md
8c), so these vectors pin OUR restatement (regression anchor for the oracle and a GPU-free record
of what the CUDA path must reproduce); they are not outputs of the reference binary.

    python tests/golden/make_golden.py      # rewrites the fixture from the current oracle
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def digest(a):
    a = np.ascontiguousarray(a)
    return {"shape": list(a.shape), "dtype": str(a.dtype), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
            "head": a.reshape(-1)[:8].tolist(), "tail": a.reshape(-1)[-8:].tolist()}


def cases(O):
    from helpers import make_input
    out = {}
    x = O.synth_f32(5000, 2, 44100, 7)
    out["synth_f32_t7"] = digest(x)
    out["synth_s16_t7"] = digest(O.f32_to_s16(x))
    for fmt in (1, 2, 3, 6, 7, 8):
        xi = make_input(O, fmt, 3001, 2, track=fmt)
        out[f"gain_fmt{fmt}_v0.8"] = digest(O.gain(xi, fmt, 0.8))
        out[f"extract_fmt{fmt}"] = digest(O.extract_interleaved(xi, fmt))
    l, r = O.swr_whole(x, O.FMT_FLT, 44100, 48000)
    out["swr_44100_48000_flt"] = digest(np.stack([l, r]))
    xs = make_input(O, O.FMT_S16, 4000, 1, rate=22050, track=2)
    l, r = O.swr_whole(xs, O.FMT_S16, 22050, 48000)
    out["swr_22050_48000_s16_mono"] = digest(np.stack([l, r]))
    tr = [O.make_track(O.synth_f32(9000 + 500 * i, 2, 44100, i), O.FMT_FLT, 44100) for i in range(3)]
    l, r = O.amix(tr, [0.5, 0.25, 1.0])
    out["amix3_44100"] = digest(np.stack([l, r]))
    l, r = O.bimix(tr[0], tr[1], 0.3)
    out["bimix_bias0.3"] = digest(np.stack([l, r]))
    o, pts = O.bimix_v2(tr[0], tr[2])
    out["bimix_v2"] = digest(o)
    x48 = O.synth_f32(48000, 2, 48000, 5)
    y, offs, _ = O.soundtouch(x48, 48000, 1.0, O.pitch_node_factor(3.0))
    out["soundtouch_pitch3"] = digest(y); out["soundtouch_pitch3_offsets"] = digest(offs)
    y, offs, _ = O.soundtouch(x48, 48000, 1.25, O.velocity_node_pitch(1.25, True))
    out["soundtouch_tempo1.25"] = digest(y); out["soundtouch_tempo1.25_offsets"] = digest(offs)
    y, offs, _ = O.soundtouch(x48[:, :1].copy(), 48000, 1.0, O.pitch_node_factor(-4.0))
    out["soundtouch_mono_pitch-4"] = digest(y)
    out["stft_t5_ch0"] = digest(O.stft(x48[:, 0].copy()).view(np.float32))
    return out


if __name__ == "__main__":
    from oracle import oracle as O
    json.dump(cases(O), open(os.path.join(HERE, "oracle_r1.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "oracle_r1.json"))
