"""Records what the REAL SoundTouch library produces for the reference's call sequence (audio-velocity.cpp:367-435) into
tests/golden/st_real.npz, so that the oracle's SoundTouch model -- and through it the CUDA kernels, which match the
oracle bit for bit -- can be pinned on machines without the library.

    NODEY_REAL_SOUNDTOUCH=/path/to/libSoundTouchDll.so python tests/golden/make_st_golden.py

NOT RUN YET: no SoundTouch exists in the build image (no st_real.npz is committed; tests/test_st_real.py says "parity
unpinned" and skips the fixture comparison).  The cases below are the parameter sets of tests/test_gpu_soundtouch.py plus
BASELINE.json configs[1] (pitch +3 semitones, tempo 1.25 keep-pitch, 48 kHz stereo).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# tag -> (sample_rate, channels, seconds, velocity (= setRate), kind of setPitch argument)
CASES = {
    "pitch_p3_48k": (48000, 2, 4.0, 1.0, ("semitones", 3.0)),
    "pitch_m3_48k": (48000, 2, 4.0, 1.0, ("semitones", -3.0)),
    "tempo_1p25_keep_48k": (48000, 2, 4.0, 1.25, ("keep", 1.25)),
    "tempo_0p8_keep_48k": (48000, 2, 3.0, 0.8, ("keep", 0.8)),
    "velocity_1p25_48k": (48000, 2, 3.0, 1.25, ("none", 0.0)),
    "velocity_0p7_44k": (44100, 2, 3.0, 0.7, ("none", 0.0)),
    "pitch_p7_44k": (44100, 2, 3.0, 1.0, ("semitones", 7.0)),
    "mono_pitch_p3_48k": (48000, 1, 3.0, 1.0, ("semitones", 3.0)),
    "mono_tempo_1p5_22k": (22050, 1, 3.0, 1.5, ("keep", 1.5)),
    "stereo_11k_m5": (11025, 2, 3.0, 1.0, ("semitones", -5.0)),
    "identity_48k": (48000, 2, 2.0, 1.0, ("none", 0.0)),
    "short_0p2s": (48000, 2, 0.2, 1.0, ("semitones", 3.0)),
}


def pitch_arg(orc, kind):
    k, v = kind
    if k == "semitones":
        return orc.pitch_node_factor(v)         # std::pow(2.0f, pitch / 12.0f), audio-velocity.cpp:473-474
    if k == "keep":
        return orc.velocity_node_pitch(v, True)  # 1 / velocity, audio-velocity.cpp:457
    return 1.0


def case_input(orc, tag):
    sr, ch, secs, _, _ = CASES[tag]
    return orc.synth_f32(int(sr * secs) + 17, ch, sr, 5 + sorted(CASES).index(tag))


def generate(R=None):
    """{tag_canonical, tag_loop, tag_loop_sizes, tag_loop_flushed} from the library behind oracle/real_soundtouch.py"""
    from oracle import oracle as O
    if R is None:
        from oracle import real_soundtouch as R
    O.build()
    out = {}
    for tag, (sr, ch, secs, velocity, kind) in sorted(CASES.items()):
        x = case_input(O, tag)
        p = pitch_arg(O, kind)
        make = lambda: R.SoundTouch(sr, ch, velocity, p)
        out[f"{tag}_canonical"] = R.run_canonical(make, x)
        y, sizes, flushed = R.run_reference_loop(make, x, velocity)
        out[f"{tag}_loop"] = y
        out[f"{tag}_loop_sizes"] = np.asarray(sizes, np.int64)
        out[f"{tag}_loop_flushed"] = np.asarray(int(flushed))
    name, vid = R.version()
    out["meta"] = np.asarray(f"{name} ({vid})")
    return out


if __name__ == "__main__":
    from oracle import real_soundtouch as R
    if R.lib() is None:
        sys.exit(f"set {R.ENV} to a SoundTouch library (SoundTouchDLL C API, or oracle/st_shim.cpp over libSoundTouch)")
    name, vid = R.version()
    if vid == 0:
        sys.exit("this is the stand-in library of tests/fake_soundtouch (the oracle on both sides): not a ground truth, nothing written")
    data = generate(R)
    path = os.path.join(HERE, "st_real.npz")
    np.savez_compressed(path, **data)
    print(f"wrote {path} from {name} ({vid}): {len(CASES)} cases")
