"""SoundTouch pin (CPU).  The reference's pitch / tempo arithmetic is SoundTouch 2.3.2 (xmake.lua:16), absent from this
image: the oracle restates it and the CUDA kernels match the oracle bit for bit, so pinning the oracle pins the path.

  * NODEY_REAL_SOUNDTOUCH=/path/to/lib set  -> the oracle is compared LIVE with the real library, driven like
    audio-velocity.cpp:367-435 (1152-frame puts, min(numSamples, 3 * 1152 / velocity) receives, flush): output length
    exact (it moves by whole sequences when one WSOLA offset differs), samples within 1e-5 of the signal peak;
  * tests/golden/st_real.npz present (made by tests/golden/make_st_golden.py from a real library) -> the same
    comparison against the recorded outputs, on any machine;
  * neither -> those tests SKIP with "parity unpinned".  What always runs: the harness itself against the stand-in
    library tests/fake_soundtouch (the oracle's model behind SoundTouchDLL's C entry points -- proves the binding,
    the two driving loops and the fixture round trip work, proves nothing about SoundTouch), and the reference loop's
    early break on the oracle model.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_st_golden as G  # noqa: E402

GOLD_PATH = os.path.join(HERE, "golden", "st_real.npz")
TOL = 1e-5          # BASELINE.json: float DSP within 1e-5 relative


def _oracle_runs(orc, tag):
    from oracle import real_soundtouch as R
    sr, ch, secs, velocity, kind = G.CASES[tag]
    x = G.case_input(orc, tag)
    p = G.pitch_arg(orc, kind)
    canonical, _, _ = orc.soundtouch(x, sr, velocity, p, 1152)
    loop, sizes, flushed = orc.soundtouch_reference_loop(x, sr, velocity, p, 1152)
    return x, canonical, loop, sizes, flushed


def _compare(orc, tag, ref_canonical, ref_loop, ref_sizes, ref_flushed, exact):
    x, canonical, loop, sizes, flushed = _oracle_runs(orc, tag)
    peak = max(float(np.abs(x).max()), 1e-9)
    assert canonical.shape == ref_canonical.shape, f"{tag}: canonical output length {canonical.shape} vs library {ref_canonical.shape}"
    assert loop.shape == ref_loop.shape, f"{tag}: reference-loop output length"
    assert list(sizes) == list(ref_sizes), f"{tag}: receive sizes of the reference loop"
    assert bool(flushed) == bool(ref_flushed), f"{tag}: whether the loop reached flush()"
    if exact:
        assert np.array_equal(canonical.view(np.uint32), ref_canonical.view(np.uint32)) and np.array_equal(loop.view(np.uint32), ref_loop.view(np.uint32))
    else:
        assert np.abs(canonical - ref_canonical).max() <= TOL * peak, f"{tag}: canonical samples"
        assert np.abs(loop - ref_loop).max() <= TOL * peak, f"{tag}: reference-loop samples"


# ---- the real thing --------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", sorted(G.CASES))
def test_oracle_against_the_real_soundtouch_library(orc, tag):
    from oracle import real_soundtouch as R
    if not R.library_path():
        pytest.skip(f"parity unpinned: {R.ENV} is not set (no SoundTouch library in this image)")
    name, vid = R.version()
    if vid == 0:
        pytest.skip("parity unpinned: NODEY_REAL_SOUNDTOUCH points at the stand-in library")
    sr, ch, secs, velocity, kind = G.CASES[tag]
    x = G.case_input(orc, tag)
    p = G.pitch_arg(orc, kind)
    make = lambda: R.SoundTouch(sr, ch, velocity, p)
    ref_canonical = R.run_canonical(make, x)
    ref_loop, ref_sizes, ref_flushed = R.run_reference_loop(make, x, velocity)
    _compare(orc, tag, ref_canonical, ref_loop, ref_sizes, ref_flushed, exact=False)


@pytest.mark.parametrize("tag", sorted(G.CASES))
def test_oracle_against_the_recorded_soundtouch_outputs(orc, tag):
    if not os.path.exists(GOLD_PATH):
        pytest.skip("parity unpinned: tests/golden/st_real.npz has not been generated (needs a real SoundTouch, see make_st_golden.py)")
    gold = np.load(GOLD_PATH)
    _compare(orc, tag, gold[f"{tag}_canonical"], gold[f"{tag}_loop"], gold[f"{tag}_loop_sizes"], int(gold[f"{tag}_loop_flushed"]), exact=False)


# ---- the harness, against the stand-in ---------------------------------------------------------------------
@pytest.fixture(scope="module")
def standin(tmp_path_factory):
    """tests/fake_soundtouch built against the oracle's C file; bound in a SUBPROCESS-free way by pointing the module at it"""
    out = tmp_path_factory.mktemp("fake_st") / "libfake_soundtouch.so"
    src = os.path.join(HERE, "fake_soundtouch", "fake_soundtouch.c")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", str(out), src,
                    os.path.join(ROOT, "oracle", "nodey_oracle.c"), "-lm"] + (["-mfma"] if " fma " in open("/proc/cpuinfo").read() else []),
                   check=True)
    from oracle import real_soundtouch as R
    old_env, old_lib = os.environ.get(R.ENV), R._lib
    os.environ[R.ENV] = str(out)
    R._lib = None
    yield R
    R._lib = old_lib
    if old_env is None:
        os.environ.pop(R.ENV, None)
    else:
        os.environ[R.ENV] = old_env


@pytest.mark.parametrize("tag", ["pitch_p3_48k", "tempo_1p25_keep_48k", "mono_tempo_1p5_22k", "velocity_0p7_44k", "short_0p2s"])
def test_harness_runs_end_to_end_on_the_standin(orc, standin, tag):
    """NOT a parity result (the stand-in is the oracle's own model): the binding, both driving loops and the comparison
    code work, and the C loop restatement (orc_soundtouch_reference_loop) equals the Python one run on the C API."""
    R = standin
    assert R.version()[1] == 0
    sr, ch, secs, velocity, kind = G.CASES[tag]
    x = G.case_input(orc, tag)
    p = G.pitch_arg(orc, kind)
    make = lambda: R.SoundTouch(sr, ch, velocity, p)
    ref_canonical = R.run_canonical(make, x)
    ref_loop, ref_sizes, ref_flushed = R.run_reference_loop(make, x, velocity)
    _compare(orc, tag, ref_canonical, ref_loop, ref_sizes, ref_flushed, exact=True)


def test_fixture_generator_round_trip_on_the_standin(orc, standin, tmp_path):
    data = G.generate(standin)
    assert set(data) == {f"{t}_{k}" for t in G.CASES for k in ("canonical", "loop", "loop_sizes", "loop_flushed")} | {"meta"}
    path = tmp_path / "st.npz"
    np.savez_compressed(path, **data)
    back = np.load(path)
    tag = "pitch_p3_48k"
    _compare(orc, tag, back[f"{tag}_canonical"], back[f"{tag}_loop"], back[f"{tag}_loop_sizes"], int(back[f"{tag}_loop_flushed"]), exact=True)
    # the generator script refuses to record the stand-in as ground truth
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_st_golden.py")], capture_output=True, text=True,
                       env=dict(os.environ))
    assert r.returncode != 0 and "stand-in" in (r.stderr + r.stdout)


# ---- App. C7 on the oracle model ---------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["pitch_p3_48k", "tempo_1p25_keep_48k", "velocity_0p7_44k", "mono_pitch_p3_48k"])
def test_reference_loop_drops_the_tail_it_never_flushes(orc, tag):
    """audio-velocity.cpp:414: `if (numSamples() == 0 && input_stream_eof) break;` comes BEFORE the flush branch, and the
    receive above it empties the FIFO whenever more than 1152 / velocity samples are queued -- so with a frame available
    at every turn the loop ends without flush() and what SoundTouch still holds is lost.  The loop's output is a strict
    prefix of the canonical (always flushed) render; the engine's `reference_schedule` mode reproduces length and frame sizes."""
    x, canonical, loop, sizes, flushed = _oracle_runs(orc, tag)
    assert sum(sizes) == len(loop) <= len(canonical)
    assert np.array_equal(loop.view(np.uint32), canonical[:len(loop)].view(np.uint32)), "the loop's samples are a prefix of the canonical render"
    if not flushed:
        assert len(loop) < len(canonical)
        assert len(canonical) - len(loop) < 48000 // 4      # the lost tail is SoundTouch's latency, a fraction of a second
