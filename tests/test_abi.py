"""CPU: the C-ABI library loads and exports every symbol include/nodey_cuda.h declares; host-only
entry points (no device work) behave.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "nodey_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nodey_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import nodey
    lib = nodey.lib()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libnodey_cuda.so does not export {n}"
    assert sorted(nodey.SYMBOLS) == names, "binding symbol list and header disagree"


def test_version_and_error_text():
    import nodey
    lib = nodey.lib()
    assert lib.nodey_version() >= 100
    assert isinstance(lib.nodey_last_error(), bytes)


def test_stft_frame_count_host_only():
    import nodey
    assert nodey.stft_frames(4095) == 0
    assert nodey.stft_frames(4096) == 1
    assert nodey.stft_frames(48000 * 3600) == 168747


def test_amix_plan_matches_oracle_node(orc):
    """audio_amix bookkeeping (host only) against the frame-by-frame oracle node: total length, placement
    of every input's resampled frames (incl. the gaps of inputs above 48 kHz), output frame sizes."""
    import nodey
    cases = [([44100], [30011], [1152]), ([44100, 48000, 22050], [30000, 20000, 9000], [1152, 1024, 4096]),
             ([96000], [50000], [1152]), ([44100, 44100], [100, 40000], [1152, 1152]), ([48000], [5000], [1152]),
             ([8000, 44100], [4000, 10], [1152, 1152])]
    for rates, lens, fs in cases:
        total, segs, runs = nodey.amix_plan(rates, [nodey.uniform_runs(l, f) for l, f in zip(lens, fs)])
        xs = [orc.synth_f32(l, 2, r, i) for i, (r, l) in enumerate(zip(rates, lens))]
        tr = [orc.make_track(x, orc.FMT_FLT, r, f) for x, r, f in zip(xs, rates, fs)]
        ol, orr = orc.amix(tr, [1.0] * len(rates))
        assert total == len(ol), (rates, lens)
        assert sum(l * c for l, c in runs) == total
        exp = np.zeros(total, np.float32)
        for i in range(len(rates)):
            wl, _ = orc.swr_whole(xs[i], orc.FMT_FLT, rates[i], 48000, flush=True)
            pl = np.zeros(total, np.float32)
            for (ii, o, s_, l) in segs:
                if ii == i:
                    pl[o:o + l] = wl[s_:s_ + l]
            exp = (exp + pl * np.float32(1.0)).astype(np.float32)
        assert np.array_equal(exp, ol), (rates, lens)


def test_amix_plan_rejects_17_inputs():
    import nodey
    import pytest
    with pytest.raises(nodey.NodeyError):
        nodey.amix_plan([44100] * 17, [nodey.uniform_runs(100)] * 17)


def test_host_library_exports_every_symbol_of_the_engine_header():
    """libnodey_host.so (the C facade of the C++ host layer) exports what include/nodey_engine.h declares"""
    import engine
    text = open(os.path.join(ROOT, "include", "nodey_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(nodey_engine_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 20
    lib = engine.lib()
    for n in names:
        assert hasattr(lib, n), f"libnodey_host.so does not export {n}"
