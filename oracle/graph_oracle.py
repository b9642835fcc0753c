"""Config-5 render graph composed from the CPU oracle's node functions (TEST INFRASTRUCTURE ONLY:
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs).

Same graph as nodey-audio-editor_b200/bindings/pipeline.py, node by node, each node a restatement
of the reference processor (see nodey_oracle.h).  Frames between nodes follow the reference: amix
emits nb samples per iteration (run-length encoded sizes are handed to the next amix), the
SoundTouch nodes are driven with 1152-sample putSamples chunks.
"""
import numpy as np

from . import oracle as O

IN_RATE = 44100
GROUP = 16


def track_gain(t):
    return float(np.float32(0.5) + np.float32(0.0625) * np.float32(t % 9))


def uniform_runs(n, fs=1152):
    q, r = divmod(int(n), int(fs))
    return ([(fs, q)] if q else []) + ([(r, 1)] if r else [])


def amix_runs(in_runs_list):
    """frame sizes an amix node emits for same-rate (48 kHz) inputs: nb = min frame size per iteration,
    1152 once all inputs ran dry, until every input's buffered surplus is drained (audio-amix.cpp:190-195, 290, 320)."""
    its = [[l for l, c in runs for _ in range(c)] for runs in in_runs_list]
    return its


def track_chain(x441, gain, frame_size=1152, pitch_semitones=3.0, velocity=1.25, keep=None):
    """audio_amix(1) -> pitch_modifier -> velocity_modifier -> audio_volume_adjust for one track.
    x441: [n, 2] float32.  Returns interleaved float32 [m2, 2]."""
    l, r = O.amix([O.make_track(x441, O.FMT_FLT, IN_RATE, frame_size)], [1.0])
    xi = O.extract_interleaved(np.stack([l, r]), O.FMT_FLTP)
    y1, _, _ = O.soundtouch(xi, 48000, 1.0, O.pitch_node_factor(pitch_semitones), frame_size, want_offsets=False)
    y2, _, _ = O.soundtouch(y1, 48000, velocity, O.velocity_node_pitch(velocity, True), frame_size, want_offsets=False)
    if keep is not None:
        keep["amix1"] = np.stack([l, r]); keep["pitch"] = y1; keep["tempo"] = y2
    return O.gain(y2, O.FMT_FLT, gain)


def _amix_with_runs(tracks_planar_or_packed, fmts, runs_list, vols):
    """amix over 48 kHz inputs; returns planes and the node's own output frame runs."""
    tr = [O.make_track(x, f, 48000, 1152, runs=runs) for x, f, runs in zip(tracks_planar_or_packed, fmts, runs_list)]
    l, r = O.amix(tr, vols)
    # output frame sizes: nb per iteration = min over inputs that still have a frame, else 1152
    seqs = [[ln for ln, c in runs for _ in range(c)] for runs in runs_list]
    out, total, m = [], 0, 0
    while total < len(l):
        present = [s[m] for s in seqs if m < len(s)]
        nb = min(present) if present else 1152
        out.append(nb); total += nb; m += 1
    assert total == len(l)
    runs = []
    for nb in out:
        if runs and runs[-1][0] == nb:
            runs[-1] = (nb, runs[-1][1] + 1)
        else:
            runs.append((nb, 1))
    return np.stack([l, r]), runs


def render(tracks, first_track=0, frame_size=1152, group_vol=1.0 / 16, master_vol=1.0 / 16, threads=1, spectrum=True):
    """tracks: list of [n, 2] float32 at 44.1 kHz, a multiple of 16.  Returns (bus [2, total], spectrum or None)."""
    assert len(tracks) % GROUP == 0
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:   # ctypes releases the GIL: one oracle chain per host thread
            ys = list(ex.map(lambda a: track_chain(a[1], track_gain(first_track + a[0]), frame_size), enumerate(tracks)))
    else:
        ys = [track_chain(x, track_gain(first_track + i), frame_size) for i, x in enumerate(tracks)]
    buses, bus_runs = [], []
    for g in range(len(tracks) // GROUP):
        grp = ys[g * GROUP:(g + 1) * GROUP]
        b, runs = _amix_with_runs(grp, [O.FMT_FLT] * GROUP, [uniform_runs(len(y), frame_size) for y in grp], [group_vol] * GROUP)
        buses.append(b); bus_runs.append(runs)
    bus, _ = _amix_with_runs(buses, [O.FMT_FLTP] * len(buses), bus_runs, [master_vol] * len(buses))
    spec = None
    if spectrum:
        spec = np.stack([O.stft(bus[c].copy()) for c in range(2)])
    return bus, spec
