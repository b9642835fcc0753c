"""ctypes binding of the CPU oracle (oracle/nodey_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.  Pinning: the reference has no
fixtures; the libswresample model is pinned against a real libswresample (oracle/real_swr.py,
tests/golden/swr_real.npz), the SoundTouch model is PARITY UNPINNED (see nodey_oracle.c / .h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnodey_oracle.so")

FMT_U8, FMT_S16, FMT_S32, FMT_FLT, FMT_DBL, FMT_U8P, FMT_S16P, FMT_S32P, FMT_FLTP, FMT_DBLP = range(10)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("nodey_oracle.c", "nodey_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


class Track(C.Structure):
    _fields_ = [("plane0", C.c_void_p), ("plane1", C.c_void_p), ("fmt", C.c_int), ("rate", C.c_int),
                ("ch", C.c_int), ("nframes", C.c_int64), ("frame_size", C.c_int), ("pts0", C.c_double),
                ("run_len", C.POINTER(C.c_int64)), ("run_count", C.POINTER(C.c_int64)), ("nruns", C.c_int)]


class StInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_int), ("channels", C.c_int), ("rate", C.c_double), ("tempo", C.c_double),
                ("overlap", C.c_int), ("seek_window", C.c_int), ("seek_length", C.c_int), ("sample_req", C.c_int),
                ("nominal_skip", C.c_double), ("tdstretch_first", C.c_int), ("n_sequences", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, i64, f32p = C.c_void_p, C.c_int64, C.POINTER(C.c_float)
    L.orc_version.restype = C.c_char_p
    L.orc_synth_f32.argtypes = [vp, i64, C.c_int, C.c_int, C.c_int, i64]
    L.orc_f32_to_s16.argtypes = [vp, vp, i64]
    L.orc_synth_hash.argtypes = [C.c_uint32, C.c_uint64]
    L.orc_synth_hash.restype = C.c_uint32
    L.orc_gain.argtypes = [vp, vp, C.c_int, i64, C.c_float]
    L.orc_extract_interleaved.argtypes = [vp, vp, vp, C.c_int, i64, C.c_int]
    L.orc_extract_interleaved.restype = C.c_int
    L.orc_split.argtypes = [vp, vp, vp, vp, C.c_int, i64]
    L.orc_swr_create.argtypes = [C.c_int] * 5
    L.orc_swr_create.restype = vp
    L.orc_swr_free.argtypes = [vp]
    L.orc_swr_convert.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int]
    L.orc_swr_convert.restype = C.c_int
    L.orc_swr_plan.argtypes = [vp, C.POINTER(C.c_int)]
    L.orc_swr_filter_bank.argtypes = [vp]
    L.orc_swr_filter_bank.restype = f32p
    L.orc_swr_whole.argtypes = [C.c_int] * 5 + [vp, vp, i64, C.c_int, vp, vp, i64]
    L.orc_swr_whole.restype = i64
    L.orc_swr_out_count.argtypes = [C.c_int, C.c_int, C.c_int, i64, C.c_int]
    L.orc_swr_out_count.restype = i64
    L.orc_amix.argtypes = [C.POINTER(Track), C.c_int, vp, C.c_int, vp, vp, i64]
    L.orc_amix.restype = i64
    L.orc_bimix.argtypes = [C.POINTER(Track), C.POINTER(Track), C.c_float, C.c_int, vp, vp, i64]
    L.orc_bimix.restype = i64
    L.orc_bimix_v2.argtypes = [C.POINTER(Track), C.POINTER(Track), C.c_int, vp, i64, C.POINTER(C.c_double)]
    L.orc_bimix_v2.restype = i64
    L.orc_soundtouch.argtypes = [vp, i64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, vp, i64, vp, i64,
                                 C.POINTER(StInfo)]
    L.orc_soundtouch.restype = i64
    L.orc_st_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float]
    L.orc_st_create.restype = vp
    L.orc_st_destroy.argtypes = [vp]
    L.orc_st_destroy.restype = None
    L.orc_st_put.argtypes = [vp, vp, i64]
    L.orc_st_put.restype = None
    L.orc_st_num_samples.argtypes = [vp]
    L.orc_st_num_samples.restype = i64
    L.orc_st_receive.argtypes = [vp, vp, i64]
    L.orc_st_receive.restype = i64
    L.orc_st_flush.argtypes = [vp]
    L.orc_st_flush.restype = None
    L.orc_soundtouch_reference_loop.argtypes = [vp, i64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, vp, i64, vp, i64,
                                                C.POINTER(i64), C.POINTER(C.c_int)]
    L.orc_soundtouch_reference_loop.restype = i64
    L.orc_pitch_node_factor.argtypes = [C.c_float]
    L.orc_pitch_node_factor.restype = C.c_float
    L.orc_velocity_node_pitch.argtypes = [C.c_float, C.c_int]
    L.orc_velocity_node_pitch.restype = C.c_float
    L.orc_stft_frames.argtypes = [i64, C.c_int, C.c_int]
    L.orc_stft_frames.restype = i64
    L.orc_stft.argtypes = [vp, i64, C.c_int, C.c_int, vp]
    L.orc_stft.restype = i64
    L.orc_hann_window.argtypes = [vp, C.c_int]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


_NP = {FMT_S16: np.int16, FMT_S16P: np.int16, FMT_S32: np.int32, FMT_S32P: np.int32,
       FMT_FLT: np.float32, FMT_FLTP: np.float32}


def np_dtype(fmt):
    return _NP[fmt]


def is_planar(fmt):
    return fmt >= FMT_U8P


# ---------------------------------------------------------------------------------------------
def synth_f32(nframes, nch, sample_rate, track=0, frame0=0):
    out = np.empty((nframes, nch), np.float32)
    lib().orc_synth_f32(_p(out), nframes, nch, sample_rate, track, frame0)
    return out


def f32_to_s16(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.shape, np.int16)
    lib().orc_f32_to_s16(_p(out), _p(x), x.size)
    return out


def gain(src, fmt, volume):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_gain(_p(dst), _p(src), fmt, src.size, np.float32(volume))
    return dst


def planes_of(x, fmt):
    """x: packed -> array [frames, ch]; planar -> array [ch, frames]. Returns (p0, p1, nframes, nch)."""
    x = np.ascontiguousarray(x)
    if is_planar(fmt):
        nch, n = x.shape
        return x[0], (x[1] if nch > 1 else None), n, nch
    n, nch = x.shape
    return x, None, n, nch


def extract_interleaved(x, fmt):
    p0, p1, n, nch = planes_of(x, fmt)
    out = np.empty((n, nch), np.float32)
    rc = lib().orc_extract_interleaved(_p(out), _p(p0), _p(p1), fmt, n, nch)
    if rc != 0:
        raise ValueError("Unsupported sample format")
    return out


def split(x, fmt):
    p0, p1, n, nch = planes_of(x, fmt)
    assert nch == 2
    l = np.empty(n, np_dtype(fmt)); r = np.empty(n, np_dtype(fmt))
    lib().orc_split(_p(l), _p(r), _p(p0), _p(p1), fmt, n)
    return l, r


class Swr:
    def __init__(self, in_rate, out_rate, in_fmt, in_ch, quirk=0):
        self.h = lib().orc_swr_create(in_rate, out_rate, in_fmt, in_ch, quirk)
        self.fmt, self.ch = in_fmt, in_ch

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_swr_free(self.h)
            self.h = None

    def plan(self):
        a = (C.c_int * 8)()
        lib().orc_swr_plan(self.h, a)
        keys = ["phase_count", "filter_length", "filter_alloc", "dst_incr_div", "dst_incr_mod", "src_incr",
                "index0", "linear"]
        return dict(zip(keys, list(a)))

    def filter_bank(self):
        p = self.plan()
        n = (p["phase_count"] + 1) * p["filter_alloc"]
        ptr = lib().orc_swr_filter_bank(self.h)
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(p["phase_count"] + 1, p["filter_alloc"]).copy()

    def convert(self, x, out_count):
        """x None => flush call. Returns (left, right) float32 arrays of the produced length."""
        ol = np.zeros(out_count, np.float32); orr = np.zeros(out_count, np.float32)
        if x is None:
            n = lib().orc_swr_convert(self.h, _p(ol), _p(orr), out_count, None, None, 0)
        else:
            p0, p1, nfr, nch = planes_of(x, self.fmt)
            n = lib().orc_swr_convert(self.h, _p(ol), _p(orr), out_count, _p(p0), _p(p1), nfr)
        return ol[:n], orr[:n]


def swr_out_count(in_rate, out_rate, in_frames, flush=True, quirk=0):
    return lib().orc_swr_out_count(in_rate, out_rate, quirk, in_frames, 1 if flush else 0)


def swr_whole(x, fmt, in_rate, out_rate=48000, flush=True, quirk=0):
    p0, p1, n, nch = planes_of(x, fmt)
    cap = swr_out_count(in_rate, out_rate, n, True, quirk) + 64
    ol = np.zeros(cap, np.float32); orr = np.zeros(cap, np.float32)
    got = lib().orc_swr_whole(in_rate, out_rate, fmt, nch, quirk, _p(p0), _p(p1), n, 1 if flush else 0,
                              _p(ol), _p(orr), cap)
    return ol[:got], orr[:got]


def make_track(x, fmt, rate, frame_size=1152, pts0=0.0, runs=None):
    """runs: optional [(frame_len, count), ...] frame sizes (e.g. the output frames of an amix node)."""
    p0, p1, n, nch = planes_of(x, fmt)
    if runs:
        rl = (C.c_int64 * len(runs))(*[int(r[0]) for r in runs])
        rc = (C.c_int64 * len(runs))(*[int(r[1]) for r in runs])
        assert sum(int(a) * int(b) for a, b in runs) == n, "frame runs do not cover the track"
        t = Track(_p(p0), _p(p1), fmt, rate, nch, n, frame_size, pts0, rl, rc, len(runs))
        t._keep = (x, p0, p1, rl, rc)
        return t
    t = Track(_p(p0), _p(p1), fmt, rate, nch, n, frame_size, pts0, None, None, 0)
    t._keep = (x, p0, p1)
    return t


def amix(tracks, volumes, quirk=0):
    arr = (Track * len(tracks))(*tracks)
    vol = np.ascontiguousarray(volumes, np.float32)
    cap = max(int(t.nframes * 48000 // t.rate) for t in tracks) + 8 * 4096
    ol = np.zeros(cap, np.float32); orr = np.zeros(cap, np.float32)
    n = lib().orc_amix(arr, len(tracks), _p(vol), quirk, _p(ol), _p(orr), cap)
    return ol[:n], orr[:n]


def bimix(tl, tr, bias, quirk=0):
    # the loop emits 1152 samples per iteration while both inputs still have a frame, whatever the frame size
    # (audio-bimix.cpp:174-181), so tiny frames make the stream far longer than the resampled inputs
    def iterations(t):
        if t.nruns:
            return sum(int(t.run_count[k]) for k in range(t.nruns))
        return -(-int(t.nframes) // max(int(t.frame_size), 1))
    cap = max(int(t.nframes * 48000 // t.rate) for t in (tl, tr)) + 1152 * (max(iterations(tl), iterations(tr)) + 64) + 8 * 4096
    ol = np.zeros(cap, np.float32); orr = np.zeros(cap, np.float32)
    n = lib().orc_bimix(C.byref(tl), C.byref(tr), np.float32(bias), quirk, _p(ol), _p(orr), cap)
    return ol[:n], orr[:n]


def bimix_v2(tl, tr, quirk=0):
    cap = int(max(t.pts0 * 48000 + t.nframes * 48000 // t.rate for t in (tl, tr))) + 8 * 4096
    out = np.zeros((cap, 2), np.float32)
    pts = C.c_double(0)
    n = lib().orc_bimix_v2(C.byref(tl), C.byref(tr), quirk, _p(out), cap, C.byref(pts))
    return out[:n], pts.value


def soundtouch(x, sample_rate, rate_arg, pitch_arg, frame_size=1152, want_offsets=True):
    x = np.ascontiguousarray(x, np.float32)
    n, nch = x.shape
    eff = float(np.float32(rate_arg)) * 1.0
    cap = int(n / max(eff, 1e-3)) + 65536
    out = np.zeros((cap, nch), np.float32)
    ocap = n // 256 + 64
    offs = np.zeros(ocap, np.int32)
    info = StInfo()
    got = lib().orc_soundtouch(_p(x), n, nch, sample_rate, np.float32(rate_arg), np.float32(pitch_arg), frame_size,
                               _p(out), cap, _p(offs) if want_offsets else None, ocap, C.byref(info))
    nseq = int(info.n_sequences)
    return out[:got], offs[:max(nseq - 1, 0)], info


def soundtouch_reference_loop(x, sample_rate, velocity, pitch_arg, frame_size=1152):
    """soundtouch_process_payload (audio-velocity.cpp:286-441) literally, one input frame per loop turn: returns
    (samples, chunk sizes the node pushes downstream, flushed) -- flushed is False when the loop took the early break
    before flush() and the tail SoundTouch still held was lost."""
    x = np.ascontiguousarray(x, np.float32)
    n, nch = x.shape
    cap = int(n / max(float(velocity), 1e-3)) + 65536
    out = np.zeros((cap, nch), np.float32)
    ccap = n // max(frame_size, 1) + 64
    chunks = np.zeros(ccap, np.int64)
    nchunks = C.c_int64()
    flushed = C.c_int()
    got = lib().orc_soundtouch_reference_loop(_p(x), n, nch, sample_rate, np.float32(velocity), np.float32(pitch_arg), frame_size,
                                              _p(out), cap, _p(chunks), ccap, C.byref(nchunks), C.byref(flushed))
    return out[:got], chunks[:min(nchunks.value, ccap)].tolist(), bool(flushed.value)


class St:
    """the streaming SoundTouch model (SoundTouch's public calls): putSamples / numSamples / receiveSamples / flush"""

    def __init__(self, sample_rate, nch, rate_arg, pitch_arg):
        self.nch = nch
        self.h = lib().orc_st_create(sample_rate, nch, np.float32(rate_arg), np.float32(pitch_arg))

    def put(self, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.nch)
        lib().orc_st_put(self.h, _p(x), x.shape[0])

    def num_samples(self):
        return int(lib().orc_st_num_samples(self.h))

    def receive(self, max_frames):
        out = np.zeros((max(int(max_frames), 0), self.nch), np.float32)
        n = lib().orc_st_receive(self.h, _p(out), out.shape[0]) if out.shape[0] else 0
        return out[:n]

    def flush(self):
        lib().orc_st_flush(self.h)

    def close(self):
        if self.h:
            lib().orc_st_destroy(self.h)
            self.h = None

    __del__ = close


def pitch_node_factor(semitones):
    return float(lib().orc_pitch_node_factor(np.float32(semitones)))


def velocity_node_pitch(velocity, keep_pitch):
    return float(lib().orc_velocity_node_pitch(np.float32(velocity), 1 if keep_pitch else 0))


def stft_frames(n, nfft=4096, hop=1024):
    return lib().orc_stft_frames(n, nfft, hop)


def stft(x, nfft=4096, hop=1024):
    x = np.ascontiguousarray(x, np.float32)
    m = stft_frames(x.size, nfft, hop)
    out = np.zeros((m, nfft // 2 + 1, 2), np.float32)
    lib().orc_stft(_p(x), x.size, nfft, hop, _p(out))
    return out.view(np.complex64)[..., 0]


def hann(nfft=4096):
    w = np.empty(nfft, np.float32)
    lib().orc_hann_window(_p(w), nfft)
    return w
