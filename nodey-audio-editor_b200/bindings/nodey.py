"""ctypes binding of libnodey_cuda.so -- the C ABI declared in include/nodey_cuda.h.

This is the same boundary the C++ host layer (host/) calls; Python is used by tests/ and bench.py
only.  torch supplies device memory and streams (plumbing); every computation is a kernel of
libnodey_cuda.so.  There is NO fallback: a missing library or a failing call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NODEY_CUDA_LIB: development override (instrumented builds under tools/micro); the default is the in-tree library
LIB_PATH = os.environ.get("NODEY_CUDA_LIB") or os.path.join(os.path.dirname(_HERE), "libnodey_cuda.so")

FMT_U8, FMT_S16, FMT_S32, FMT_FLT, FMT_DBL, FMT_U8P, FMT_S16P, FMT_S32P, FMT_FLTP, FMT_DBLP = range(10)
MAX_MIX_INPUTS = 16

# every entry point include/nodey_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "nodey_version", "nodey_last_error", "nodey_device_info", "nodey_synth", "nodey_gain",
    "nodey_extract_interleaved", "nodey_split", "nodey_to_fltp_stereo", "nodey_mix", "nodey_mix_gains", "nodey_bimix",
    "nodey_downmix_half", "nodey_merge_segments", "nodey_resampler_create", "nodey_resampler_destroy",
    "nodey_resampler_info", "nodey_resampler_filter_bank", "nodey_resampler_out_count", "nodey_resampler_run",
    "nodey_resampler_run_mode", "nodey_resample_mix", "nodey_resample_tracks", "nodey_resample_tracks_chunks", "nodey_resample_tracks_chunk", "nodey_resampler_segment", "nodey_preview_pack", "nodey_gain_tracks", "nodey_resampler_producible", "nodey_resampler_flush_reflect", "nodey_stft_frames", "nodey_stft",
    "nodey_soundtouch_create", "nodey_soundtouch_destroy", "nodey_soundtouch_info", "nodey_soundtouch_out_frames",
    "nodey_soundtouch_run", "nodey_soundtouch_run_tracks", "nodey_soundtouch_reference_schedule", "nodey_soundtouch_chunks", "nodey_soundtouch_run_chunk", "nodey_soundtouch_run_tracks_chunk", "nodey_soundtouch_set_cluster", "nodey_soundtouch_set_candidates_per_thread", "nodey_soundtouch_set_unfused", "nodey_amix_plan", "nodey_profile_enable", "nodey_profile_launches",
    "nodey_profile_report",
    "nodey_set_device", "nodey_get_device", "nodey_device_count", "nodey_device_synchronize", "nodey_stream_create", "nodey_stream_create_priority", "nodey_stream_destroy",
    "nodey_stream_synchronize", "nodey_event_create", "nodey_event_destroy", "nodey_event_record",
    "nodey_event_synchronize", "nodey_event_elapsed_ms", "nodey_stream_wait_event", "nodey_malloc", "nodey_free", "nodey_trim_memory", "nodey_memory_stats", "nodey_memory_reserved", "nodey_set_memory_policy",
    "nodey_bus_nccl_version", "nodey_bus_unique_id", "nodey_bus_create", "nodey_bus_destroy", "nodey_bus_info",
    "nodey_bus_reduce", "nodey_bus_allreduce",
    "nodey_peer_alloc", "nodey_peer_free", "nodey_peer_export", "nodey_peer_open", "nodey_peer_close",
    "nodey_memset", "nodey_memcpy_h2d", "nodey_memcpy2d_h2d", "nodey_memcpy_d2h", "nodey_memcpy_d2d", "nodey_host_alloc", "nodey_host_free",
]


class NodeyError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"nodey error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.nodey_version.restype = i32
    L.nodey_last_error.restype = C.c_char_p
    L.nodey_device_info.argtypes = [C.POINTER(i32)] * 3 + [C.POINTER(i64)]
    L.nodey_synth.argtypes = [vp, vp, i64, i32, i32, i32, i64, vp]
    L.nodey_gain.argtypes = [vp, vp, i32, i64, C.c_float, vp]
    L.nodey_extract_interleaved.argtypes = [vp, vp, vp, i32, i64, i32, vp]
    L.nodey_split.argtypes = [vp, vp, vp, vp, i32, i64, vp]
    L.nodey_to_fltp_stereo.argtypes = [vp, vp, vp, vp, i32, i32, i64, vp]
    L.nodey_mix.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_float), i32, i64, vp]
    L.nodey_mix_gains.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_float), C.POINTER(C.c_float), i32, i64, vp]
    L.nodey_bimix.argtypes = [vp, vp, vp, vp, i64, vp, vp, i64, C.c_float, i64, vp]
    L.nodey_downmix_half.argtypes = [vp, vp, vp, i64, vp]
    L.nodey_merge_segments.argtypes = [vp, vp, vp] + [C.POINTER(i64)] * 4 + [i32, vp]
    L.nodey_resampler_create.argtypes = [C.POINTER(vp), i32, i32, i32]
    L.nodey_resampler_destroy.argtypes = [vp]
    L.nodey_resampler_destroy.restype = None
    L.nodey_resampler_info.argtypes = [vp, C.POINTER(i32)]
    L.nodey_resampler_filter_bank.argtypes = [vp]
    L.nodey_resampler_filter_bank.restype = C.POINTER(C.c_float)
    L.nodey_resampler_out_count.argtypes = [vp, i64, i32]
    L.nodey_resampler_out_count.restype = i64
    L.nodey_resampler_run.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64, i32, i64, vp]
    L.nodey_resampler_run_mode.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64, i32, i64, i32, vp]
    L.nodey_resample_tracks.argtypes = [vp, vp, vp, i64, C.POINTER(vp), C.POINTER(vp), i32, i32, i64, C.POINTER(C.c_float), i32, i32, i64, i64, vp]
    L.nodey_resample_tracks_chunks.argtypes = [vp, i32, i64, i64, i32, C.POINTER(i64), C.POINTER(i64), i32]
    L.nodey_resample_tracks_chunk.argtypes = [vp, vp, vp, i64, C.POINTER(vp), C.POINTER(vp), i32, i32, i64, C.POINTER(C.c_float), i32, i32, i64, i64, i32, i32, vp]
    L.nodey_preview_pack.argtypes = [vp, vp, vp, i64, vp]
    L.nodey_gain_tracks.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_float), i32, i32, vp]
    L.nodey_resampler_segment.argtypes = [vp, i64, i64, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.nodey_resample_mix.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32),
                                     C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_float), i32, i32, i64, vp]
    L.nodey_stft_frames.argtypes = [i64, i32, i32]
    L.nodey_stft_frames.restype = i64
    L.nodey_stft.argtypes = [vp, vp, i64, i32, i32, i64, i32, i32, vp]
    L.nodey_soundtouch_create.argtypes = [C.POINTER(vp), i32, i32, C.c_float, C.c_float]
    L.nodey_soundtouch_destroy.argtypes = [vp]
    L.nodey_soundtouch_destroy.restype = None
    L.nodey_soundtouch_info.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_double)]
    L.nodey_soundtouch_out_frames.argtypes = [vp, i64, i32, C.POINTER(i64)]
    L.nodey_soundtouch_out_frames.restype = i64
    L.nodey_soundtouch_set_cluster.argtypes = [vp, i32]
    L.nodey_soundtouch_set_candidates_per_thread.argtypes = [vp, i32]
    L.nodey_soundtouch_set_unfused.argtypes = [vp, i32]
    L.nodey_soundtouch_run.argtypes = [vp, vp, i64, vp, i64, i32, i64, i32, i64, vp, i64, vp]
    L.nodey_soundtouch_run_tracks.argtypes = [vp, vp, i64, vp, vp, i32, i64, i32, i64, vp, i64, vp]
    L.nodey_soundtouch_reference_schedule.argtypes = [vp, i64, i32, C.c_float, C.POINTER(i64), C.POINTER(i64), i64, C.POINTER(i64), C.POINTER(i32)]
    L.nodey_soundtouch_reference_schedule.restype = i64
    L.nodey_soundtouch_chunks.argtypes = [vp, i64, i32, i64, i32, C.POINTER(i64), C.POINTER(i64), i32]
    L.nodey_soundtouch_run_chunk.argtypes = [vp, vp, i64, vp, i64, i32, i64, i32, i64, vp, i64, i32, i32, i32, vp]
    L.nodey_soundtouch_run_tracks_chunk.argtypes = [vp, vp, i64, vp, vp, i32, i64, i32, i64, vp, i64, i32, i32, i32, vp]
    L.nodey_amix_plan.argtypes = [C.POINTER(i32), i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), i32,
                                  C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), i64, C.POINTER(i64),
                                  C.POINTER(i64), C.POINTER(i64), i64, C.POINTER(i64)]
    L.nodey_amix_plan.restype = i64
    L.nodey_profile_enable.argtypes = [i32]
    L.nodey_profile_enable.restype = None
    L.nodey_profile_launches.restype = C.c_uint64
    L.nodey_profile_report.argtypes = [C.c_char_p, i32]
    L.nodey_bus_nccl_version.argtypes = [C.POINTER(i32)]
    L.nodey_bus_unique_id.argtypes = [vp]
    L.nodey_bus_create.argtypes = [C.POINTER(vp), vp, i32, i32]
    L.nodey_bus_destroy.argtypes = [vp]
    L.nodey_bus_destroy.restype = None
    L.nodey_bus_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.nodey_bus_reduce.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
    L.nodey_bus_allreduce.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    L.nodey_peer_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.nodey_peer_free.argtypes = [vp]
    L.nodey_peer_export.argtypes = [vp, vp]
    L.nodey_peer_open.argtypes = [C.POINTER(vp), vp]
    L.nodey_peer_close.argtypes = [vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise NodeyError(rc, lib().nodey_last_error().decode("utf-8", "replace"))


def _torch():
    import torch
    return torch


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


_TORCH_DT = {FMT_S16: "int16", FMT_S16P: "int16", FMT_S32: "int32", FMT_S32P: "int32",
             FMT_FLT: "float32", FMT_FLTP: "float32"}


def torch_dtype(fmt):
    return getattr(_torch(), _TORCH_DT[fmt])


def is_planar(fmt):
    return fmt >= FMT_U8P


def planes_of(x, fmt):
    """x: packed -> tensor [frames, ch]; planar -> tensor [ch, frames] (contiguous). (p0, p1, nframes, nch)"""
    assert x.is_contiguous()
    if is_planar(fmt):
        nch, n = x.shape
        return x[0], (x[1] if nch > 1 else None), n, nch
    n, nch = x.shape
    return x, None, n, nch


def device_info():
    a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    check(lib().nodey_device_info(C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return {"sm_count": a.value, "cc": (b.value, c.value), "total_mem": d.value}


def synth(nframes, nch, sample_rate, track=0, frame0=0, want_s16=False, device="cuda"):
    t = _torch()
    f = t.empty((nframes, nch), dtype=t.float32, device=device)
    s = t.empty((nframes, nch), dtype=t.int16, device=device) if want_s16 else None
    check(lib().nodey_synth(_dp(f), _dp(s), nframes, nch, sample_rate, track, frame0, _stream()))
    return (f, s) if want_s16 else f


def gain(src, fmt, volume, out=None):
    t = _torch()
    dst = t.empty_like(src) if out is None else out
    check(lib().nodey_gain(_dp(dst), _dp(src), fmt, src.numel(), C.c_float(volume), _stream()))
    return dst


def extract_interleaved(x, fmt):
    t = _torch()
    p0, p1, n, nch = planes_of(x, fmt)
    out = t.empty((n, nch), dtype=t.float32, device=x.device)
    check(lib().nodey_extract_interleaved(_dp(out), _dp(p0), _dp(p1), fmt, n, nch, _stream()))
    return out


def split(x, fmt):
    t = _torch()
    p0, p1, n, nch = planes_of(x, fmt)
    assert nch == 2
    l = t.empty(n, dtype=x.dtype, device=x.device)
    r = t.empty(n, dtype=x.dtype, device=x.device)
    check(lib().nodey_split(_dp(l), _dp(r), _dp(p0), _dp(p1), fmt, n, _stream()))
    return l, r


def to_fltp_stereo(x, fmt):
    t = _torch()
    p0, p1, n, nch = planes_of(x, fmt)
    out = t.empty((2, n), dtype=t.float32, device=x.device)
    check(lib().nodey_to_fltp_stereo(_dp(out[0]), _dp(out[1]), _dp(p0), _dp(p1), fmt, nch, n, _stream()))
    return out


def mix(inputs, volumes, nframes=None):
    """inputs: list of [2, len_i] float32 planar tensors; zeros past each input's own length."""
    t = _torch()
    nin = len(inputs)
    n = max(int(x.shape[1]) for x in inputs) if nframes is None else nframes
    out = t.empty((2, n), dtype=t.float32, device=inputs[0].device)
    pl = (C.c_void_p * nin)(*[x[0].data_ptr() for x in inputs])
    pr = (C.c_void_p * nin)(*[x[1].data_ptr() for x in inputs])
    ln = (C.c_int64 * nin)(*[int(x.shape[1]) for x in inputs])
    vol = (C.c_float * nin)(*[float(v) for v in volumes])
    check(lib().nodey_mix(_dp(out[0]), _dp(out[1]), pl, pr, ln, vol, nin, n, _stream()))
    return out


def bimix(left, right, bias, nframes=None):
    t = _torch()
    n = max(left.shape[1], right.shape[1]) if nframes is None else nframes
    out = t.empty((2, n), dtype=t.float32, device=left.device)
    check(lib().nodey_bimix(_dp(out[0]), _dp(out[1]), _dp(left[0]), _dp(left[1]), left.shape[1],
                            _dp(right[0]), _dp(right[1]), right.shape[1], C.c_float(bias), n, _stream()))
    return out


def downmix_half(x):
    t = _torch()
    out = t.empty(x.shape[1], dtype=t.float32, device=x.device)
    check(lib().nodey_downmix_half(_dp(out), _dp(x[0]), _dp(x[1]), x.shape[1], _stream()))
    return out


def merge_segments(left, right, segs):
    """segs: list of (out_start, length, l_start or -1, r_start or -1)."""
    t = _torch()
    total = max(s[0] + s[1] for s in segs) if segs else 0
    out = t.empty((total, 2), dtype=t.float32, device=left.device)
    arr = [(C.c_int64 * len(segs))(*[int(s[i]) for s in segs]) for i in range(4)]
    check(lib().nodey_merge_segments(_dp(out), _dp(left), _dp(right), arr[0], arr[1], arr[2], arr[3], len(segs), _stream()))
    return out


class Resampler:
    """A7: swr_convert with library defaults (44.1 k -> 48 k: 160 phases x 32 taps)."""

    def __init__(self, in_rate, out_rate=48000, quirk=0):
        self.h = C.c_void_p()
        check(lib().nodey_resampler_create(C.byref(self.h), in_rate, out_rate, quirk))
        self.in_rate, self.out_rate = in_rate, out_rate

    def close(self):
        if getattr(self, "h", None):
            try:
                lib().nodey_resampler_destroy(self.h)
            except TypeError:        # interpreter shutdown: module globals are already gone
                pass
            self.h = None

    __del__ = close

    def info(self):
        a = (C.c_int * 8)()
        check(lib().nodey_resampler_info(self.h, a))
        keys = ["phase_count", "filter_length", "filter_alloc", "dst_incr_div", "dst_incr_mod", "src_incr",
                "index0", "linear"]
        return dict(zip(keys, list(a)))

    def filter_bank(self):
        p = self.info()
        ptr = lib().nodey_resampler_filter_bank(self.h)
        n = (p["phase_count"] + 1) * p["filter_alloc"]
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(p["phase_count"] + 1, p["filter_alloc"]).copy()

    def out_count(self, in_frames, flush=True):
        return lib().nodey_resampler_out_count(self.h, in_frames, 1 if flush else 0)

    def run(self, x, fmt, flush=True, out_frames=None, mode=0, out=None):
        t = _torch()
        p0, p1, n, nch = planes_of(x, fmt)
        m = self.out_count(n, flush) if out_frames is None else out_frames
        if out is None:
            out = t.empty((2, m), dtype=t.float32, device=x.device)
        check(lib().nodey_resampler_run_mode(self.h, _dp(out[0]), _dp(out[1]), _dp(p0), _dp(p1), fmt, nch, n,
                                             1 if flush else 0, m, mode, _stream()))
        return out

    def resample_mix(self, inputs, fmts, volumes, flush=True, out_lens=None, out_frames=None, out=None):
        t = _torch()
        nin = len(inputs)
        pls = [planes_of(x, f) for x, f in zip(inputs, fmts)]
        lens = [self.out_count(p[2], flush) for p in pls] if out_lens is None else list(out_lens)
        m = max(lens) if out_frames is None else out_frames
        if out is None:
            out = t.empty((2, m), dtype=t.float32, device=inputs[0].device)
        p0 = (C.c_void_p * nin)(*[p[0].data_ptr() for p in pls])
        p1 = (C.c_void_p * nin)(*[(p[1].data_ptr() if p[1] is not None else 0) for p in pls])
        fm = (C.c_int * nin)(*fmts)
        ch = (C.c_int * nin)(*[p[3] for p in pls])
        inf = (C.c_int64 * nin)(*[p[2] for p in pls])
        ol = (C.c_int64 * nin)(*lens)
        vol = (C.c_float * nin)(*[float(v) for v in volumes])
        check(lib().nodey_resample_mix(self.h, _dp(out[0]), _dp(out[1]), p0, p1, fm, ch, inf, ol, vol, nin,
                                       1 if flush else 0, m, _stream()))
        return out


def _resample_tracks(self, inputs, fmt, volumes, flush=True, out_len=None, out_frames=None):
    """batch of independent single-input mixers (audio_amix(1) per track) in one launch -> [ntracks, 2, out_frames]"""
    t = _torch()
    ntr = len(inputs)
    pls = [planes_of(x, fmt) for x in inputs]
    n, nch = pls[0][2], pls[0][3]
    ol = self.out_count(n, flush) if out_len is None else out_len
    m = ol if out_frames is None else out_frames
    stride = (m + 63) & ~63
    out = t.empty((ntr, 2, stride), dtype=t.float32, device=inputs[0].device)
    p0 = (C.c_void_p * ntr)(*[p[0].data_ptr() for p in pls])
    p1 = (C.c_void_p * ntr)(*[(p[1].data_ptr() if p[1] is not None else 0) for p in pls])
    vol = (C.c_float * ntr)(*[float(v) for v in volumes])
    check(lib().nodey_resample_tracks(self.h, _dp(out[0, 0]), _dp(out[0, 1]), out.stride(0), p0, p1, fmt, nch, n, vol, ntr,
                                      1 if flush else 0, ol, m, _stream()))
    return out[:, :, :m]


Resampler.resample_tracks = _resample_tracks


def _resample_tracks_chunked(self, inputs, fmt, volumes, nchunks, flush=True, poison=None):
    """nodey_resample_tracks cut into chunk launches -> ([ntracks, 2, out_frames], [(in_need, out_ready), ...]).
    poison(c, in_need): optional hook before chunk c (tests overwrite the input beyond in_need)."""
    t = _torch()
    ntr = len(inputs)
    pls = [planes_of(x, fmt) for x in inputs]
    n, nch = pls[0][2], pls[0][3]
    m = self.out_count(n, flush)
    a = (C.c_int64 * max(nchunks, 1))(); b = (C.c_int64 * max(nchunks, 1))()
    k = lib().nodey_resample_tracks_chunks(self.h, nch, n, m, nchunks, a, b, max(nchunks, 1))
    if k < 0:
        raise NodeyError(k, "nodey_resample_tracks_chunks")
    plan = [(a[c], b[c]) for c in range(k)]
    stride = (m + 63) & ~63
    out = t.full((ntr, 2, stride), float("nan"), dtype=t.float32, device=inputs[0].device)
    p0 = (C.c_void_p * ntr)(*[p[0].data_ptr() for p in pls])
    p1 = (C.c_void_p * ntr)(*[(p[1].data_ptr() if p[1] is not None else 0) for p in pls])
    vol = (C.c_float * ntr)(*[float(v) for v in volumes])
    for c in range(k):
        if poison is not None:
            poison(c, plan[c][0])
        check(lib().nodey_resample_tracks_chunk(self.h, _dp(out[0, 0]), _dp(out[0, 1]), out.stride(0), p0, p1, fmt, nch, n, vol, ntr,
                                                1 if flush else 0, m, m, c, k, _stream()))
    return out[:, :, :m], plan


Resampler.resample_tracks_chunked = _resample_tracks_chunked


def _resampler_segment(self, n_in, k0, k1):
    """input slice and recipe for outputs [k0, k1) of the flushed conversion of n_in frames (host arithmetic only)"""
    in0, in1, skip, flush = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
    check(lib().nodey_resampler_segment(self.h, n_in, k0, k1, C.byref(in0), C.byref(in1), C.byref(skip), C.byref(flush)))
    return dict(in0=in0.value, in1=in1.value, skip=skip.value, flush=bool(flush.value))


Resampler.segment = _resampler_segment


def stft_frames(n, nfft=4096, hop=1024):
    return lib().nodey_stft_frames(n, nfft, hop)


def stft(x, interleaved, nfft=4096, hop=1024, out=None):
    """x: [frames, nch] (interleaved) or [nch, frames] (planar) float32 -> complex64 [nch, M, nfft/2+1]."""
    t = _torch()
    if interleaved:
        n, nch = x.shape
        stride = 0
    else:
        nch, n = x.shape
        stride = x.stride(0)
    m = stft_frames(n, nfft, hop)
    if out is None:
        out = t.empty((nch, m, nfft // 2 + 1), dtype=t.complex64, device=x.device)
    check(lib().nodey_stft(_dp(out), _dp(x), n, nch, 1 if interleaved else 0, stride, nfft, hop, _stream()))
    return out


def uniform_runs(nframes, frame_size=1152):
    """frame sizes of a stream cut into frame_size chunks, run-length encoded [(len, count), ...]"""
    q, r = divmod(int(nframes), int(frame_size))
    runs = []
    if q:
        runs.append((frame_size, q))
    if r:
        runs.append((r, 1))
    return runs


def amix_plan(in_rates, in_runs, quirk=0, seg_cap=4096, run_cap=4096):
    """Host-only bookkeeping of audio_amix.  in_runs[i] = [(frame_len, count), ...].
    Returns (total_out_frames, [(input, out_start, src_start, len), ...], out_runs)."""
    nin = len(in_rates)
    off = [0]
    rl, rc = [], []
    for runs in in_runs:
        for l, c in runs:
            rl.append(int(l)); rc.append(int(c))
        off.append(len(rl))
    si = (C.c_int32 * seg_cap)(); so = (C.c_int64 * seg_cap)(); ss = (C.c_int64 * seg_cap)(); sl = (C.c_int64 * seg_cap)()
    orl = (C.c_int64 * run_cap)(); orc = (C.c_int64 * run_cap)()
    nseg = C.c_int64(); nrun = C.c_int64()
    n = max(len(rl), 1)
    total = lib().nodey_amix_plan((C.c_int * nin)(*in_rates), nin, (C.c_int64 * (nin + 1))(*off),
                                  (C.c_int64 * n)(*rl), (C.c_int64 * n)(*rc), quirk, si, so, ss, sl, seg_cap,
                                  C.byref(nseg), orl, orc, run_cap, C.byref(nrun))
    if total < 0:
        raise NodeyError(total, lib().nodey_last_error().decode())
    if nseg.value > seg_cap or nrun.value > run_cap:
        return amix_plan(in_rates, in_runs, quirk, max(seg_cap, nseg.value), max(run_cap, nrun.value))
    return (total, [(si[k], so[k], ss[k], sl[k]) for k in range(nseg.value)],
            [(orl[k], orc[k]) for k in range(nrun.value)])


def memory_stats(reset_peak=False):
    """(live bytes, peak bytes) of device memory held through the library's allocator"""
    live, peak = C.c_int64(), C.c_int64()
    check(lib().nodey_memory_stats(C.byref(live), C.byref(peak), 1 if reset_peak else 0))
    return live.value, peak.value


def memory_reserved():
    """(reserved bytes, peak reserved bytes): device memory obtained from the driver through the library's allocator"""
    r, p = C.c_int64(), C.c_int64()
    check(lib().nodey_memory_reserved(C.byref(r), C.byref(p)))
    return r.value, p.value


def profile_enable(on):
    lib().nodey_profile_enable(1 if on else 0)


def profile_launches():
    return int(lib().nodey_profile_launches())


def profile_report():
    import json
    buf = C.create_string_buffer(1 << 16)
    lib().nodey_profile_report(buf, len(buf))
    return json.loads(buf.value.decode())


class SoundTouch:
    """A9: the SoundTouch object of pitch_modifier / velocity_modifier as a whole-track batch op."""

    def __init__(self, sample_rate, channels, rate=1.0, pitch=1.0):
        self.h = C.c_void_p()
        check(lib().nodey_soundtouch_create(C.byref(self.h), sample_rate, channels, C.c_float(rate), C.c_float(pitch)))
        self.ch = channels

    @classmethod
    def pitch_node(cls, sample_rate, channels, semitones):
        return cls(sample_rate, channels, 1.0, float(np.float32(2.0) ** (np.float32(semitones) / np.float32(12.0))))

    @classmethod
    def velocity_node(cls, sample_rate, channels, velocity, keep_pitch):
        p = float(np.float32(1) / np.float32(velocity)) if keep_pitch else 1.0
        return cls(sample_rate, channels, velocity, p)

    def close(self):
        if getattr(self, "h", None):
            try:
                lib().nodey_soundtouch_destroy(self.h)
            except TypeError:        # interpreter shutdown: module globals are already gone
                pass
            self.h = None

    __del__ = close

    def set_cluster(self, cluster):
        check(lib().nodey_soundtouch_set_cluster(self.h, cluster))

    def set_candidates_per_thread(self, kt):
        check(lib().nodey_soundtouch_set_candidates_per_thread(self.h, kt))

    def set_unfused(self, unfused):
        check(lib().nodey_soundtouch_set_unfused(self.h, 1 if unfused else 0))

    def info(self):
        a = (C.c_int * 8)(); d = (C.c_double * 3)()
        check(lib().nodey_soundtouch_info(self.h, a, d))
        keys = ["overlap", "seek_window", "seek_length", "sample_req", "tdstretch_first", "prefill", "channels",
                "sample_rate"]
        out = dict(zip(keys, list(a)))
        out.update(rate=d[0], tempo=d[1], nominal_skip=d[2])
        return out

    def run_tracks(self, tracks, frame_size=1152, want_offsets=False):
        """tracks: list of [frames, ch] interleaved tensors or (left, right) plane pairs, read in place (no staging copy)."""
        t = _torch()
        ntr = len(tracks)
        planar = isinstance(tracks[0], (tuple, list))
        first = tracks[0][0] if planar else tracks[0]
        n = first.shape[0]
        m, nseq = self.out_frames(n, frame_size)
        out = t.empty((ntr, m, self.ch), dtype=t.float32, device=first.device)
        offs = t.zeros((ntr, max(nseq - 1, 1)), dtype=t.int32, device=first.device) if want_offsets else None
        pa = (C.c_void_p * ntr)(*[(tr[0] if planar else tr).data_ptr() for tr in tracks])
        pb = (C.c_void_p * ntr)(*[tr[1].data_ptr() for tr in tracks]) if planar else None
        check(lib().nodey_soundtouch_run_tracks(self.h, _dp(out), out.stride(0), pa, pb, ntr, n, frame_size, m,
                                                _dp(offs), offs.stride(0) if offs is not None else 0, _stream()))
        if want_offsets:
            return out, offs[:, :max(nseq - 1, 0)]
        return out

    def out_frames(self, in_frames, frame_size=1152):
        nseq = C.c_int64()
        n = lib().nodey_soundtouch_out_frames(self.h, in_frames, frame_size, C.byref(nseq))
        if n < 0:
            raise NodeyError(n, "nodey_soundtouch_out_frames")
        return n, nseq.value

    def reference_schedule(self, in_frames, velocity, frame_size=1152, cap=1 << 16):
        """(total frames, [(frame size, count), ...], flushed) of the reference's node loop (App. C7)"""
        a = (C.c_int64 * cap)(); b = (C.c_int64 * cap)()
        n, fl = C.c_int64(), C.c_int()
        total = lib().nodey_soundtouch_reference_schedule(self.h, in_frames, frame_size, C.c_float(velocity), a, b, cap, C.byref(n), C.byref(fl))
        if total < 0:
            raise NodeyError(total, "nodey_soundtouch_reference_schedule")
        return total, [(a[k], b[k]) for k in range(min(n.value, cap))], bool(fl.value)

    def chunks(self, in_frames, frame_size=1152, want=8):
        """[(in_need, out_ready), ...]: chunk c may run once in_need input frames are final and makes out_ready output frames final"""
        m, _ = self.out_frames(in_frames, frame_size)
        a = (C.c_int64 * max(want, 1))(); b = (C.c_int64 * max(want, 1))()
        n = lib().nodey_soundtouch_chunks(self.h, in_frames, frame_size, m, want, a, b, max(want, 1))
        if n < 0:
            raise NodeyError(n, "nodey_soundtouch_chunks")
        return [(a[k], b[k]) for k in range(n)]

    def run_chunked(self, x, nchunks, frame_size=1152, out=None, poison=None):
        """the render of run() cut into nchunks launches (nodey_soundtouch_run_chunk).  x: [ntracks, frames, ch].
        poison(c, in_need): optional hook called before chunk c (tests overwrite the input beyond in_need to prove that the
        chunk does not read it).  Returns (out, offsets, chunk plan)."""
        t = _torch()
        assert x.dim() == 3 and x.is_contiguous() and x.shape[2] == self.ch
        ntr, n, ch = x.shape
        m, nseq = self.out_frames(n, frame_size)
        plan = self.chunks(n, frame_size, nchunks)
        if out is None:
            out = t.empty((ntr, m, ch), dtype=t.float32, device=x.device)
        offs = t.zeros((ntr, max(nseq - 1, 1)), dtype=t.int32, device=x.device)
        for c in range(len(plan)):
            if poison is not None:
                poison(c, plan[c][0])
            check(lib().nodey_soundtouch_run_chunk(self.h, _dp(out), out.stride(0), _dp(x), x.stride(0), ntr, n, frame_size, m,
                                                   _dp(offs), offs.stride(0), c, len(plan), 0, _stream()))
        return out, offs[:, :max(nseq - 1, 0)], plan

    def run(self, x, frame_size=1152, want_offsets=False, out=None):
        """x: [ntracks, frames, ch] or [frames, ch] float32 (interleaved). Returns same rank."""
        t = _torch()
        single = x.dim() == 2
        xb = x.unsqueeze(0) if single else x
        assert xb.is_contiguous() and xb.shape[2] == self.ch
        ntr, n, ch = xb.shape
        m, nseq = self.out_frames(n, frame_size)
        if out is None:
            out = t.empty((ntr, m, ch), dtype=t.float32, device=x.device)
        offs = t.zeros((ntr, max(nseq - 1, 1)), dtype=t.int32, device=x.device) if want_offsets else None
        check(lib().nodey_soundtouch_run(self.h, _dp(out), out.stride(0), _dp(xb), xb.stride(0), ntr, n, frame_size, m,
                                         _dp(offs), offs.stride(0) if offs is not None else 0, _stream()))
        res = out[0] if single else out
        if want_offsets:
            o = offs[:, :max(nseq - 1, 0)]
            return res, (o[0] if single else o)
        return res


BUS_ID_BYTES = 128


def bus_unique_id():
    """the 128 bytes rank 0 hands to every rank before Bus(...) (nodey_bus_unique_id)"""
    buf = C.create_string_buffer(BUS_ID_BYTES)
    check(lib().nodey_bus_unique_id(buf))
    return buf.raw


class Bus:
    """Master-bus reduce across the ranks of one box (nodey_bus_*): NCCL sum of the partial FLTP buses.
    Every rank constructs it with the same id after selecting its device; the call returns when all have joined."""

    def __init__(self, unique_id, rank, nranks):
        assert len(unique_id) == BUS_ID_BYTES
        self.h = C.c_void_p()
        self.rank, self.nranks = rank, nranks
        check(lib().nodey_bus_create(C.byref(self.h), C.c_char_p(bytes(unique_id)), rank, nranks))

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.nodey_bus_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the library or the CUDA context may be gone already
            pass

    def info(self):
        r, n, d = C.c_int(), C.c_int(), C.c_int()
        check(lib().nodey_bus_info(self.h, C.byref(r), C.byref(n), C.byref(d)))
        return {"rank": r.value, "nranks": n.value, "device": d.value}

    def reduce_ptrs(self, send_l, send_r, recv_l, recv_r, nframes, root=0, stream=None):
        """raw device addresses (ints; 0 / None = absent plane)"""
        vp = lambda a: C.c_void_p(a) if a else None
        check(lib().nodey_bus_reduce(self.h, vp(send_l), vp(send_r), vp(recv_l), vp(recv_r), nframes, root,
                                     stream if stream is not None else _stream()))

    def reduce(self, send, recv, root=0):
        """send / recv: planar [nch, frames] float32 CUDA tensors (nch 1 or 2); recv is written on the root only"""
        nch, n = send.shape
        assert send.stride(1) == 1 and (recv is None or recv.stride(1) == 1)
        self.reduce_ptrs(send[0].data_ptr(), send[1].data_ptr() if nch == 2 else 0,
                         recv[0].data_ptr() if recv is not None else 0,
                         recv[1].data_ptr() if (recv is not None and nch == 2) else 0, n, root)
        return recv

    def allreduce(self, send, recv):
        nch, n = send.shape
        assert send.stride(1) == 1 and recv.stride(1) == 1
        check(lib().nodey_bus_allreduce(self.h, _dp(send[0]), _dp(send[1]) if nch == 2 else None,
                                        _dp(recv[0]), _dp(recv[1]) if nch == 2 else None, n, _stream()))
        return recv


PEER_HANDLE_BYTES = 64


class PeerBlock:
    """Device block other processes of the box can map (nodey_peer_*): .ptr, .handle (64 bytes to hand round)."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        check(lib().nodey_peer_alloc(C.byref(p), nbytes))
        self.ptr, self.nbytes = p.value, nbytes
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        check(lib().nodey_peer_export(C.c_void_p(self.ptr), buf))
        self.handle = buf.raw

    def close(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.nodey_peer_free(C.c_void_p(self.ptr))
            self.ptr = None

    __del__ = close


def peer_open(handle):
    """device address (int) of another rank's PeerBlock in this process; nodey.peer_close() when done"""
    assert len(handle) == PEER_HANDLE_BYTES
    p = C.c_void_p()
    check(lib().nodey_peer_open(C.byref(p), C.c_char_p(bytes(handle))))
    return p.value


def peer_close(ptr):
    check(lib().nodey_peer_close(C.c_void_p(ptr)))


def mix_ptrs(out_l, out_r, in_l, in_r, lens, volumes, nframes, stream=None):
    """nodey_mix on raw device addresses (ints): local or peer-mapped inputs alike"""
    nin = len(in_l)
    pl = (C.c_void_p * nin)(*in_l)
    pr = (C.c_void_p * nin)(*in_r)
    ln = (C.c_int64 * nin)(*[int(v) for v in lens])
    vol = (C.c_float * nin)(*[float(v) for v in volumes])
    check(lib().nodey_mix(C.c_void_p(out_l), C.c_void_p(out_r), pl, pr, ln, vol, nin, nframes,
                          stream if stream is not None else _stream()))
