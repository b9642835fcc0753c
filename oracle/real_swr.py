"""ctypes binding of a REAL libswresample, used only to pin the oracle's restatement (test infrastructure).

The reference links FFmpeg 7.1's libswresample (xmake.lua:12), which is not vendored.  This image happens to
carry a stock build of the same library inside the opencv-python-headless wheel
(`site-packages/opencv_python_headless.libs/libswresample-*.so.6.1.100`, FFmpeg 8.0: libswresample 6.1 /
libavutil 60.8; the resampler core -- resample.c, resample_template.c, audioconvert.c, rematrix.c -- is
the code FFmpeg 7.1 ships as libswresample 5.3).  The context is configured exactly like the reference's call
sites do (audio-amix.cpp:212-240, audio-bimix.cpp:198-240): `swr_alloc`, `av_opt_set_chlayout/int/sample_fmt`
for in/out layout, rate and format, everything else left at the library default, `swr_init`, then
`swr_convert` per frame and `swr_convert(ctx, out, n, 0, 0)` to flush.

`tests/golden/make_swr_golden.py` uses it to write `tests/golden/swr_real.npz` (outputs of the real
library); `tests/test_swr_real.py` checks the oracle against that fixture everywhere and against the
live library where it is present.  Nothing outside tests/ may import this module.
"""
import ctypes as C
import glob
import os
import site
import sys

import numpy as np

FMT_U8, FMT_S16, FMT_S32, FMT_FLT, FMT_DBL, FMT_U8P, FMT_S16P, FMT_S32P, FMT_FLTP = range(9)
_BPS = {FMT_S16: 2, FMT_S32: 4, FMT_FLT: 4, FMT_S16P: 2, FMT_S32P: 4, FMT_FLTP: 4}
_NP = {FMT_S16: np.int16, FMT_S32: np.int32, FMT_FLT: np.float32, FMT_S16P: np.int16, FMT_S32P: np.int32,
       FMT_FLTP: np.float32}


class AVChannelLayout(C.Structure):
    _fields_ = [("order", C.c_int), ("nb_channels", C.c_int), ("mask", C.c_uint64), ("opaque", C.c_void_p)]


_libs = None


def find():
    """(libavutil path, libswresample path) or None.  NODEY_REAL_AVUTIL / NODEY_REAL_SWRESAMPLE point the checks at
    another build (e.g. the FFmpeg 7.1 libraries the reference links: libavutil.so.59 / libswresample.so.5)."""
    eu, es = os.environ.get("NODEY_REAL_AVUTIL"), os.environ.get("NODEY_REAL_SWRESAMPLE")
    if eu and es and os.path.exists(eu) and os.path.exists(es):
        return eu, es
    roots = list(site.getsitepackages()) + [p for p in sys.path if p.endswith("site-packages")]
    for r in dict.fromkeys(roots):
        d = os.path.join(r, "opencv_python_headless.libs")
        u = sorted(glob.glob(os.path.join(d, "libavutil-*.so*")))
        s = sorted(glob.glob(os.path.join(d, "libswresample-*.so*")))
        if u and s:
            return u[0], s[0]
    return None


def available():
    return find() is not None


def libs():
    global _libs
    if _libs is None:
        p = find()
        if p is None:
            raise RuntimeError("no libswresample in this image (opencv_python_headless.libs not found)")
        d = os.path.dirname(p[0])
        for dep in ("libdrm-*.so*", "libcrypto-*.so*"):   # libavutil's NEEDED entries; the wheel has no RPATH
            for f in sorted(glob.glob(os.path.join(d, dep))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
        u = C.CDLL(p[0], mode=C.RTLD_GLOBAL)
        s = C.CDLL(p[1], mode=C.RTLD_GLOBAL)
        u.av_force_cpu_flags.argtypes = [C.c_int]
        u.av_get_cpu_flags.restype = C.c_int
        u.av_channel_layout_default.argtypes = [C.POINTER(AVChannelLayout), C.c_int]
        u.av_opt_set_chlayout.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(AVChannelLayout), C.c_int]
        u.av_opt_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int]
        u.av_opt_set_sample_fmt.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        u.av_opt_get_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
        u.av_opt_get_double.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
        u.av_version_info.restype = C.c_char_p
        s.swr_alloc.restype = C.c_void_p
        s.swr_init.argtypes = [C.c_void_p]
        s.swr_free.argtypes = [C.POINTER(C.c_void_p)]
        s.swr_convert.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_int]
        s.swr_get_delay.argtypes = [C.c_void_p, C.c_int64]
        s.swr_get_delay.restype = C.c_int64
        s.swr_get_out_samples.argtypes = [C.c_void_p, C.c_int]
        s.swresample_version.restype = C.c_uint
        _libs = (u, s)
    return _libs


def versions():
    u, s = libs()
    v = s.swresample_version()
    return {"ffmpeg": u.av_version_info().decode(), "libswresample": f"{v >> 16}.{(v >> 8) & 255}.{v & 255}"}


def force_c_path(on=True):
    """on: av_force_cpu_flags(0) -> swri_resample_dsp_init keeps the C template (resample_template.c)
    instead of the SSE/AVX/FMA3 assembly, whose lane-wise partial sums round differently.  Must be called
    before a context is initialised.  off: back to auto detection (-1)."""
    libs()[0].av_force_cpu_flags(0 if on else -1)


class RealSwr:
    """One SwrContext set up like the reference's: input layout mono / stereo by channel count,
    output stereo FLTP at `out_rate`, all other options default."""

    def __init__(self, in_rate, out_rate, in_fmt, in_ch, out_fmt=FMT_FLTP):
        u, s = libs()
        self.fmt, self.ch, self.out_fmt = in_fmt, in_ch, out_fmt
        self.h = C.c_void_p(s.swr_alloc())
        lin, lout = AVChannelLayout(), AVChannelLayout()
        u.av_channel_layout_default(C.byref(lin), in_ch)
        u.av_channel_layout_default(C.byref(lout), 2)
        assert u.av_opt_set_chlayout(self.h, b"in_chlayout", C.byref(lin), 0) == 0
        assert u.av_opt_set_int(self.h, b"in_sample_rate", in_rate, 0) == 0
        assert u.av_opt_set_sample_fmt(self.h, b"in_sample_fmt", in_fmt, 0) == 0
        assert u.av_opt_set_chlayout(self.h, b"out_chlayout", C.byref(lout), 0) == 0
        assert u.av_opt_set_int(self.h, b"out_sample_rate", out_rate, 0) == 0
        assert u.av_opt_set_sample_fmt(self.h, b"out_sample_fmt", out_fmt, 0) == 0
        rc = s.swr_init(self.h)
        if rc < 0:
            raise RuntimeError(f"swr_init returned {rc}")

    def option(self, name, kind="int"):
        u, _ = libs()
        if kind == "int":
            v = C.c_int64()
            assert u.av_opt_get_int(self.h, name.encode(), 0, C.byref(v)) == 0
            return v.value
        v = C.c_double()
        assert u.av_opt_get_double(self.h, name.encode(), 0, C.byref(v)) == 0
        return v.value

    def close(self):
        if self.h:
            libs()[1].swr_free(C.byref(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def convert(self, x, out_count):
        """x: None (flush), (n, ch) array for packed formats or (ch, n) for planar ones.
        Returns (left, right) float32 arrays of the length swr_convert returned (out_fmt FLTP), or the
        packed (n, 2) float array for out_fmt FLT."""
        _, s = libs()
        if self.out_fmt == FMT_FLTP:
            ol = np.zeros(out_count + 64, np.float32); orr = np.zeros(out_count + 64, np.float32)
            outp = (C.c_void_p * 2)(ol.ctypes.data, orr.ctypes.data)
        else:
            o = np.zeros((out_count + 64, 2), np.float32)
            outp = (C.c_void_p * 2)(o.ctypes.data, None)
        if x is None:
            n = s.swr_convert(self.h, outp, out_count, None, 0)
        else:
            x = np.ascontiguousarray(x, _NP[self.fmt])
            if self.fmt >= FMT_U8P:
                planes = [np.ascontiguousarray(x[c]) for c in range(self.ch)]
                nfr = planes[0].shape[0]
                inp = (C.c_void_p * 2)(*[p.ctypes.data for p in planes], *([None] * (2 - self.ch)))
            else:
                nfr = x.shape[0]
                inp = (C.c_void_p * 2)(x.ctypes.data, None)
            n = s.swr_convert(self.h, outp, out_count, inp, nfr)
        if n < 0:
            raise RuntimeError(f"swr_convert returned {n}")
        if self.out_fmt == FMT_FLTP:
            return ol[:n].copy(), orr[:n].copy()
        return o[:n].copy()


def whole(x, fmt, in_rate, out_rate, in_ch, frame=None, flush=True):
    """Feed x (layout as RealSwr.convert) in one call or in frames of `frame` samples with an output
    capacity large enough to take everything, then flush until 0.  Returns (left, right, counts)."""
    r = RealSwr(in_rate, out_rate, fmt, in_ch)
    planar = fmt >= FMT_U8P
    n = x.shape[1] if planar else x.shape[0]
    frame = frame or max(n, 1)
    L, R, counts = [], [], []
    for p in range(0, n, frame):
        xs = x[:, p:p + frame] if planar else x[p:p + frame]
        m = xs.shape[1] if planar else xs.shape[0]
        l, rr = r.convert(xs, int(m * out_rate / in_rate) + 4096)
        L.append(l); R.append(rr); counts.append(len(l))
    if flush:
        while True:
            l, rr = r.convert(None, 1 << 16)
            counts.append(len(l))
            if len(l) == 0:
                break
            L.append(l); R.append(rr)
    r.close()
    cat = lambda a: np.concatenate(a) if a else np.zeros(0, np.float32)
    return cat(L), cat(R), counts
