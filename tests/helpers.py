import numpy as np

FMT_S16, FMT_S32, FMT_FLT, FMT_S16P, FMT_S32P, FMT_FLTP = 1, 2, 3, 6, 7, 8
ALL_FMTS = [FMT_S16, FMT_S32, FMT_FLT, FMT_S16P, FMT_S32P, FMT_FLTP]


def make_input(orc, fmt, nframes, nch, rate=44100, track=0):
    """Synthetic track in the given AVSampleFormat: packed [frames, ch] or planar [ch, frames]."""
    x = orc.synth_f32(nframes, nch, rate, track)
    if fmt in (FMT_S16, FMT_S16P):
        y = orc.f32_to_s16(x)
    elif fmt in (FMT_S32, FMT_S32P):
        y = np.clip(np.rint(x.astype(np.float64) * 2147483647.0), -2147483648, 2147483647).astype(np.int32)
    else:
        y = x
    if fmt >= 5:
        y = np.ascontiguousarray(y.T)
    return y


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} vs {b.dtype}"
    if not np.array_equal(bits(a), bits(b)):
        bad = np.flatnonzero(a.reshape(-1).view(np.uint8 if a.itemsize == 1 else f"u{a.itemsize}") !=
                             b.reshape(-1).view(np.uint8 if b.itemsize == 1 else f"u{b.itemsize}"))
        i = int(bad[0])
        raise AssertionError(f"{what}: {bad.size} of {a.size} elements differ; first at {i}: "
                             f"{a.reshape(-1)[i]!r} vs {b.reshape(-1)[i]!r}")
