"""The REAL SoundTouch library driven the way the reference drives it (TEST INFRASTRUCTURE ONLY, like everything
under oracle/).  Counterpart of real_swr.py for the pitch / tempo path:

    reference:  src/processor/audio-velocity.cpp:367-435 -- new SoundTouch; setSampleRate; setChannels;
                setRate(velocity); setPitch(pitch); putSamples per 1152-sample frame; receiveSamples of
                min(numSamples, 3 * 1152 / velocity) whenever more than 1152 / velocity are queued; flush() at the end.
    library:    SoundTouch 2.3.2 (xmake.lua:16), float-sample build.

No SoundTouch binary or source exists in this image, so nothing here has run against the real library yet: the
SoundTouch half of the oracle stays PARITY UNPINNED until someone points NODEY_REAL_SOUNDTOUCH at one and runs
tests/test_st_real.py (and, to keep the result, tests/golden/make_st_golden.py).  What can be checked here is the
harness itself: tests/fake_soundtouch builds a stand-in library with SoundTouchDLL's C entry points on top of the
oracle's streaming model, and the whole chain (binding, driving loops, comparisons, fixture round trip) runs against it.

Two ways to reach the library:
  * SoundTouchDLL's C API (libSoundTouchDll.so / SoundTouchDLL.dll: soundtouch_createInstance, soundtouch_setRate, ...,
    source/SoundTouchDLL/SoundTouchDLL.h of the SoundTouch distribution) -- bound directly with ctypes;
  * a plain libSoundTouch.so (C++ symbols only): build the 40-line shim oracle/st_shim.cpp against the library's
    headers (`make -C oracle st_shim SOUNDTOUCH_INC=... SOUNDTOUCH_LIB=...`), which exports the same C names, and
    point NODEY_REAL_SOUNDTOUCH at oracle/_ref/libnodey_st_shim.so.
"""
import ctypes as C
import os

import numpy as np

ENV = "NODEY_REAL_SOUNDTOUCH"
_lib = None

# setting ids of SoundTouch.h (SETTING_*), for the record of what the defaults were
SETTINGS = {"USE_AA_FILTER": 0, "AA_FILTER_LENGTH": 1, "USE_QUICKSEEK": 2, "SEQUENCE_MS": 3, "SEEKWINDOW_MS": 4, "OVERLAP_MS": 5,
            "NOMINAL_INPUT_SEQUENCE": 6, "NOMINAL_OUTPUT_SEQUENCE": 7, "INITIAL_LATENCY": 8}


def library_path():
    return os.environ.get(ENV, "")


def available():
    try:
        return lib() is not None
    except OSError:
        return False


def lib():
    """the library named by NODEY_REAL_SOUNDTOUCH; None when the variable is unset; OSError when it cannot be used"""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path:
        return None
    L = C.CDLL(path)
    if not hasattr(L, "soundtouch_createInstance"):
        raise OSError(f"{path} does not export SoundTouchDLL's C API (soundtouch_createInstance): for a plain "
                      "libSoundTouch.so build oracle/st_shim.cpp and use that (see oracle/real_soundtouch.py)")
    vp, u32, f32 = C.c_void_p, C.c_uint, C.c_float
    L.soundtouch_createInstance.restype = vp
    L.soundtouch_destroyInstance.argtypes = [vp]
    L.soundtouch_destroyInstance.restype = None
    L.soundtouch_getVersionString.restype = C.c_char_p
    L.soundtouch_getVersionId.restype = u32
    for name in ("soundtouch_setRate", "soundtouch_setTempo", "soundtouch_setPitch"):
        getattr(L, name).argtypes = [vp, f32]
        getattr(L, name).restype = None
    L.soundtouch_setChannels.argtypes = [vp, u32]
    L.soundtouch_setSampleRate.argtypes = [vp, u32]
    L.soundtouch_flush.argtypes = [vp]
    L.soundtouch_flush.restype = None
    L.soundtouch_putSamples.argtypes = [vp, vp, u32]
    L.soundtouch_receiveSamples.argtypes = [vp, vp, u32]
    L.soundtouch_receiveSamples.restype = u32
    L.soundtouch_numSamples.argtypes = [vp]
    L.soundtouch_numSamples.restype = u32
    L.soundtouch_getSetting.argtypes = [vp, C.c_int]
    L.soundtouch_getSetting.restype = C.c_int
    _lib = L
    return L


def version():
    L = lib()
    return (L.soundtouch_getVersionString() or b"").decode("ascii", "replace"), int(L.soundtouch_getVersionId())


class SoundTouch:
    """one SoundTouch object configured like audio-velocity.cpp:381-385"""

    def __init__(self, sample_rate, channels, rate, pitch):
        L = lib()
        if L is None:
            raise OSError(f"{ENV} is not set")
        self.L, self.ch = L, channels
        self.h = L.soundtouch_createInstance()
        if not self.h:
            raise OSError("soundtouch_createInstance failed")
        L.soundtouch_setSampleRate(self.h, sample_rate)
        L.soundtouch_setChannels(self.h, channels)
        L.soundtouch_setRate(self.h, C.c_float(rate))
        L.soundtouch_setPitch(self.h, C.c_float(pitch))

    def settings(self):
        return {k: int(self.L.soundtouch_getSetting(self.h, v)) for k, v in SETTINGS.items()}

    def put(self, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.ch)
        if x.shape[0]:
            self.L.soundtouch_putSamples(self.h, x.ctypes.data_as(C.c_void_p), x.shape[0])

    def num_samples(self):
        return int(self.L.soundtouch_numSamples(self.h))

    def receive(self, max_frames):
        out = np.zeros((max(int(max_frames), 0), self.ch), np.float32)
        n = int(self.L.soundtouch_receiveSamples(self.h, out.ctypes.data_as(C.c_void_p), out.shape[0])) if out.shape[0] else 0
        return out[:n]

    def flush(self):
        self.L.soundtouch_flush(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.soundtouch_destroyInstance(self.h)
            self.h = None

    __del__ = close


def frames_of(x, frame_size):
    return [x[k:k + frame_size] for k in range(0, x.shape[0], frame_size)]


def run_canonical(make, x, frame_size=1152):
    """the engine's canonical schedule: putSamples per frame, drain everything after each put,
    ALWAYS flush.  `make()` returns an object with put / receive / num_samples / flush (the real library, the oracle's
    streaming model, ...).  SoundTouch is FIFO driven, so the samples do not depend on the receive sizes."""
    st = make()
    outs = []
    for fr in frames_of(x, frame_size):
        st.put(fr)
        outs.append(st.receive(st.num_samples()))
    st.flush()
    outs.append(st.receive(st.num_samples()))
    return np.concatenate(outs) if outs else np.zeros((0, x.shape[1]), np.float32)


def run_reference_loop(make, x, velocity, frame_size=1152):
    """soundtouch_process_payload (audio-velocity.cpp:286-441) with a frame available at every loop turn.
    Returns (samples, receive sizes, flushed): flushed False = the loop left through the early `break` at :414 with
    SoundTouch still holding the tail (never flushed)."""
    time_ratio = float(np.float32(1.0) / np.float32(velocity))         # const double time_ratio = 1.0f / velocity;
    min_samples = int(time_ratio * 1152) & 0xFFFFFFFF
    max_samples = int(time_ratio * 1152 * 3) & 0xFFFFFFFF
    frames = frames_of(x, frame_size)
    st, k, eof, flushed = None, 0, False, False
    outs, sizes = [], []
    while True:
        if not eof:
            if k >= len(frames):
                eof = True
            else:
                if st is None:
                    st = make()
                st.put(frames[k])
                k += 1
        if st is None:
            if eof:
                break
            continue
        if st.num_samples() == 0 and eof:
            break
        if st.num_samples() > min_samples:
            got = st.receive(min(st.num_samples(), max_samples))
            outs.append(got); sizes.append(len(got))
        elif eof:
            st.flush()
            flushed = True
            remaining = st.num_samples()
            if remaining > 0:
                got = st.receive(remaining)
                outs.append(got); sizes.append(len(got))
            break
    y = np.concatenate(outs) if outs else np.zeros((0, x.shape[1]), np.float32)
    return y, sizes, flushed
