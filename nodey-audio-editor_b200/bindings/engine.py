"""ctypes binding of libnodey_host.so (include/nodey_engine.h): the C++ host layer -- Processor plugin
API, Graph/project JSON, level-batched Runner -- as seen by a non-C++ caller.  Used by tests/ and
bench.py; no fallback path."""
import ctypes as C
import json
import os

import numpy as np

import nodey

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libnodey_host.so")

E_INVALID, E_FILE, E_GRAPH, E_NODE = -1, -2, -3, -4
_lib = None


class EngineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"engine error {code}: {message}")
        self.code, self.message = code, message


def lib():
    global _lib
    if _lib is not None:
        return _lib
    nodey.lib()      # libnodey_cuda.so first (same directory, $ORIGIN rpath)
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, cp = C.c_void_p, C.c_int64, C.c_int, C.c_char_p
    L.nodey_engine_last_error.restype = cp
    L.nodey_engine_create.argtypes = [C.POINTER(vp), cp]
    L.nodey_engine_destroy.argtypes = [vp]
    L.nodey_engine_destroy.restype = None
    L.nodey_engine_serialize.argtypes = [vp, cp, i32]
    L.nodey_engine_check.argtypes = [vp]
    L.nodey_engine_node_count.argtypes = [vp]
    L.nodey_engine_node_info.argtypes = [vp, i32, C.POINTER(i32), cp, i32, C.POINTER(i32)]
    L.nodey_engine_set_volume.argtypes = [vp, i32, C.c_float]
    L.nodey_engine_bind_source.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, i64, i32, C.c_double]
    L.nodey_engine_run.argtypes = [vp]
    L.nodey_engine_product.argtypes = [vp, i32, cp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                       C.POINTER(i64), C.POINTER(C.c_double), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32)]
    L.nodey_engine_product_runs.argtypes = [vp, i32, cp, C.POINTER(i64), C.POINTER(i64), i32]
    L.nodey_engine_output.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(C.c_double),
                                      C.POINTER(vp), C.POINTER(vp)]
    L.nodey_engine_set_export_path.argtypes = [vp, cp]
    L.nodey_engine_set_preview.argtypes = [vp, i32]
    L.nodey_engine_set_export_kbps.argtypes = [vp, i32]
    L.nodey_engine_diagnostics.argtypes = [vp, cp, i32]
    L.nodey_engine_level_timings.argtypes = [vp] + [C.POINTER(i32)] * 4 + [C.POINTER(C.c_double)] * 3 + [i32]
    L.nodey_engine_encode_mp3.argtypes = [cp, vp, vp, i32, i32, i32, i64, i32, C.c_double, i32, C.POINTER(C.c_double)]
    L.nodey_engine_preview.argtypes = [vp, C.POINTER(i64), C.POINTER(vp), C.POINTER(i64), i32]
    L.nodey_engine_set_schedule.argtypes = [vp, cp, cp]
    L.nodey_engine_export_plan.argtypes = [i32, C.c_double, i32, i64, C.POINTER(i64), C.POINTER(i64), i32, C.POINTER(C.c_double),
                                           C.POINTER(i64), C.POINTER(C.c_double), i32]
    L.nodey_engine_product_stamp.argtypes = [vp, i32, cp, C.POINTER(i32), C.POINTER(C.c_double)]
    L.nodey_engine_probe_wav.argtypes = [cp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i32)]
    _lib = L
    return L


def probe_wav(path):
    """(format, sample_rate, channels, frames, frame_size) audio_input would publish for a RIFF/WAVE file; header only"""
    f, r, c, fs = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    n = C.c_int64()
    _check(lib().nodey_engine_probe_wav(os.fsencode(path), C.byref(f), C.byref(r), C.byref(c), C.byref(n), C.byref(fs)))
    return f.value, r.value, c.value, n.value, fs.value


def set_release_products(release):
    """memory policy of the next runs: True = links drop their products once consumed (only sink / spectrum survive)"""
    _check(lib().nodey_engine_set_release_products(1 if release else 0))


def register_examples():
    """registers the example processors outside the reference's set ("frame_gain_example": frame-interface node)"""
    _check(lib().nodey_engine_register_examples())


def _check(rc):
    if rc < 0:
        raise EngineError(rc, lib().nodey_engine_last_error().decode("utf-8", "replace"))
    return rc


_NP = {1: np.int16, 2: np.int32, 3: np.float32, 6: np.int16, 7: np.int32, 8: np.float32}


class Product:
    def __init__(self, kind, fmt, rate, ch, frames, pts, p0, p1, extra):
        self.kind, self.fmt, self.rate, self.ch, self.frames, self.pts = kind, fmt, rate, ch, frames, pts
        self.p0, self.p1, self.bins = p0, p1, extra

    def numpy(self):
        """device -> host copy.  audio packed: [frames, ch]; planar: [ch, frames]; spectrum: complex64 [ch, frames, bins]"""
        L = nodey.lib()

        def fetch(ptr, shape, dtype):
            out = np.empty(shape, dtype)
            if out.size:
                nodey.check(L.nodey_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes, None))
                nodey.check(L.nodey_device_synchronize())
            return out
        if self.kind == 2:
            return fetch(self.p0, (self.ch, self.frames, self.bins), np.complex64)
        dt = _NP[self.fmt]
        if self.fmt >= 5:
            planes = [fetch(self.p0, (self.frames,), dt)]
            if self.ch == 2:
                planes.append(fetch(self.p1, (self.frames,), dt))
            return np.stack(planes)
        return fetch(self.p0, (self.frames, self.ch), dt)


def mp3_available():
    return bool(lib().nodey_engine_mp3_available())


def encode_mp3(path, data, fmt, rate, frame_size=1152, pts=0.0, kbps=320, time=0.0):
    """The sink's MP3 leg on a HOST stream (numpy; packed [frames, ch] or planar [ch, frames]): do_export's LAME call
    sequence.  Returns Process_context::time afterwards."""
    data = np.ascontiguousarray(data)
    planar = fmt >= 5
    ch, frames = data.shape if planar else data.shape[::-1]
    p0 = data.ctypes.data
    p1 = data[1].ctypes.data if (planar and ch == 2) else 0
    t = C.c_double(time)
    _check(lib().nodey_engine_encode_mp3(path.encode(), C.c_void_p(p0), C.c_void_p(p1) if p1 else None, fmt, rate, ch, frames,
                                         frame_size, pts, kbps, C.byref(t)))
    return t.value


STAMP_START, STAMP_END_US, STAMP_START_FLOAT_US = 0, 1, 2


def export_plan(stamp, origin, rate, runs, frames=None, time=0.0):
    """do_export's bookkeeping for a stream cut into `runs` = [(frame size, count), ...] whose frames are stamped by rule
    `stamp` (include/nodey_engine.h): (silence samples in front of every frame, every frame's stamp in seconds, time after)"""
    n = len(runs)
    rl = (C.c_int64 * max(n, 1))(*[r[0] for r in runs])
    rc = (C.c_int64 * max(n, 1))(*[r[1] for r in runs])
    total = sum(a * b for a, b in runs) if frames is None else frames
    cap = sum(b for _, b in runs)
    sil = (C.c_int64 * max(cap, 1))()
    pts = (C.c_double * max(cap, 1))()
    t = C.c_double(time)
    k = _check(lib().nodey_engine_export_plan(stamp, origin, rate, total, rl, rc, n, C.byref(t), sil, pts, cap))
    return list(sil[:k]), list(pts[:k]), t.value


class Engine:
    def __init__(self, project):
        text = project if isinstance(project, str) else json.dumps(project)
        self.h = C.c_void_p()
        _check(lib().nodey_engine_create(C.byref(self.h), text.encode()))
        self._keep = []

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.nodey_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def serialize(self):
        n = _check(lib().nodey_engine_serialize(self.h, None, 0))
        buf = C.create_string_buffer(n + 1)
        _check(lib().nodey_engine_serialize(self.h, buf, n + 1))
        return buf.value.decode()

    def check(self):
        _check(lib().nodey_engine_check(self.h))

    def nodes(self):
        out = []
        for k in range(_check(lib().nodey_engine_node_count(self.h))):
            nid, lvl = C.c_int(), C.c_int()
            buf = C.create_string_buffer(128)
            _check(lib().nodey_engine_node_info(self.h, k, C.byref(nid), buf, 128, C.byref(lvl)))
            out.append((nid.value, buf.value.decode(), lvl.value))
        return out

    def set_volume(self, node_id, volume):
        _check(lib().nodey_engine_set_volume(self.h, node_id, volume))

    def bind_source(self, index, data, fmt, rate, frame_size=1152, pts=0.0):
        """data: numpy array (host; packed [frames, ch] or planar [ch, frames]) or a torch CUDA tensor of that shape"""
        is_torch = hasattr(data, "data_ptr")
        planar = fmt >= 5
        if planar:
            ch, frames = data.shape
        else:
            frames, ch = data.shape
        if is_torch:
            assert data.is_contiguous()
            on_dev = 1 if data.is_cuda else 0
            p0 = data.data_ptr()
            p1 = data[1].data_ptr() if (planar and ch == 2) else 0
        else:
            data = np.ascontiguousarray(data)
            on_dev = 0
            p0 = data.ctypes.data
            p1 = data[1].ctypes.data if (planar and ch == 2) else 0
        self._keep.append(data)
        _check(lib().nodey_engine_bind_source(self.h, index, C.c_void_p(p0), C.c_void_p(p1) if p1 else None, on_dev, fmt, rate,
                                              ch, frames, frame_size, pts))

    def run(self):
        _check(lib().nodey_engine_run(self.h))

    def set_export_path(self, path):
        """empty = memory only; *.wav = 32-bit float WAV; any other path = MP3 through LAME like the reference's export"""
        _check(lib().nodey_engine_set_export_path(self.h, (path or "").encode()))

    def set_export_kbps(self, kbps):
        """bit rate of an MP3 export (any export path that does not end in .wav), 8..320, default 320"""
        _check(lib().nodey_engine_set_export_kbps(self.h, int(kbps)))

    def set_preview(self, on=True):
        """the next runs take the sink's preview path (swr without flush -> clamp -> packed 48 kHz stereo float)"""
        _check(lib().nodey_engine_set_preview(self.h, 1 if on else 0))

    def preview(self):
        """(packed [frames, 2] float32 numpy array, chunk sizes) of the last preview run"""
        frames, ptr = C.c_int64(), C.c_void_p()
        cap = 1 << 16
        chunks = (C.c_int64 * cap)()
        n = _check(lib().nodey_engine_preview(self.h, C.byref(frames), C.byref(ptr), chunks, cap))
        out = np.empty((frames.value, 2), np.float32)
        if frames.value:
            nodey.check(nodey.lib().nodey_memcpy_d2h(out.ctypes.data_as(C.c_void_p), ptr, out.nbytes, None))
            nodey.check(nodey.lib().nodey_stream_synchronize(None))
        return out, [chunks[k] for k in range(min(n, cap))]

    def diagnostics(self):
        """the overlay's Audio block of the last run as text"""
        n = _check(lib().nodey_engine_diagnostics(self.h, None, 0))
        buf = C.create_string_buffer(n + 1)
        _check(lib().nodey_engine_diagnostics(self.h, buf, n + 1))
        return buf.value.decode()

    def level_timings(self):
        """[{wave, level, lane, nodes, enqueue_ms, device_ms, start_ms}, ...] of the last run, in enqueue order"""
        cap = 4096
        iv = [(C.c_int * cap)() for _ in range(4)]
        dv = [(C.c_double * cap)() for _ in range(3)]
        n = _check(lib().nodey_engine_level_timings(self.h, *iv, *dv, cap))
        keys = ("wave", "level", "lane", "nodes", "enqueue_ms", "device_ms", "start_ms")
        return [dict(zip(keys, [a[k] for a in iv] + [a[k] for a in dv])) for k in range(min(n, cap))]

    def product(self, node_id, pin):
        kind, fmt, rate, ch, extra = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        frames, pts = C.c_int64(), C.c_double()
        p0, p1 = C.c_void_p(), C.c_void_p()
        _check(lib().nodey_engine_product(self.h, node_id, pin.encode(), C.byref(kind), C.byref(fmt), C.byref(rate), C.byref(ch),
                                          C.byref(frames), C.byref(pts), C.byref(p0), C.byref(p1), C.byref(extra)))
        return Product(kind.value, fmt.value, rate.value, ch.value, frames.value, pts.value, p0.value, p1.value, extra.value)

    def product_runs(self, node_id, pin):
        a = (C.c_int64 * 4096)(); b = (C.c_int64 * 4096)()
        n = _check(lib().nodey_engine_product_runs(self.h, node_id, pin.encode(), a, b, 4096))
        return [(a[k], b[k]) for k in range(min(n, 4096))]

    def set_schedule(self, **settings):
        """infra::Runner::Schedule of this engine's runs, e.g. set_schedule(wave_pins=8, compute_lanes=2, stream_chunks=4,
        wave_pattern="32,64")"""
        for key, value in settings.items():
            text = ",".join(str(v) for v in value) if isinstance(value, (list, tuple)) else str(int(value) if isinstance(value, bool) else value)
            _check(lib().nodey_engine_set_schedule(self.h, key.encode(), text.encode()))

    def product_stamp(self, node_id, pin):
        """(rule, origin) by which the product's frames are stamped (STAMP_*)"""
        st, og = C.c_int(), C.c_double()
        _check(lib().nodey_engine_product_stamp(self.h, node_id, pin.encode(), C.byref(st), C.byref(og)))
        return st.value, og.value

    def output(self):
        fmt, rate, ch = C.c_int(), C.c_int(), C.c_int()
        frames, pts = C.c_int64(), C.c_double()
        p0, p1 = C.c_void_p(), C.c_void_p()
        _check(lib().nodey_engine_output(self.h, C.byref(fmt), C.byref(rate), C.byref(ch), C.byref(frames), C.byref(pts),
                                         C.byref(p0), C.byref(p1)))
        return Product(1, fmt.value, rate.value, ch.value, frames.value, pts.value, p0.value, p1.value, 0)


# ---------------------------------------------------------------------------------------------------
# project builders (the JSON format of src/infra/graph.cpp:284-372)
# ---------------------------------------------------------------------------------------------------
class Project:
    def __init__(self):
        self.nodes, self.links = {}, []

    def add(self, identifier, info=None, node_id=None):
        nid = len(self.nodes) if node_id is None else node_id
        self.nodes[str(nid)] = {"identifier": identifier, "info": info if info is not None else None, "position": {"x": 0.0, "y": 0.0}}
        return nid

    def link(self, a, a_pin, b, b_pin):
        self.links.append({"from": {"node": a, "pin": a_pin}, "to": {"node": b, "pin": b_pin}})

    def json(self):
        return {"nodes": self.nodes, "links": self.links}


def amix_info(volumes):
    info = {"input_num": len(volumes)}
    for i, v in enumerate(volumes):
        info[f"volumes{i}"] = float(v)
        info[f"locks{i}"] = False
    return info


def config5_project(n_tracks, gains, pitch=3.0, velocity=1.25, group_vol=1.0 / 16, master_vol=1.0 / 16, spectrum=True):
    """audio_input(n) -> [amix(1) -> pitch -> velocity -> gain] x n -> amix(16) x n/16 -> amix(n/16) -> output (+ spectrum)."""
    assert n_tracks % 16 == 0 and n_tracks // 16 <= 16
    p = Project()
    src = p.add("audio_input", {"file_path": [""] * n_tracks})
    gain_nodes = []
    for t in range(n_tracks):
        a = p.add("audio_amix", amix_info([1.0]))
        pm = p.add("pitch_modifier", {"pitch": pitch})
        vm = p.add("velocity_modifier", {"velocity": velocity, "keep_pitch": True})
        g = p.add("audio_volume_adjust", {"volume": gains[t]})
        p.link(src, f"output_{t}", a, "input_1")
        p.link(a, "output", pm, "input")
        p.link(pm, "output", vm, "input")
        p.link(vm, "output", g, "input")
        gain_nodes.append(g)
    groups = []
    for gi in range(n_tracks // 16):
        m = p.add("audio_amix", amix_info([group_vol] * 16))
        for k in range(16):
            p.link(gain_nodes[gi * 16 + k], "output", m, f"input_{k + 1}")
        groups.append(m)
    master = p.add("audio_amix", amix_info([master_vol] * len(groups)))
    for k, m in enumerate(groups):
        p.link(m, "output", master, f"input_{k + 1}")
    out = p.add("audio_output")
    p.link(master, "output", out, "input")
    spec = None
    if spectrum:
        spec = p.add("audio_spectrum", {"fft_size": 4096, "hop": 1024, "window": "hann"})
        p.link(master, "output", spec, "input")
    return p, {"input": src, "gains": gain_nodes, "groups": groups, "master": master, "output": out, "spectrum": spec}
