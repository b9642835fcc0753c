"""bench_configs.py -- BASELINE.json configs[0..3] measured next to the headline (configs[4], bench.py).

Every config runs through the reference-facing plugin API (project JSON -> Graph -> Runner -> node classes -> C ABI)
on ONE GPU and reports, as BASELINE.md section 2 asks per config:

  value / ms       realtime factor and device time with the source resident in HBM
  e2e              the same with the source in pinned host memory (H2D inside the timed region) and the result
                   copied back (D2H), bytes counted
  roofline         the dominant kernel of the render (CUDA events around every launch of a separate, profiled run):
                   achieved = the config's ALGORITHMIC bytes (unique source bytes read once + final output
                   bytes written once) / the device time of the kernels that move them, against the measured copy peak
  cpu_baseline     the oracle port of the same nodes on ONE host thread (the reference's Runner drives a graph from one
                   thread, src/infra/runner.cpp:151), on a bounded sample, 1152-sample frames
  parity           GPU result vs the oracle on that sample (bit exact; spectrum 1e-5 of the frame peak)

Only bench.py imports this module; the oracle is used as the checker and as the timed CPU baseline, never on the GPU path.
"""
import ctypes as C
import os
import time

import numpy as np


def _timed(torch, fn, reps):
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def _bits_equal(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


class _Graph:
    """one project bound to its sources; result(eng) -> [(device pointer, bytes), ...] of what a user reads back"""

    def __init__(self, mods, project, sources, result, reps):
        self.torch, self.nodey, self.engine = mods
        self.eng = self.engine.Engine(project.json())
        self.sources, self.result, self.reps = sources, result, reps

    def bind(self, sources):
        self.eng._keep = []
        for i, (x, fmt, rate) in enumerate(sources):
            self.eng.bind_source(i, x, fmt, rate)

    def measure(self):
        torch, nodey = self.torch, self.nodey
        L = nodey.lib()
        self.bind(self.sources)
        for _ in range(2):
            self.eng.run()
        ms = _timed(torch, self.eng.run, self.reps)
        # per-kernel device time: one launch per WSOLA chain, so that events around a launch time that kernel alone
        saved = os.environ.get("NODEY_ST_CHUNKS")
        os.environ["NODEY_ST_CHUNKS"] = "1"
        self.eng.run()
        nodey.profile_enable(True)
        self.eng.run()
        kernels = nodey.profile_report()
        nodey.profile_enable(False)
        if saved is None:
            del os.environ["NODEY_ST_CHUNKS"]
        else:
            os.environ["NODEY_ST_CHUNKS"] = saved
        # end to end: pinned host sources, result copied back
        host = [(x.cpu().pin_memory(), fmt, rate) for x, fmt, rate in self.sources]
        h2d = sum(x.numel() * x.element_size() for x, _, _ in host)
        self.bind(host)
        self.eng.run()
        sizes = [nb for _, nb in self.result(self.eng)]
        hout = [torch.empty((nb,), dtype=torch.uint8, pin_memory=True) for nb in sizes]

        def step():
            self.eng.run()
            for (ptr, nb), h in zip(self.result(self.eng), hout):
                nodey.check(L.nodey_memcpy_d2h(C.c_void_p(h.data_ptr()), C.c_void_p(ptr), nb, None))
            torch.cuda.synchronize()

        step()
        ms_e2e = _timed(torch, step, self.reps)
        self.bind(self.sources)
        return ms, ms_e2e, h2d, sum(sizes), kernels

    def close(self):
        self.eng.close()


def _audio_result(eng):
    out = eng.output()
    plane = out.frames * 4 * (1 if out.fmt >= 5 else out.ch)
    return [(out.p0, plane)] + ([(out.p1, plane)] if out.fmt >= 5 and out.ch == 2 else [])


def _roofline(kernels, algo_bytes, peak, movers):
    """dominant kernel by device time; achieved = algorithmic bytes of the config / device time of the kernels that
    move them (`movers`: the kernels the config's bytes pass through, in the order they run)"""
    if not kernels:
        return None
    name, st = max(kernels.items(), key=lambda kv: kv[1]["ms"])
    ms_movers = sum(v["ms"] for k, v in kernels.items() if k in movers) or st["ms"]
    gbs = algo_bytes / (ms_movers * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": name, "launch_ms": st["ms"] / max(st["launches"], 1), "achieved": gbs, "peak": peak, "unit": "GB/s",
            "frac": gbs / peak, "algorithmic_bytes": algo_bytes,
            "kernels_ms": {k: round(v["ms"], 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}}


def run_configs(torch, nodey, engine, peak, reps=3):
    """-> {"configs[0]": {...}, ..., "configs[3]": {...}}"""
    from oracle import oracle as O
    from oracle import graph_oracle as G
    O.build()
    mods = (torch, nodey, engine)
    L = nodey.lib()
    out = {}

    def synth(n, rate, track, s16=False):
        x = torch.empty((n, 2), dtype=torch.float32, device="cuda")
        nodey.check(L.nodey_synth(nodey._dp(x), None, n, 2, rate, track, 0, None))
        if s16:
            return torch.from_numpy(O.f32_to_s16(x.cpu().numpy())).cuda()
        return x

    # ---- configs[0]: 60 s stereo 44.1 kHz S16 -> gain 0.8 / gain 0.5 -> audio_amix(2) -> output (reference nodes only) ----
    n0 = 44100 * 60
    p = engine.Project()
    src = p.add("audio_input", {"file_path": [""]})
    g0 = p.add("audio_volume_adjust", {"volume": 0.8}); g1 = p.add("audio_volume_adjust", {"volume": 0.5})
    mix = p.add("audio_amix", engine.amix_info([0.5, 0.5]))
    sink = p.add("audio_output")
    p.link(src, "output_0", g0, "input"); p.link(src, "output_0", g1, "input")
    p.link(g0, "output", mix, "input_1"); p.link(g1, "output", mix, "input_2"); p.link(mix, "output", sink, "input")
    x0 = synth(n0, 44100, 0, s16=True)
    g = _Graph(mods, p, [(x0, nodey.FMT_S16, 44100)], _audio_result, reps)
    ms, ms_e2e, h2d, d2h, kernels = g.measure()
    got = g.eng.output().numpy()
    xh = x0.cpu().numpy()
    t0 = time.perf_counter()
    a = O.gain(xh, O.FMT_S16, 0.8); b = O.gain(xh, O.FMT_S16, 0.5)
    rl, rr = O.amix([O.make_track(a, O.FMT_S16, 44100), O.make_track(b, O.FMT_S16, 44100)], [0.5, 0.5])
    cpu_s = time.perf_counter() - t0
    frames_out = got.shape[1]
    out["configs[0]"] = {
        "workload": "60 s stereo 44.1 kHz S16 -> audio_volume_adjust x2 -> audio_amix(2) (swr 44.1 -> 48 kHz) -> audio_output",
        "value": 60.0 / (ms * 1e-3), "unit": "audio-s/s", "ms": ms,
        "e2e": {"value": 60.0 / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms": ms_e2e, "h2d_bytes": h2d, "d2h_bytes": d2h},
        "roofline": _roofline(kernels, n0 * 4 + frames_out * 8, peak, {"gain_s16_kernel", "resample_tile_kernel", "mix_kernel"}),
        "cpu_baseline": {"value": 60.0 / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port", "sample": "the whole 60 s graph"},
        "parity": {"bit_exact": bool(_bits_equal(got, np.stack([rl, rr]))), "against": "oracle gain + amix, full length"}}
    g.close()
    del x0

    # ---- configs[1]: 10 min stereo 48 kHz float -> pitch_modifier(+3) -> velocity_modifier(1.25, keep pitch) ----
    def st_project():
        q = engine.Project()
        s_ = q.add("audio_input", {"file_path": [""]})
        pm = q.add("pitch_modifier", {"pitch": 3.0})
        vm = q.add("velocity_modifier", {"velocity": 1.25, "keep_pitch": True})
        o_ = q.add("audio_output")
        q.link(s_, "output_0", pm, "input"); q.link(pm, "output", vm, "input"); q.link(vm, "output", o_, "input")
        return q
    n1 = 48000 * 600
    x1 = synth(n1, 48000, 1)
    g = _Graph(mods, st_project(), [(x1, nodey.FMT_FLT, 48000)], _audio_result, reps)
    ms, ms_e2e, h2d, d2h, kernels = g.measure()
    m2 = g.eng.output().frames
    g.close()
    # parity + CPU on the first 60 s (the oracle runs at about 100 audio-s/s on one core)
    ns = 48000 * 60
    gs = _Graph(mods, st_project(), [(x1[:ns].contiguous(), nodey.FMT_FLT, 48000)], _audio_result, 1)
    gs.bind(gs.sources); gs.eng.run()
    got = gs.eng.output().numpy()
    xh = x1[:ns].cpu().numpy()
    t0 = time.perf_counter()
    y1, _, _ = O.soundtouch(xh, 48000, 1.0, O.pitch_node_factor(3.0), 1152, want_offsets=False)
    y2, _, _ = O.soundtouch(y1, 48000, 1.25, O.velocity_node_pitch(1.25, True), 1152, want_offsets=False)
    cpu_s = time.perf_counter() - t0
    gs.close()
    m1 = n1                       # the pitch node keeps the duration
    out["configs[1]"] = {
        "workload": "10 min stereo 48 kHz float -> pitch_modifier(+3 semitones) -> velocity_modifier(1.25, keep_pitch) -> audio_output (one track: "
                    "the WSOLA chains are sequential, 4-CTA clusters; chunked launches overlap the two nodes)",
        "value": 600.0 / (ms * 1e-3), "unit": "audio-s/s", "ms": ms,
        "e2e": {"value": 600.0 / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms": ms_e2e, "h2d_bytes": h2d, "d2h_bytes": d2h},
        "roofline": dict(_roofline(kernels, (n1 + m1) * 8 + (m1 + m2) * 8, peak, {"tds_offsets_kernel", "st_post_kernel"}),
                         note="one stream: bounded by the latency of the sequential WSOLA chain (FP32 / latency), not by HBM (SURVEY.md 8d)"),
        "cpu_baseline": {"value": 60.0 / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port", "sample": "first 60 s of the track through both nodes"},
        "parity": {"bit_exact": bool(_bits_equal(got, y2)), "against": "oracle SoundTouch model (parity unpinned: no SoundTouch library in the image), first 60 s"}}
    del x1

    # ---- configs[2]: 16 x 5 min stereo 44.1 kHz float -> audio_amix(16, 1/16) ----
    def mix_project():
        q = engine.Project()
        s_ = q.add("audio_input", {"file_path": [""] * 16})
        m_ = q.add("audio_amix", engine.amix_info([1.0 / 16] * 16))
        o_ = q.add("audio_output")
        for i in range(16):
            q.link(s_, f"output_{i}", m_, f"input_{i + 1}")
        q.link(m_, "output", o_, "input")
        return q
    n2 = 44100 * 300
    xs = [synth(n2, 44100, t) for t in range(16)]
    g = _Graph(mods, mix_project(), [(x, nodey.FMT_FLT, 44100) for x in xs], _audio_result, reps)
    ms, ms_e2e, h2d, d2h, kernels = g.measure()
    frames_out = g.eng.output().frames
    g.close()
    ns = 44100 * 20
    gs = _Graph(mods, mix_project(), [(x[:ns].contiguous(), nodey.FMT_FLT, 44100) for x in xs], _audio_result, 1)
    gs.bind(gs.sources); gs.eng.run()
    got = gs.eng.output().numpy()
    hs = [x[:ns].cpu().numpy() for x in xs]
    t0 = time.perf_counter()
    rl, rr = O.amix([O.make_track(h, O.FMT_FLT, 44100) for h in hs], [1.0 / 16] * 16)
    cpu_s = time.perf_counter() - t0
    gs.close()
    out["configs[2]"] = {
        "workload": "16 x 5 min stereo 44.1 kHz float -> audio_amix(16, volumes 1/16): swr 44.1 -> 48 kHz polyphase FIR fused with the mix",
        "value": 300.0 / (ms * 1e-3), "unit": "audio-s/s", "ms": ms,
        "e2e": {"value": 300.0 / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms": ms_e2e, "h2d_bytes": h2d, "d2h_bytes": d2h},
        "roofline": dict(_roofline(kernels, 16 * n2 * 8 + frames_out * 8, peak, {"resample_tile_kernel"}),
                         note="co-limited by FP32: 64 fused multiply-adds per output sample and input (SURVEY.md 8d)"),
        "cpu_baseline": {"value": 20.0 / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port", "sample": "first 20 s of the 16 tracks"},
        "parity": {"bit_exact": bool(_bits_equal(got, np.stack([rl, rr]))), "against": "oracle amix node (libswresample model pinned to the real library), first 20 s"}}
    del xs

    # ---- configs[3]: 1 h stereo 48 kHz float -> audio_spectrum(4096, 1024, hann) ----
    n3 = 48000 * 3600
    q = engine.Project()
    s_ = q.add("audio_input", {"file_path": [""]})
    sp = q.add("audio_spectrum", {"fft_size": 4096, "hop": 1024, "window": "hann"})
    q.link(s_, "output_0", sp, "input")
    x3 = synth(n3, 48000, 2)

    def spec_result(eng):
        pr = eng.product(sp, "output")
        return [(pr.p0, pr.ch * pr.frames * pr.bins * 8)]
    g = _Graph(mods, q, [(x3, nodey.FMT_FLT, 48000)], spec_result, reps)
    ms, ms_e2e, h2d, d2h, kernels = g.measure()
    pr = g.eng.product(sp, "output")
    frames = pr.frames
    # parity: spot frames against the oracle's double DFT; CPU: the oracle STFT of the first 30 s (both channels)
    ok = True
    spec_bytes = pr.ch * pr.frames * pr.bins * 8
    for f in (0, 1, frames // 3, frames - 1):
        seg = x3[f * 1024:f * 1024 + 4096].cpu().numpy()
        for c in range(2):
            ref = O.stft(seg[:, c].copy())[0]
            gotf = np.empty(2049, np.complex64)
            nodey.check(L.nodey_memcpy_d2h(gotf.ctypes.data_as(C.c_void_p), C.c_void_p(pr.p0 + ((c * frames + f) * 2049) * 8), 2049 * 8, None))
            nodey.check(L.nodey_device_synchronize())
            ok = ok and bool(np.abs(gotf - ref).max() <= 1e-5 * np.abs(ref).max())
    xh = x3[:48000 * 30].cpu().numpy()
    t0 = time.perf_counter()
    for c in range(2):
        O.stft(xh[:, c].copy())
    cpu_s = time.perf_counter() - t0
    g.close()
    out["configs[3]"] = {
        "workload": "1 h stereo 48 kHz float -> audio_spectrum (4096-point periodic Hann STFT, hop 1024, 2 x 168747 x 2049 complex64)",
        "value": 3600.0 / (ms * 1e-3), "unit": "audio-s/s", "ms": ms,
        "e2e": {"value": 3600.0 / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms": ms_e2e, "h2d_bytes": h2d, "d2h_bytes": d2h},
        "roofline": _roofline(kernels, n3 * 8 + spec_bytes, peak, {"stft4096_kernel"}),
        "cpu_baseline": {"value": 30.0 / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port",
                         "sample": "first 30 s, both channels (double-precision radix-2 FFT per frame: the oracle is a checker, not a tuned FFT)"},
        "parity": {"within_1e-5_of_frame_peak": ok, "against": "oracle double DFT, frames 0, 1, M/3, M-1 of both channels"}}
    del x3
    torch.cuda.empty_cache()
    nodey.check(L.nodey_trim_memory())
    return out


def run_segments(torch, nodey, dist, world, rank, dev):
    """Time-segment sharding of ONE long stream over the ranks (bindings/segments.py): every rank computes a
    contiguous range of the output from its slice of the input; no collective.  Two cases: configs[3] (1 h STFT) and one
    configs[2] track (5 min, 44.1 -> 48 kHz).  Device time of the rank's own range, max over ranks, next to the time of the
    whole stream on one GPU; each rank checks its range bit for bit against the whole-stream result it computes locally."""
    import segments as SG
    L = nodey.lib()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag):
        t = torch.tensor([1 if flag else 0], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    res = {}
    # ---- STFT, 1 h stereo 48 kHz (planar) ----
    n = 48000 * 3600
    x = torch.empty((n, 2), dtype=torch.float32, device=dev)
    nodey.check(L.nodey_synth(nodey._dp(x), None, n, 2, 48000, 2, 0, None))
    xp = x.T.contiguous()
    del x
    seg = SG.stft_segment(n, world, rank)
    xs = xp[:, seg["in0"]:seg["in1"]].contiguous()
    mine = nodey.stft(xs, False)
    buf = torch.empty_like(mine)
    ms_part = _timed(torch, lambda: nodey.stft(xs, False, out=buf), 3)
    whole = nodey.stft(xp, False)
    wbuf = torch.empty_like(whole)
    ms_whole = _timed(torch, lambda: nodey.stft(xp, False, out=wbuf), 3)
    same = torch.equal(mine.view(torch.float32), whole[:, seg["m0"]:seg["m1"]].contiguous().view(torch.float32))
    del mine, buf, whole, wbuf, xs, xp
    torch.cuda.empty_cache()
    res["stft_1h"] = {"ms_one_gpu": ms_whole, "ms_sharded": max_over_ranks(ms_part), "bit_identical": all_true(same),
                      "halo_frames": 3072, "what": "configs[3]: spectrum frames split evenly, each rank reads its input range + 3072 frames"}
    # ---- resampler, one 5 min stereo 44.1 kHz track ----
    n = 44100 * 300
    x = torch.empty((n, 2), dtype=torch.float32, device=dev)
    nodey.check(L.nodey_synth(nodey._dp(x), None, n, 2, 44100, 3, 0, None))
    rs = nodey.Resampler(44100, 48000)
    total = rs.out_count(n, True)
    sg = SG.resample_ranges(rs.info(), n, total, world, rank)
    xs = x[sg["in0"]:sg["in1"]].contiguous()
    want = sg["skip"] + sg["k1"] - sg["k0"]
    part = rs.run(xs, nodey.FMT_FLT, flush=sg["flush"], out_frames=want)
    pbuf = torch.empty_like(part)
    ms_part = _timed(torch, lambda: rs.run(xs, nodey.FMT_FLT, flush=sg["flush"], out_frames=want, out=pbuf), 5)
    whole = rs.run(x, nodey.FMT_FLT)
    wbuf = torch.empty_like(whole)
    ms_whole = _timed(torch, lambda: rs.run(x, nodey.FMT_FLT, out=wbuf), 5)
    same = torch.equal(part[:, sg["skip"]:], whole[:, sg["k0"]:sg["k1"]])
    res["resample_5min"] = {"ms_one_gpu": ms_whole, "ms_sharded": max_over_ranks(ms_part), "bit_identical": all_true(same),
                            "what": "one configs[2] track: output periods split evenly, each rank starts whole periods (>= 15 input frames) early"}
    del x, part, pbuf, whole, wbuf
    torch.cuda.empty_cache()
    for v in res.values():
        v["speedup"] = v["ms_one_gpu"] / v["ms_sharded"] if v["ms_sharded"] > 0 else None
        v["n_gpus"] = world
    return res
