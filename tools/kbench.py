"""Per-kernel micro-benchmark (device-resident, CUDA events): achieved algorithmic GB/s of each
node kernel against the measured HBM copy peak.  Development tool; bench.py is the contract."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import nodey as nd  # noqa: E402

PEAK = 6546.9
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, bytes_, t, extra=""):
    gbs = bytes_ / t / 1e9
    print(f"{name:34s} {t*1e3:9.3f} ms  {gbs:8.1f} GB/s  {gbs/PEAK*100:5.1f}% of measured peak {extra}", flush=True)


def main():
    print(nd.device_info())
    n = 48000 * 60 * 20     # 20 min stereo: 460 MB float
    x = nd.synth(n, 2, 48000)
    out = torch.empty_like(x)
    report("gain f32", x.numel() * 8, timeit(lambda: nd.gain(x, nd.FMT_FLT, 0.8, out=out)))
    _, s = nd.synth(n, 2, 48000, want_s16=True)
    so = torch.empty_like(s)
    report("gain s16", s.numel() * 4, timeit(lambda: nd.gain(s, nd.FMT_S16, 0.8, out=so)))
    report("split f32", x.numel() * 8, timeit(lambda: nd.split(x, nd.FMT_FLT)))
    report("extract s16->f32", s.numel() * 6, timeit(lambda: nd.extract_interleaved(s, nd.FMT_S16)))
    del out, so, s
    ins = [nd.synth(48000 * 300, 2, 48000, track=i).T.contiguous() for i in range(16)]
    report("mix 16 x 5min", sum(i.numel() for i in ins) * 4 + ins[0].numel() * 4,
           timeit(lambda: nd.mix(ins, [1 / 16] * 16)))
    del ins
    # config 3: 16 x 5 min 44.1k -> 48k + mix (fused) and single-track resample
    n441 = 44100 * 300
    xs = [nd.synth(n441, 2, 44100, track=i) for i in range(16)]
    r = nd.Resampler(44100, 48000)
    m = r.out_count(n441, True)
    o = torch.empty((2, m), dtype=torch.float32, device="cuda")
    report("resample 5min tile", n441 * 8 + m * 8, timeit(lambda: r.run(xs[0], nd.FMT_FLT, out=o)))
    report("resample 5min generic", n441 * 8 + m * 8, timeit(lambda: r.run(xs[0], nd.FMT_FLT, mode=1, out=o)))
    t = timeit(lambda: r.resample_mix(xs, [nd.FMT_FLT] * 16, [1 / 16] * 16, out=o))
    report("config3 resample+mix16 fused", 16 * n441 * 8 + m * 8, t, f"flops {16*m*2*64/t/1e12:.1f} TF/s")
    del xs, o
    # config 4: 1 h stereo 48 k STFT
    n4 = 48000 * 3600
    x4 = nd.synth(n4, 2, 48000)
    mfr = nd.stft_frames(n4)
    so = torch.empty((2, mfr, 2049), dtype=torch.complex64, device="cuda")
    t = timeit(lambda: nd.stft(x4, True, out=so), iters=5)
    report("config4 stft 1h stereo (interl.)", n4 * 8 + so.numel() * 8, t)
    x4p = x4.T.contiguous()
    del x4
    t = timeit(lambda: nd.stft(x4p, False, out=so), iters=5)
    report("config4 stft 1h stereo (planar)", n4 * 8 + so.numel() * 8, t)


if __name__ == "__main__":
    main()
