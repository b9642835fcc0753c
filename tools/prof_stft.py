"""Development: the STFT alone (10 min stereo, planar and interleaved) for ncu captures and timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey as nd
secs = int(os.environ.get("SECS", "600"))
x = nd.synth(48000 * secs, 2, 48000)
xp = x.T.contiguous()
so = torch.empty((2, nd.stft_frames(48000 * secs), 2049), dtype=torch.complex64, device="cuda")
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it
byt = 48000 * secs * 8 + so.numel() * 8
for name, fn in (("planar", lambda: nd.stft(xp, False, out=so)), ("interleaved", lambda: nd.stft(x, True, out=so))):
    ms = t(fn)
    print(f"stft {secs} s stereo {name}: {ms:.3f} ms  {byt / ms / 1e6:.0f} GB/s", flush=True)
import numpy as np
xs = xp[:, :4096 * 8].cpu().numpy().astype(np.float64)
w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(4096) / 4096)
got = nd.stft(xp[:, :4096 * 8].contiguous(), False).cpu().numpy()
worst = 0.0
for c in range(2):
    for m in range(got.shape[1]):
        ref = np.fft.rfft(xs[c, m * 1024:m * 1024 + 4096] * w)
        worst = max(worst, float(np.abs(got[c, m] - ref).max() / np.abs(ref).max()))
print(f"max error relative to the frame peak: {worst:.3e} (bar 1e-5)")
