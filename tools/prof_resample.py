"""Development: run the fused resample+mix (config 3, shortened) and the STFT for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey as nd
n441 = 44100 * 60
xs = [nd.synth(n441, 2, 44100, track=i) for i in range(16)]
r = nd.Resampler(44100, 48000)
m = r.out_count(n441, True)
o = torch.empty((2, m), dtype=torch.float32, device="cuda")
for _ in range(3):
    r.resample_mix(xs, [nd.FMT_FLT] * 16, [1 / 16] * 16, out=o)
    r.run(xs[0], nd.FMT_FLT, out=o)
x4 = nd.synth(48000 * 600, 2, 48000).T.contiguous()
so = torch.empty((2, nd.stft_frames(48000 * 600), 2049), dtype=torch.complex64, device="cuda")
for _ in range(3):
    nd.stft(x4, False, out=so)
torch.cuda.synchronize()
print("ok")
