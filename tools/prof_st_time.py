"""Development: device time of the SoundTouch pitch node's two kernels (N tracks x SECS s), event timed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey as nd
N = int(os.environ.get("N", "256")); secs = int(os.environ.get("SECS", "60"))
n = 48000 * secs
x = torch.empty((N, n, 2), dtype=torch.float32, device="cuda").normal_(0, 0.1)
sp = nd.SoundTouch.pitch_node(48000, 2, 3.0)
m1, _ = sp.out_frames(n)
y = torch.empty((N, m1, 2), dtype=torch.float32, device="cuda")
sp.run(x, out=y)
torch.cuda.synchronize()
nd.profile_enable(True)
for _ in range(2):
    sp.run(x, out=y)
torch.cuda.synchronize()
rep = nd.profile_report()
nd.profile_enable(False)
print({k: round(v["ms"] / v["launches"], 3) for k, v in rep.items()}, "stagger", os.environ.get("NODEY_ST_STAGGER"))
