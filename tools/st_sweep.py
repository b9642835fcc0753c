"""Development: SoundTouch batch time vs batch size and cluster size (pitch +3 then tempo 1.25, 3 min tracks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey as nd
n = 48000 * 180
x = torch.empty((256, n, 2), dtype=torch.float32, device="cuda").normal_(0, 0.1)
sp = nd.SoundTouch.pitch_node(48000, 2, 3.0)
m1, _ = sp.out_frames(n)
y = torch.empty((256, m1, 2), dtype=torch.float32, device="cuda")
def t(fn, it=2):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it
for N in (32, 48, 64, 74, 96, 128, 148, 192, 256):
    row = []
    for cl in (1, 2, 4):
        sp.set_cluster(cl)
        row.append(t(lambda: sp.run(x[:N], out=y[:N])))
    print(f"N={N:4d}  pitch node ms: CL1 {row[0]:7.2f}  CL2 {row[1]:7.2f}  CL4 {row[2]:7.2f}   per-track best {min(row)/N:.3f} ms", flush=True)
