"""Development: SoundTouch node time for the parameter regions that take different kernel paths
(rate > 1: TDStretch first, fused tail; rate <= 1: cubic -> FIR -> TDStretch, separate kernels; mono)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey as nd
N, secs = 128, 60
n = 48000 * secs
def t(fn, it=2):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it
for ch in (2, 1):
    x = torch.empty((N, n, ch), dtype=torch.float32, device="cuda").normal_(0, 0.1)
    for label, sp in (("pitch +3", nd.SoundTouch.pitch_node(48000, ch, 3.0)), ("pitch -3", nd.SoundTouch.pitch_node(48000, ch, -3.0)),
                      ("tempo 1.25 keep", nd.SoundTouch.velocity_node(48000, ch, 1.25, True)), ("tempo 0.8 keep", nd.SoundTouch.velocity_node(48000, ch, 0.8, True)),
                      ("velocity 1.25", nd.SoundTouch.velocity_node(48000, ch, 1.25, False)), ("velocity 0.8", nd.SoundTouch.velocity_node(48000, ch, 0.8, False))):
        m, nseq = sp.out_frames(n)
        y = torch.empty((N, m, ch), dtype=torch.float32, device="cuda")
        nd.profile_enable(True)
        sp.run(x, out=y); torch.cuda.synchronize()
        rep = nd.profile_report(); nd.profile_enable(False)
        ms = t(lambda: sp.run(x, out=y))
        print(f"ch={ch} {label:16s} {ms:8.2f} ms  ({nseq} seq)  " + ", ".join(f"{k.replace('_kernel','')} {v['ms']:.1f}" for k, v in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])), flush=True)
