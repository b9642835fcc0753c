#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_soundtouch.py tests/test_gpu_guards.py -m gpu -q -x 2>&1 | tail -3
SECS=60 timeout 600 python - <<'PY'
import os, sys
sys.path.insert(0, "nodey-audio-editor_b200/bindings")
import torch
import nodey as nd
n = 48000 * 60
x = torch.empty((256, n, 2), dtype=torch.float32, device="cuda").normal_(0, 0.1)
xp = x.permute(0, 2, 1).contiguous()
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it
for name, sp in (("pitch +3", nd.SoundTouch.pitch_node(48000, 2, 3.0)), ("tempo 1.25", nd.SoundTouch.velocity_node(48000, 2, 1.25, True))):
    m1, _ = sp.out_frames(n)
    y = torch.empty((256, m1, 2), dtype=torch.float32, device="cuda")
    nd.profile_enable(True)
    sp.run(x, out=y); torch.cuda.synchronize()
    nd.profile_enable(True)
    for _ in range(3): sp.run(x, out=y)
    torch.cuda.synchronize()
    print(name, "256 x 60 s interleaved input (3 runs):", nd.profile_report())
    nd.profile_enable(False)
    print(name, "whole node ms:", t(lambda: sp.run(x, out=y)))
PY
T=256 timeout 300 python tools/chain_trace.py 2>&1 | head -4
