"""Development: diagnostics of one end-to-end run (pinned host sources) of the config-5 project."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey, engine, pipeline
T, secs = int(os.environ.get("T", "256")), 180
n = 44100 * secs
x = torch.empty((T, n, 2), dtype=torch.float32, pin_memory=True)
x.view(-1)[::4096] = 0.1
p, ids = engine.config5_project(T, [pipeline.track_gain(t) for t in range(T)])
e = engine.Engine(p.json())
for t in range(T):
    e.bind_source(t, x[t], 3, 44100)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); e.run(); torch.cuda.synchronize()
    print(f"run {it}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
print(e.diagnostics())
