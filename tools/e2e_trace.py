"""Development: end-to-end engine time on the config-5 project with pinned host sources, for several
wave patterns / compute-lane counts (NODEY_WAVES, NODEY_WAVE, NODEY_COMPUTE_LANES)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey, engine, pipeline
T, secs = 256, 180
n = 44100 * secs
x = torch.empty((T, n, 2), dtype=torch.float32, pin_memory=True)
x.view(-1)[::4096] = 0.1      # touch every page (the values do not matter for the timing)
p, ids = engine.config5_project(T, [pipeline.track_gain(t) for t in range(T)])
e = engine.Engine(p.json())
for t in range(T):
    e.bind_source(t, x[t], 3, 44100)
def timed(label, reps=3):
    ts = []
    for it in range(reps):
        t0 = time.perf_counter(); e.run(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label}: " + " ".join(f"{t:.1f}" for t in ts) + " ms", flush=True)
timed("default (warm-up)")
timed("default")
patterns = sys.argv[1:] or ["32,64,64,64,32", "32", "16,32,48,64,48,32,16", "32,48,48,48,48,32", "48,64,64,48,32", "32,64,64,48,32,16", "24,40,64,64,40,24"]
for lanes in os.environ.get("LANES", "2,3").split(","):
    os.environ["NODEY_COMPUTE_LANES"] = lanes
    for pat in patterns:
        os.environ["NODEY_WAVES"] = pat
        timed(f"waves {pat}, lanes={lanes}")
