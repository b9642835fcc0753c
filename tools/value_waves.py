"""Development: device-resident render time of the config-5 project for several wave patterns / lane counts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey, engine, pipeline
T, secs = int(os.environ.get("T", "256")), 180
n = 44100 * secs
x = torch.empty((T, n, 2), dtype=torch.float32, device="cuda").normal_(0, 0.1)
p, ids = engine.config5_project(T, [pipeline.track_gain(t) for t in range(T)])
e = engine.Engine(p.json())
for t in range(T):
    e.bind_source(t, x[t], 3, 44100)
def timed(label, reps=4):
    ts = []
    for it in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); e.run(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label}: " + " ".join(f"{t:.1f}" for t in ts) + " ms", flush=True)
timed("single wave")
for pat, lanes in [(p, l) for p in sys.argv[1:] for l in ("2", "3")]:
    os.environ["NODEY_WAVES"] = pat; os.environ["NODEY_COMPUTE_LANES"] = lanes
    timed(f"waves {pat} lanes {lanes}")
