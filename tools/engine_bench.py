"""Development timing of the C++ engine on the config-5 project (device-resident sources)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey, engine, pipeline

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 180
n = 44100 * secs
x = torch.empty((T, n, 2), dtype=torch.float32, device="cuda")
for t in range(T):
    nodey.check(nodey.lib().nodey_synth(nodey._dp(x[t]), None, n, 2, 44100, t, 0, None))
torch.cuda.synchronize()
p, ids = engine.config5_project(T, [pipeline.track_gain(t) for t in range(T)])
e = engine.Engine(p.json())
for t in range(T):
    e.bind_source(t, x[t], 3, 44100)
for it in range(4):
    l0 = nodey.profile_launches()
    t0 = time.perf_counter(); e.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"run {it}: {dt*1e3:.1f} ms  -> {T*secs/dt:.0f} audio-s/s, launches {nodey.profile_launches()-l0}", flush=True)
nodey.profile_enable(True); e.run(); rep = nodey.profile_report(); nodey.profile_enable(False)
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:28s} {v['launches']:5d} {v['ms']:9.2f} ms")
print("sum ms", sum(v["ms"] for v in rep.values()))
