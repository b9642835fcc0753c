"""Development: time of the WSOLA offsets kernel alone (profiling events) vs candidates per thread, 256 / 32 tracks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey as nd
secs = int(os.environ.get("SECS", "60"))
n = 48000 * secs
x = torch.empty((256, n, 2), dtype=torch.float32, device="cuda").normal_(0, 0.1)
nodes = {"pitch +3": nd.SoundTouch.pitch_node(48000, 2, 3.0), "tempo 1.25": nd.SoundTouch.velocity_node(48000, 2, 1.25, True)}
def t(fn, it=2):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it
for name, sp in nodes.items():
    m1, _ = sp.out_frames(n)
    y = torch.empty((256, m1, 2), dtype=torch.float32, device="cuda")
    for N, cl in ((256, 1), (128, 1), (128, 2), (32, 4)):
        sp.set_cluster(cl)
        row = []
        for kt in ((8, 11, 12, 13, 14, 15, 16) if cl == 1 else ((8,) if cl == 2 else (4,))):
            sp.set_candidates_per_thread(kt)
            for rt in ((False, True) if kt <= 8 else (False,)):
                if rt: os.environ["NODEY_TDS_RUNTIME_SK"] = "1"
                row.append(f"kt{kt}{'r' if rt else ''} {t(lambda: sp.run(x[:N], out=y[:N])):7.2f}")
                os.environ.pop("NODEY_TDS_RUNTIME_SK", None)
        sp.set_candidates_per_thread(0)
        print(f"{name:10s} N={N:4d} CL{cl} whole node ms ({secs} s tracks): " + "  ".join(row), flush=True)
