"""Development: where one WSOLA sequence spends its time (clock64 phase timers, instrumented build).
make -C tools/micro libnodey_cuda_timing.so; NODEY_CUDA_LIB=tools/micro/libnodey_cuda_timing.so python tools/tds_phases.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NODEY_CUDA_LIB", os.path.join(ROOT, "tools", "micro", "libnodey_cuda_timing.so"))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import torch
import nodey as nd
secs = int(os.environ.get("SECS", "60"))
n = 48000 * secs
x = torch.empty((256, n, 2), dtype=torch.float32, device="cuda").normal_(0, 0.1)
names = ["gather mid+issue copies", "sync A", "corr lane sums", "sync B", "combine+warp redux", "sync C", "cta reduce+publish", "copies landed+norms(i+1)+cluster wait"]
for label, sp in (("pitch +3", nd.SoundTouch.pitch_node(48000, 2, 3.0)), ("tempo 1.25", nd.SoundTouch.velocity_node(48000, 2, 1.25, True))):
    m1, nseq = sp.out_frames(n)
    y = torch.empty((256, m1, 2), dtype=torch.float32, device="cuda")
    for N, cl in ((256, 1), (128, 2), (128, 1), (64, 4), (32, 4), (32, 2), (8, 4)):
        sp.set_cluster(cl)
        sp.run(x[:N], out=y[:N]); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); sp.run(x[:N], out=y[:N]); b.record(); b.synchronize()
        ph = (C.c_ulonglong * 8)()
        nd.lib().nodey_debug_tds_phases(ph)
        tot = sum(ph)
        per = [p / max(1, nseq - 1) for p in ph]
        print(f"{label} N={N} CL={cl}: node {a.elapsed_time(b):.1f} ms, {nseq} seq, {tot / max(1, nseq - 1):.0f} clk/seq | " +
              ", ".join(f"{nm} {v:.0f}" for nm, v in zip(names, per)), flush=True)
