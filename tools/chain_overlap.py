"""Development: do the pitch node's and the tempo node's WSOLA chains overlap when they run as chunk launches on two
streams?  C ABI only (no engine): serial whole-track launches vs. chunked launches with events, T tracks x S seconds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
import ctypes as C
import torch
import nodey as nd

T = int(os.environ.get("T", "32")); S = int(os.environ.get("S", "180")); NC = int(os.environ.get("NC", "8"))
n = 48000 * S
L = nd.lib()
x = torch.empty((T, n, 2), device="cuda").uniform_(-0.5, 0.5)
st1 = nd.SoundTouch.pitch_node(48000, 2, 3.0)
st2 = nd.SoundTouch.velocity_node(48000, 2, 1.25, True)
m1, q1 = st1.out_frames(n); m2, q2 = st2.out_frames(m1)
y1 = torch.empty((T, m1, 2), device="cuda"); y2 = torch.empty((T, m2, 2), device="cuda")
o1 = torch.zeros((T, q1), dtype=torch.int32, device="cuda"); o2 = torch.zeros((T, q2), dtype=torch.int32, device="cuda")
p1 = st1.chunks(n, 1152, NC); p2 = st2.chunks(m1, 1152, NC)
print("pitch chunks", p1[:3], "...", "tempo chunks", p2[:3], flush=True)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()

def chunk(st, out, xin, nin, m, offs, c, nc, stream):
    nd.check(L.nodey_soundtouch_run_chunk(st.h, nd._dp(out), out.stride(0), nd._dp(xin), xin.stride(0), T, nin, 1152, m,
                                          nd._dp(offs), offs.stride(0), c, nc, 0, C.c_void_p(stream.cuda_stream)))

def serial():
    for c in range(len(p1)): chunk(st1, y1, x, n, m1, o1, c, len(p1), sA)
    for c in range(len(p2)): chunk(st2, y2, y1, m1, m2, o2, c, len(p2), sA)

def pipelined():
    evs = []
    for c in range(len(p1)):
        chunk(st1, y1, x, n, m1, o1, c, len(p1), sA)
        e = torch.cuda.Event(); e.record(sA); evs.append(e)
    waited = 0
    sB.wait_stream(torch.cuda.current_stream())
    for d in range(len(p2)):
        while waited < len(p1) and (waited == 0 or p1[waited - 1][1] < p2[d][0]):
            sB.wait_event(evs[waited]); waited += 1
        chunk(st2, y2, y1, m1, m2, o2, d, len(p2), sB)
    sA.wait_stream(sB)

def interleaved():
    # same, but the host enqueues pitch chunk c and then whatever tempo chunks it has unblocked
    evs, waited, d = [], 0, 0
    for c in range(len(p1)):
        chunk(st1, y1, x, n, m1, o1, c, len(p1), sA)
        e = torch.cuda.Event(); e.record(sA); evs.append(e)
        while d < len(p2) and p2[d][0] <= p1[c][1]:
            while waited <= c: sB.wait_event(evs[waited]); waited += 1
            chunk(st2, y2, y1, m1, m2, o2, d, len(p2), sB); d += 1
    while d < len(p2):
        chunk(st2, y2, y1, m1, m2, o2, d, len(p2), sB); d += 1
    sA.wait_stream(sB)

def timed(label, fn, reps=4):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(sA); fn(); b.record(sA); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"T={T} NC={NC} {label}: " + " ".join(f"{t:.2f}" for t in ts) + " ms", flush=True)
    return y2.clone(), o2.clone()

ra = timed("serial   ", serial)
rb = timed("pipelined", pipelined)
rc = timed("interleav", interleaved)
print("bit identical:", torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1]) and torch.equal(ra[0], rc[0]))
