"""Development: pinned host -> device rate for the bench's upload pattern (256 x 63.5 MB on one stream)."""
import time, torch
T, n = 256, 44100 * 180
x = torch.empty((T, n, 2), dtype=torch.float32, pin_memory=True); x.normal_(0, 0.1)
d = torch.empty((T, n, 2), dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
for chunk in (1, 4, 16):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for t in range(0, T, chunk): d[t:t + chunk].copy_(x[t:t + chunk], non_blocking=True)
        s.synchronize(); dt = time.perf_counter() - t0
    print(f"{T // chunk} copies of {chunk * n * 8 / 1e6:.1f} MB: {dt * 1e3:.1f} ms, {x.numel() * 4 / dt / 1e9:.1f} GB/s", flush=True)
# while the SMs and HBM are busy
a = torch.empty(1 << 30, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s):
    for t in range(T): d[t].copy_(x[t], non_blocking=True)
for _ in range(60): b.copy_(a)
s.synchronize(); dt = time.perf_counter() - t0
print(f"with concurrent device copies: {dt * 1e3:.1f} ms, {x.numel() * 4 / dt / 1e9:.1f} GB/s")
