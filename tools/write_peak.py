"""Development: what a write-dominated stream can reach on this GPU (the measured copy peak counts read + write bytes)."""
import torch
def t(fn, it=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
ms = t(lambda: b.copy_(a)); print(f"copy 2 GiB -> 2 GiB: {ms:.3f} ms  {4 * n / ms / 1e6:.0f} GB/s (read + write)")
ms = t(lambda: b.zero_()); print(f"memset 2 GiB: {ms:.3f} ms  {2 * n / ms / 1e6:.0f} GB/s (write only)")
ms = t(lambda: b.fill_(1.5)); print(f"fill 2 GiB: {ms:.3f} ms  {2 * n / ms / 1e6:.0f} GB/s (write only)")
ms = t(lambda: a.sum()); print(f"sum 2 GiB: {ms:.3f} ms  {2 * n / ms / 1e6:.0f} GB/s (read only)")
c = torch.empty(n // 4, dtype=torch.bfloat16, device="cuda").normal_()
d = b.view(4, n // 4)
ms = t(lambda: torch.add(c, 1.0, out=d[0])); 
ms = t(lambda: d.copy_(c.expand(4, n // 4))); print(f"broadcast copy 0.5 GiB -> 2 GiB: {ms:.3f} ms  {2.5 * n / ms / 1e6:.0f} GB/s (1 read : 4 write)")
