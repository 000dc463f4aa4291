"""Host-to-device / device-to-host copy bandwidth of this box from pinned memory (context for the e2e number)."""
import torch, time
n = 1 << 27   # 1 GiB of float64
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    for size in (n, n // 16):
        hs, ds = h[:size], d[:size]
        f = (lambda: ds.copy_(hs, non_blocking=True)) if name == "H2D" else (lambda: hs.copy_(ds, non_blocking=True))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if size == n else 40
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} {size * 8 / 2**20:7.0f} MiB per copy: {reps * size * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s")
