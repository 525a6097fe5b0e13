#!/usr/bin/env python
"""Host->device copy bandwidth from pinned memory for one bench step's input (256 x 3 x 224 x 224 bf16 = 77 MB): one copy on one
stream vs the same bytes split over 2 / 4 / 8 streams.  Decides how approx/pipeline.py issues its copies."""
import sys
import torch

n = 256 * 3 * 224 * 224
host = torch.empty(n, dtype=torch.bfloat16).pin_memory()
dev = torch.empty(n, dtype=torch.bfloat16, device="cuda")
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunk = n // parts
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dev[i * chunk:(i + 1) * chunk].copy_(host[i * chunk:(i + 1) * chunk], non_blocking=True)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    iters = 20
    for _ in range(iters):
        go()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    print(f"H2D {n * 2 / 1e6:.1f} MB in {parts} part(s): {ms:.3f} ms = {n * 2 / ms / 1e6:.1f} GB/s", flush=True)
d2h = torch.empty(256 * 512 * 7 * 7, dtype=torch.bfloat16).pin_memory()
src = torch.empty_like(d2h, device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(20):
    d2h.copy_(src, non_blocking=True)
b.record(); torch.cuda.synchronize()
print(f"D2H {d2h.numel() * 2 / 1e6:.1f} MB: {a.elapsed_time(b) / 20:.3f} ms = {d2h.numel() * 2 / (a.elapsed_time(b) / 20) / 1e6:.1f} GB/s")
