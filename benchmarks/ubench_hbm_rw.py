#!/usr/bin/env python
"""Write-only / read-only / copy HBM bandwidth on this GPU through library kernels (cudaMemset-style fill, device copy, reduction),
buffers of 1.6 and 3.3 GB (far above the 126 MB L2).  Context for the fused conv's first layer, which writes 4x what it reads:
is the write direction alone able to reach the copy figure MEASURED_PEAKS.json quotes (reads + writes summed)?

    python benchmarks/ubench_hbm_rw.py
"""
import json
import torch


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    for gb in (1.644, 3.288):
        n = int(gb * 1e9 / 2)
        x = torch.empty(n, dtype=torch.bfloat16, device="cuda")
        y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
        x.normal_()
        ms_fill = t(lambda: y.fill_(1.0))
        ms_zero = t(lambda: y.zero_())
        ms_copy = t(lambda: y.copy_(x))
        xi = x.view(torch.int32)
        ms_read = t(lambda: xi.sum())
        print(json.dumps({"bytes": n * 2, "fill_ms": round(ms_fill, 4), "fill_write_GBps": round(n * 2 / ms_fill / 1e6, 1),
                          "memset_ms": round(ms_zero, 4), "memset_write_GBps": round(n * 2 / ms_zero / 1e6, 1),
                          "copy_ms": round(ms_copy, 4), "copy_rw_GBps": round(2 * n * 2 / ms_copy / 1e6, 1),
                          "reduce_ms": round(ms_read, 4), "reduce_read_GBps": round(n * 2 / ms_read / 1e6, 1)}))


if __name__ == "__main__":
    main()
