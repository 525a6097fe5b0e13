#!/usr/bin/env python
"""Roofline of the individual MSCA depthwise / strip kernels (reference layers/msca.py:18-19,
layers/depth_seperable_conv.py:12-16, core/msca_rep.py:44-45) at the MSCAN-T shapes of BASELINE config 4.

Per kernel: time, achieved HBM GB/s on the algorithmic bytes (read x once + write y once) as a fraction of the
measured copy peak, and FP32 FMA rate as a fraction of the FFMA peak (148 SMs x 128 lanes x clock) — whichever is
higher is the binding roofline.  One JSON object per line.

    python benchmarks/bench_dwconv.py [--batch 32] [--iters 20] [--dtype bf16|f32] [--generic]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402

SHAPES = [(32, 128), (64, 64), (160, 32), (256, 16)]
KERNELS = [(5, 5), (1, 7), (7, 1), (1, 11), (11, 1), (1, 21), (21, 1), (21, 21)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--generic", action="store_true", help="force the generic one-thread-per-output kernel (CNA_DW_GENERIC=1)")
    a = ap.parse_args()
    if a.generic:
        os.environ["CNA_DW_GENERIC"] = "1"
    from approx import backend as B
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    esz = 2 if dt == torch.bfloat16 else 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    ffma_peak = 148 * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12      # TFMA/s
    for (C, H) in SHAPES:
        x = torch.randn(a.batch, H, H, C, device="cuda").to(dt)
        for (kh, kw) in KERNELS:
            w = torch.randn(C, 1, kh, kw, device="cuda") / (kh * kw) ** 0.5
            wpk = B.pack_dw(w, dt)
            bias = torch.randn(C, device="cuda")
            fn = lambda: B.dwconv(x, wpk, bias, 1, kh, kw, (1, 1), (kh // 2, kw // 2))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            elems = a.batch * H * H * C
            gbs = 2 * elems * esz / ms / 1e6
            tfma = elems * kh * kw / ms / 1e9
            print(json.dumps({"bench": "dwconv", "kernel": f"{kh}x{kw}", "C": C, "H": H, "batch": a.batch, "dtype": a.dtype,
                              "path": "generic" if a.generic else "tiled", "ms": round(ms, 4), "hbm_GBps": round(gbs, 1),
                              "frac_hbm_peak": round(gbs / hbm, 3), "TFMA_per_s": round(tfma, 2),
                              "frac_ffma_peak": round(tfma / ffma_peak, 3)}))


if __name__ == "__main__":
    main()
