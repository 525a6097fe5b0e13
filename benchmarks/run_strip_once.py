#!/usr/bin/env python
"""Launches one depthwise strip / tile kernel a few times (target for `ncu --set full -k regex:dw_`).
    python benchmarks/run_strip_once.py [--kh 1 --kw 7 --C 32 --H 128 --batch 256 --reps 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(kh=1, kw=7, C=32, H=128, batch=256, reps=3).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    a = ap.parse_args()
    from approx import backend as B
    x = torch.randn(a.batch, a.H, a.H, a.C, device="cuda").bfloat16()
    w = torch.randn(a.C, 1, a.kh, a.kw, device="cuda") / (a.kh * a.kw) ** 0.5
    wpk = B.pack_dw(w, torch.bfloat16)
    bias = torch.randn(a.C, device="cuda")
    for _ in range(a.reps):
        B.dwconv(x, wpk, bias, 1, a.kh, a.kw, (1, 1), (a.kh // 2, a.kw // 2))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
