#!/usr/bin/env python
"""Runs the fused FFN half of an MSCAN block a few times at one stage shape (target for an ncu launch list).
    python benchmarks/run_ffn_once.py [--C 32 --exp 8 --H 128 --batch 32 --reps 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402
from torch import nn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(C=32, exp=8, H=128, batch=32, reps=3).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    a = ap.parse_args()
    from approx.layers import MSCANFFN
    C, Hd = a.C, a.C * a.exp

    class FFN(nn.Module):
        def __init__(s_):
            super().__init__()
            s_.fc1 = nn.Conv2d(C, Hd, 1)
            s_.dconv = nn.Conv2d(Hd, Hd, 3, 1, 1, groups=Hd)
            s_.fc2 = nn.Conv2d(Hd, C, 1)

    m = MSCANFFN(nn.BatchNorm2d(C), FFN(), nn.Parameter(0.01 * torch.ones(C))).cuda().bfloat16().eval()
    x = torch.randn(a.batch, a.H, a.H, C, device="cuda").bfloat16()
    with torch.no_grad():
        for _ in range(a.reps):
            m.forward_nhwc(x)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
