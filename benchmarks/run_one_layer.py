#!/usr/bin/env python
"""Runs the fused scheme-2 conv on one layer shape a few times (the command line ncu profiles).

    python benchmarks/run_one_layer.py --C 64 --N 64 --K 64 --H 224 --batch 64 [--pool 1] [--iters 3]
"""
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
    for k, v in dict(C=64, N=64, K=64, H=224, batch=64, pool=0, iters=3).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    a = ap.parse_args()
    from approx import backend as B
    torch.manual_seed(0)
    v_w = torch.randn(a.K, a.C, 3, 1, device="cuda") / (3 * a.C) ** 0.5
    h_w = torch.randn(a.N, a.K, 1, 3, device="cuda") / (3 * a.K) ** 0.5
    bias = torch.randn(a.N, device="cuda")
    wv, wh = B.lowrank2_pack(v_w, h_w)
    x = torch.randn(a.batch, a.H, a.H, B.round_up(a.C, 16), device="cuda").bfloat16()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    ev[0].record()
    for i in range(a.iters):
        y = B.lowrank2_fused(x, wv, wh, bias, relu=True, pool=bool(a.pool))
        ev[i + 1].record()
    torch.cuda.synchronize()
    print("ms per launch:", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(a.iters)], "checksum", float(y.float().abs().mean()))


if __name__ == "__main__":
    main()
