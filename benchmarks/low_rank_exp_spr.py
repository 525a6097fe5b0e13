#!/usr/bin/env python
"""The reference's scheme-1 micro-benchmark (scripts/low_rank_exp_spr.py:11-63) on this backend: speed-up of
`LowRankExpConvV1(num_base = 2..11)` over the dense conv it replaces, for AlexNet's conv2-5 shapes at batch 128.

The reference times `module(x)` on the host clock (10 warm-up, 100 repeats); here the same loop is timed with CUDA events, for
  dense   stock nn.Conv2d through torch/cuDNN (the layer being replaced), in the run's dtype
  ours    LowRankExpConvV1 of this backend (fused tcgen05 kernel for bf16 and num_base <= 16, unfused kernels otherwise)
One JSON object per (layer, num_base).

    python benchmarks/low_rank_exp_spr.py [--decomp] [--dtype bf16|f32] [--batch 128]
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
from torch import nn  # noqa: E402


def inference(module, x, repeat=100, warmup=10):
    with torch.no_grad():
        for _ in range(warmup):
            module(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(repeat):
            module(x)
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / repeat * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--decomp', action='store_true')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--repeat', type=int, default=50)
    args = ap.parse_args()
    from approx import backend as B
    from approx.layers import LowRankExpConvV1
    dt = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    b, h, w = args.batch, 224, 224
    cfgs = [[(64, 192, 5, 1, 2), (b, 64, h // 8, w // 8)], [(192, 384, 3, 1, 1), (b, 192, h // 16, w // 16)],
            [(384, 256, 3, 1, 1), (b, 384, h // 32, w // 32)], [(256, 256, 3, 1, 1), (b, 256, h // 32, w // 32)]]
    for conv_cfg, x_cfg in cfgs:
        x = torch.randn(*x_cfg, device='cuda').to(dt).contiguous(memory_format=torch.channels_last)
        dense = nn.Conv2d(*conv_cfg).cuda().to(dt).to(memory_format=torch.channels_last).eval()
        std_time = inference(dense, x, args.repeat)
        for m in range(2, 12):
            layer = LowRankExpConvV1(*conv_cfg, num_base=m, decomp=args.decomp).cuda().to(dt).eval()
            with torch.no_grad():
                layer(x)
                n0 = B.launch_count()
                layer(x)
                launches = B.launch_count() - n0
            t = inference(layer, x, args.repeat)
            print(json.dumps({"bench": "low_rank_exp_spr", "conv": conv_cfg, "input": list(x_cfg), "dtype": args.dtype,
                              "decomp": bool(args.decomp), "num_base": m, "dense_cudnn_ms": std_time * 1e3, "ours_ms": t * 1e3,
                              "speedup_vs_dense": std_time / t, "kernels_per_forward": launches,
                              "path": "fused tcgen05" if launches == 1 else "unfused"}))


if __name__ == '__main__':
    main()
