#!/usr/bin/env python
"""Scheme-1 low-rank layers (LowRankExpConvV1, reference layers/low_rank_conv.py:28-66) at the AlexNet shapes of the reference's
shipped configs (configs/low-rank-exp/low-rank-exp-v1_l2345_svd_alexnet.py: M = 8 bases), on one B200.

For each (C, N, d, H) times, in bf16 NHWC through this backend,
  v1          grouped d x d `s_conv` (channel multiplier M, with bias) then the 1x1 `d_conv`
  v1 decomp   the separable (d,1) . (1,d) form after `decomp()` then the same 1x1
each as the fused tcgen05 kernel (csrc/lowrank1_tc.cu; `us`) and as separate depthwise + 1x1 launches (`unfused_us`)
and — with --torch — stock torch/cuDNN on the same GPU (bf16, channels_last): the dense conv the layer replaces and the
reference's own nn.Conv2d chain.  Prints one JSON object per line with the algorithmic bytes / FLOPs of SURVEY §8d
(bytes = 2 (C + N) per pixel, FLOPs = 2 C M (d^2 + N), or 2 C M (2 d + N) after decomp).

    python benchmarks/bench_lowrank1.py [--batch 64] [--iters 30] [--torch]
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

# AlexNet conv2..conv5 as the reference's torchvision-style AlexNet builds them: (C, N, d, pad, H = W of the layer input)
SHAPES = [(64, 192, 5, 2, 27), (192, 384, 3, 1, 13), (384, 256, 3, 1, 13), (256, 256, 3, 1, 13)]
M = 8


def time_cuda(fn, iters, warmup=5):
    for _ in range(warmup):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--torch", action="store_true", help="also time stock torch/cuDNN on the GPU")
    args = ap.parse_args()
    from approx import backend as B
    from approx.layers import LowRankExpConvV1, _ops

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6552.6))
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    torch.manual_seed(42)
    for (C, N, d, pad, H) in SHAPES:
        Bn = args.batch
        x = torch.randn(Bn, C, H, H, device=dev).to(dt)
        xn = B.to_nhwc(x)
        px = Bn * H * H
        abytes = 2 * (C + N) * px
        out = dict(shape=f"{C}->{N} {d}x{d} @{H}", batch=Bn, M=M, algorithmic_MB=round(abytes / 1e6, 2))
        for name, dec in (("v1", False), ("v1_decomp", True)):
            layer = LowRankExpConvV1(C, N, d, 1, pad, M).to(dev).eval()
            if dec:
                layer.decomp()
            layer = layer.to(dt)
            with torch.no_grad():
                ms = time_cuda(lambda: layer.forward_nhwc(xn), args.iters)                     # fused tcgen05 kernel
                ms_u = time_cuda(lambda: layer.forward_nhwc(xn, fused=False), args.iters)      # depthwise kernels + 1x1
                t = layer.s_conv.forward_nhwc(xn) if dec else _ops.dw(layer.s_conv, xn)
                ms_s = time_cuda(lambda: layer.s_conv.forward_nhwc(xn) if dec else _ops.dw(layer.s_conv, xn), args.iters)
                ms_d = time_cuda(lambda: _ops.dense_taps(layer.d_conv, t), args.iters)
            flops = 2 * C * M * ((2 * d if dec else d * d) + N) * px
            out[name] = dict(us=round(ms * 1e3, 1), hbm_frac=round(abytes / (ms * 1e-3) / 1e9 / hbm, 3),
                             tflops=round(flops / (ms * 1e-3) / 1e12, 1),
                             unfused_us=round(ms_u * 1e3, 1), s_conv_us=round(ms_s * 1e3, 1), d_conv_us=round(ms_d * 1e3, 1),
                             intermediate_MB=round(2 * 2 * C * M * px * (2 if dec else 1) / 1e6, 1))
        if args.torch:
            from torch import nn
            dense = nn.Conv2d(C, N, d, 1, pad).to(dev).to(dt).to(memory_format=torch.channels_last).eval()
            ref = nn.Sequential(nn.Conv2d(C, C * M, d, 1, pad, groups=C), nn.Conv2d(C * M, N, 1)).to(dev).to(dt)
            ref = ref.to(memory_format=torch.channels_last).eval()
            xc = x.contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                out["torch_dense_us"] = round(time_cuda(lambda: dense(xc), args.iters) * 1e3, 1)
                out["torch_v1_us"] = round(time_cuda(lambda: ref(xc), args.iters) * 1e3, 1)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
