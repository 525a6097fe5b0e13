#!/usr/bin/env python
"""BASELINE config 4: MSCA block (SegNeXt MSCAN-T shapes at 512x512 input, batch 32), un-reparameterised vs
re-parameterised forward, on one B200.

Times each of the four MSCAN-T stage shapes (C,H) = (32,128) (64,64) (160,32) (256,16) for
  original            MSCA with 3 strip cascades + identity
  rep d0+fix          single 21x21 depthwise kernel + border fix  (MscaRep(decomp=0, fix=True))
  rep d1+fix          one 1x21 . 21x1 cascade + border fix        (MscaRep(decomp=1, fix=True))
through this backend (bf16 and fp32), and — with --torch — through stock torch modules on the same GPU
(cuDNN, channels_last bf16): the library path the kernels have to beat.  Prints one JSON object per line.

    python benchmarks/bench_msca.py [--batch 32] [--iters 20] [--dtype bf16|f32] [--torch] [--cpu]
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

SHAPES = [(32, 128, 3), (64, 64, 3), (160, 32, 5), (256, 16, 2)]   # (C, H=W, number of blocks in MSCAN-T)


def time_cuda(fn, iters, warmup=3):
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


def torch_msca(C):
    """Stock-torch restatement of the reference module (layers/msca.py, depth_seperable_conv.py) for the GPU library baseline."""
    from torch import nn

    class Cascade(nn.Module):
        def __init__(s, k):
            super().__init__()
            s.c1 = nn.Conv2d(C, C, (1, k), padding=(0, k // 2), groups=C)
            s.c2 = nn.Conv2d(C, C, (k, 1), padding=(k // 2, 0), groups=C)

        def forward(s, x):
            return s.c2(s.c1(x))

    class M(nn.Module):
        def __init__(s):
            super().__init__()
            s.conv0 = nn.Conv2d(C, C, 5, padding=2, groups=C)
            s.br = nn.ModuleList([Cascade(k) for k in (7, 11, 21)])
            s.mix = nn.Conv2d(C, C, 1)

        def forward(s, x):
            a = s.conv0(x)
            return x * s.mix(sum([b(a) for b in s.br] + [a]))
    return M()


EXP = {32: 8, 64: 8, 160: 4, 256: 4}      # MSCAN-T exp_ratios per stage width (reference models/mscan.py:141)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--torch", action="store_true", help="also time stock torch/cuDNN on the GPU")
    ap.add_argument("--block", action="store_true",
                    help="also time the attention half of an MSCAN block (norm1 -> proj_1 -> GELU -> MSCA -> proj_2 -> + shortcut, "
                         "layer scale, residual: models/mscan.py:72-80,108-109) fused into four kernels, and the same graph in torch/cuDNN")
    args = ap.parse_args()
    from approx import backend as B
    from approx.layers import MSCA
    from approx.core import MscaRep

    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    esz = 2 if dt == torch.bfloat16 else 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    total = {}
    for (C, H, nblk) in SHAPES:
        torch.manual_seed(42)
        src = MSCA(C, 5, (7, 11, 21))
        variants = {"original": src}
        for name, (d, f) in {"rep_d0_fix": (0, True), "rep_d1_fix": (1, True)}.items():
            app = MscaRep(decomp=d, fix=f)
            sub = app.initialize(src)
            import io, contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                app.optimize(sub)
            variants[name] = app.postprocess(sub)
        x = torch.randn(args.batch, H, H, C, device="cuda").to(dt)          # NHWC resident in HBM
        elems = args.batch * H * H * C
        n_in = max(4, -(-300_000_000 // (3 * elems * esz)))
        xs = [x] + [torch.randn_like(x) for _ in range(n_in - 1)]
        for name, m in variants.items():
            m = m.cuda().to(dt).eval()
            with torch.no_grad():
                n0 = B.launch_count()
                m.forward_nhwc(x)
                launches = B.launch_count() - n0
                eager_ms = time_cuda(lambda: m.forward_nhwc(x), args.iters)
                # device time: n_in forwards on distinct inputs (> L2 in total) captured as one CUDA graph
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    m.forward_nhwc(xs[0])
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    keep = [m.forward_nhwc(xi) for xi in xs]
                ms = time_cuda(graph.replay, args.iters) / len(xs)
                del keep, graph
            rec = {"bench": "msca_block", "variant": name, "dtype": args.dtype, "C": C, "H": H, "batch": args.batch,
                   "ms": ms, "eager_ms": eager_ms, "launches": launches, "images_per_s": args.batch / ms * 1e3,
                   "min_traffic_GBps": 2 * elems * esz / ms / 1e6, "frac_hbm_min_traffic": 2 * elems * esz / ms / 1e6 / hbm}
            print(json.dumps(rec))
            total[name] = total.get(name, 0.0) + ms * nblk
        if args.block and dt == torch.bfloat16:
            from torch import nn
            from approx.layers import MSCANAttention

            class Attn(nn.Module):
                def __init__(s_):
                    super().__init__()
                    s_.proj_1 = nn.Conv2d(C, C, 1)
                    s_.spatial_gating_unit = MSCA(C, 5, (7, 11, 21))
                    s_.proj_2 = nn.Conv2d(C, C, 1)

            blk = MSCANAttention(nn.BatchNorm2d(C), Attn(), nn.Parameter(0.01 * torch.ones(C))).cuda().to(dt).eval()
            with torch.no_grad():
                blk.forward_nhwc(xs[0])
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    blk.forward_nhwc(xs[0])
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                with torch.cuda.graph(graph):
                    keep = [blk.forward_nhwc(xi) for xi in xs]
                ms = time_cuda(graph.replay, args.iters) / len(xs)
                del keep, graph
                # the same graph through stock torch modules (cuDNN / cuBLAS, channels_last bf16)
                tm = torch_msca(C).cuda().to(dt).to(memory_format=torch.channels_last).eval()
                bn = nn.BatchNorm2d(C).cuda().to(dt).eval()
                p1, p2 = nn.Conv2d(C, C, 1).cuda().to(dt), nn.Conv2d(C, C, 1).cuda().to(dt)
                ls = (0.01 * torch.ones(C, device="cuda", dtype=dt)).reshape(1, C, 1, 1)
                xt = xs[0].permute(0, 3, 1, 2)

                def torch_block():
                    n = bn(xt)
                    return xt + ls * (p2(tm(torch.nn.functional.gelu(p1(n)))) + n)
                ms_t = time_cuda(torch_block, args.iters)
            print(json.dumps({"bench": "mscan_attention_half", "dtype": args.dtype, "C": C, "H": H, "batch": args.batch, "ms": ms,
                              "torch_cudnn_ms": ms_t, "speedup_over_torch": ms_t / ms, "kernels": 4,
                              "min_traffic_GBps": 2 * elems * esz / ms / 1e6}))
            total["attention_half_fused"] = total.get("attention_half_fused", 0.0) + ms * nblk
            total["attention_half_torch_cudnn"] = total.get("attention_half_torch_cudnn", 0.0) + ms_t * nblk
            # ---- the whole block: attention half + FFN half (fc1 -> 3x3 depthwise -> GELU -> fc2, models/mscan.py:27-62, 108-112)
            from approx.layers import MSCANBlock
            Hd = C * EXP[C]

            class FFN(nn.Module):
                def __init__(s_):
                    super().__init__()
                    s_.fc1 = nn.Conv2d(C, Hd, 1)
                    s_.dconv = nn.Conv2d(Hd, Hd, 3, 1, 1, groups=Hd)
                    s_.fc2 = nn.Conv2d(Hd, C, 1)

            ffn = FFN()
            full = MSCANBlock(nn.BatchNorm2d(C), Attn(), nn.Parameter(0.01 * torch.ones(C)), nn.BatchNorm2d(C), ffn,
                              nn.Parameter(0.01 * torch.ones(C))).cuda().to(dt).eval()
            with torch.no_grad():
                n0 = B.launch_count()
                full.forward_nhwc(xs[0])
                launches = B.launch_count() - n0
                full.forward_nhwc(xs[0])
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    full.forward_nhwc(xs[0])
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                with torch.cuda.graph(graph):
                    keep = [full.forward_nhwc(xi) for xi in xs]
                ms_b = time_cuda(graph.replay, args.iters) / len(xs)
                del keep, graph
                bn2 = nn.BatchNorm2d(C).cuda().to(dt).eval()
                tf = FFN().cuda().to(dt).to(memory_format=torch.channels_last).eval()

                def torch_full():
                    x1 = torch_block()
                    return x1 + ls * tf.fc2(torch.nn.functional.gelu(tf.dconv(tf.fc1(bn2(x1)))))
                ms_tb = time_cuda(torch_full, args.iters)
            hidden_bytes = args.batch * H * H * Hd * esz
            print(json.dumps({"bench": "mscan_block", "dtype": args.dtype, "C": C, "hidden": Hd, "H": H, "batch": args.batch, "ms": ms_b,
                              "ffn_half_ms": ms_b - ms, "torch_cudnn_ms": ms_tb, "speedup_over_torch": ms_tb / ms_b, "kernels": launches,
                              "ffn_hbm_GBps_on_4_hidden_passes": 4 * hidden_bytes / max(ms_b - ms, 1e-6) / 1e6}))
            total["block_fused"] = total.get("block_fused", 0.0) + ms_b * nblk
            total["block_torch_cudnn"] = total.get("block_torch_cudnn", 0.0) + ms_tb * nblk
        if args.torch:
            tm = torch_msca(C).cuda().to(dt).to(memory_format=torch.channels_last).eval()
            xt = x.permute(0, 3, 1, 2)
            with torch.no_grad():
                ms = time_cuda(lambda: tm(xt), args.iters)
            print(json.dumps({"bench": "msca_block", "variant": "torch_cudnn_original", "dtype": args.dtype, "C": C, "H": H,
                              "batch": args.batch, "ms": ms, "images_per_s": args.batch / ms * 1e3}))
            total["torch_cudnn_original"] = total.get("torch_cudnn_original", 0.0) + ms * nblk
    for name, ms in total.items():
        print(json.dumps({"bench": "mscan_t_all_13_msca_blocks", "variant": name, "dtype": args.dtype, "batch": args.batch,
                          "ms": ms, "images_per_s": args.batch / ms * 1e3}))


if __name__ == "__main__":
    main()
