#!/usr/bin/env python
"""BASELINE config 5: rank-K x batch sweep of ONE fused separable (scheme-2) conv layer, next to the reference's CPU path.

The reference's own micro-benchmark of this path is scripts/low_rank_exp_spr.py:11-63 (conv layer vs its low-rank
replacement over a range of ranks, 10 warm-up + 100 timed calls); this is the same sweep for the true scheme-2 layer on the
B200 backend: K in {4, 8, 16, 32, 64} x batch in {1, 4, 16, 64, 256, 1024} on a VGG conv3_2-shaped layer (256 -> 256 at 56x56)
by default.  Per point: images/s through `cna_lowrank2_fused_forward` (CUDA events, inputs resident in HBM), the layer's
algorithmic HBM bytes / time against the measured copy peak, and — with --cpu — the reference's two stock convs on the host
cores (fp32, all threads) on a bounded sample.  Under torchrun every rank runs the same sweep on its own GPU (batch is per GPU,
no collective) and rank 0 reports the sum.  One JSON object per line.

    python benchmarks/sweep_lowrank2.py [--C 256 --N 256 --H 56] [--cpu] [--batches 1,4,16,64,256,1024] [--ranks 4,8,16,32,64]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=256)
    ap.add_argument("--N", type=int, default=256)
    ap.add_argument("--H", type=int, default=56)
    ap.add_argument("--ranks", default="4,8,16,32,64")
    ap.add_argument("--batches", default="1,4,16,64,256,1024")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu", action="store_true", help="also time the reference's CPU path (two stock convs, fp32) per rank K")
    a = ap.parse_args()
    from approx import backend as B
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    C, N, H = a.C, a.N, a.H
    cp = B.round_up(C, 16)
    for K in [int(v) for v in a.ranks.split(",")]:
        torch.manual_seed(42)
        v_w = torch.randn(K, C, 3, 1, device=dev) / (3 * C) ** 0.5
        h_w = torch.randn(N, K, 1, 3, device=dev) / (3 * K) ** 0.5
        bias = torch.zeros(N, device=dev)
        wv, wh = B.lowrank2_pack(v_w, h_w)
        cpu_ips = None
        if a.cpu and rank == 0:
            import torch.nn.functional as F
            torch.set_num_threads(os.cpu_count())
            xc = torch.randn(8, C, H, H)
            vc, hc = v_w.cpu(), h_w.cpu()
            with torch.no_grad():
                for _ in range(2):
                    F.conv2d(F.conv2d(xc, vc, None, 1, (1, 0)), hc, bias.cpu(), 1, (0, 1))
                t0 = time.perf_counter()
                for _ in range(5):
                    F.conv2d(F.conv2d(xc, vc, None, 1, (1, 0)), hc, bias.cpu(), 1, (0, 1))
                cpu_ips = 8 * 5 / (time.perf_counter() - t0)
        for Bn in [int(v) for v in a.batches.split(",")]:
            x = torch.randn(Bn, H, H, cp, device=dev).bfloat16()
            for _ in range(3):
                B.lowrank2_fused(x, wv, wh, bias)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                B.lowrank2_fused(x, wv, wh, bias)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms = ms.item()
            if rank == 0:
                byts = Bn * (2 * H * H * (C + N)) + 6 * K * (C + N) + 2 * N
                rec = {"bench": "lowrank2_sweep", "C": C, "N": N, "H": H, "rank_K": K, "batch_per_gpu": Bn, "n_gpus": world,
                       "ms": round(ms, 4), "images_per_s": round(world * Bn / ms * 1e3, 1), "hbm_GBps": round(byts / ms / 1e6, 1),
                       "frac_hbm_peak": round(byts / ms / 1e6 / hbm, 3),
                       "tflops": round(6.0 * H * H * K * (C + N) * Bn / ms / 1e9, 1)}
                if cpu_ips is not None:
                    rec["cpu_reference_images_per_s"] = round(cpu_ips, 1)
                    rec["cpu_cores"] = os.cpu_count()
                print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
