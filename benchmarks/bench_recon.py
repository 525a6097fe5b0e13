#!/usr/bin/env python
"""BASELINE config 3: layer-wise filter reconstruction of VGG-16 conv layers, batch-sharded over N B200s with one
NCCL all-reduce of the factor gradients per step.

For each selected VGG-16 conv (default: conv2_1 64->128@112, conv3_2 256->256@56, conv4_2 512->512@28) a scheme-2
layer (rank K) is fitted to the dense conv's output: step = fused forward + output-L2 loss/gradient + tap-conv
weight/data gradients + all-reduce(SUM) + fused AdamW (approx.recon.Scheme2Reconstructor).  The teacher outputs are
computed once before the timed region (the reference's asymmetric mode caches them per batch, hooks/finetune.py:294-299).
Global batch is fixed (strong scaling): every rank takes global_batch / world images.

    python benchmarks/bench_recon.py [--global-batch 256] [--rank 64] [--steps 10]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/bench_recon.py
Prints one JSON line per layer and a total, on rank 0; times are CUDA-event, max over ranks.
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
import torch.distributed as dist  # noqa: E402

LAYERS = {"conv1_2": (64, 64, 224), "conv2_1": (64, 128, 112), "conv2_2": (128, 128, 112), "conv3_2": (256, 256, 56),
          "conv4_2": (512, 512, 28), "conv5_2": (512, 512, 14)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=256)
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layers", default="conv2_1,conv3_2,conv4_2")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--weak", action="store_true", help="weak scaling: --global-batch images PER GPU (global batch = that x world size)")
    ap.add_argument("--graph", action="store_true", help="capture the step in a CUDA graph and replay it (no host work per step)")
    args = ap.parse_args()
    from approx import backend as B
    from approx.layers import LowRankExpConvV2
    from approx.parallel import shard_range
    from approx.recon import Scheme2Reconstructor

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    if args.weak:
        args.global_batch *= world
    lo, hi = shard_range(args.global_batch, rank, world)
    total_ms = 0.0
    for name in args.layers.split(","):
        C, N, H = LAYERS[name]
        torch.manual_seed(42)                                           # identical initial factors on every rank
        teacher = torch.nn.Conv2d(C, N, 3, 1, 1).to(dev).to(dt)
        layer = LowRankExpConvV2(C, args.rank, (3, 3), (1, 1), (1, 1), out_channels=N).to(dev).to(dt)
        rec = Scheme2Reconstructor(layer, lr=1e-4, weight_decay=0.01)
        torch.manual_seed(1000 + rank)
        x = torch.randn(hi - lo, H, H, C, device=dev).to(dt)            # this rank's shard, NHWC, resident in HBM
        with torch.no_grad():                                            # teacher output once, outside the timed region
            y_old = teacher(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous()
        step = lambda: rec.step(x, y_old, args.global_batch)
        if args.graph:
            rec.capture(x, y_old, args.global_batch)
            step = rec.replay
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = B.launch_count()
        a.record()
        for _ in range(args.steps):
            loss = step()
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        total_ms += ms.item()
        if rank == 0:
            print(json.dumps({"bench": "recon_step", "layer": name, "C": C, "N": N, "H": H, "rank_K": args.rank,
                              "global_batch": args.global_batch, "n_gpus": world, "dtype": args.dtype, "cuda_graph": bool(args.graph),
                              "ms_per_step": ms.item(),
                              "images_per_s": args.global_batch / ms.item() * 1e3, "loss": float(loss),
                              "launches_per_step": (B.launch_count() - n0) / args.steps,
                              "allreduce_bytes": rec.grad.numel() * 4}))
    if rank == 0:
        print(json.dumps({"bench": "recon_total", "layers": args.layers, "n_gpus": world, "global_batch": args.global_batch,
                          "ms_per_step_sum": total_ms, "images_per_s": args.global_batch / total_ms * 1e3,
                          "scaling": "weak" if args.weak else "strong", "cuda_graph": bool(args.graph)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
