"""Multi-GPU check of the fused reduce-scatter + AdamW + all-gather kernel (csrc/peer.cu), run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu/recon_2rank.py

Every rank fits its shard of a global batch for a few graph-replayed steps; rank r additionally runs the SAME global batch alone
(world-size-1 reconstructor on a one-member group).  The sharded result must match the single-GPU full-batch result, all ranks
must hold bit-identical parameters, and no peer barrier may have timed out.  Prints RECON_MULTI_OK on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo_groups = [dist.new_group([r]) for r in range(world)]
    from approx.layers import LowRankExpConvV2
    from approx.parallel import shard_range
    from approx.recon import LayerwiseReconstructor
    shapes = [(64, 32, 128, 28), (16, 16, 64, 20), (128, 64, 128, 14)]
    Bg = 4 * world

    def make():
        torch.manual_seed(7)
        layers = [LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N).to(dev).bfloat16() for (C, K, N, H) in shapes]
        xs = [torch.randn(Bg, H, H, C, device=dev).bfloat16() for (C, K, N, H) in shapes]
        ys = [torch.randn(Bg, H, H, N, device=dev).bfloat16() for (C, K, N, H) in shapes]
        return layers, xs, ys

    steps = 4
    layers, xs, ys = make()
    lo, hi = shard_range(Bg, rank, world)
    rec = LayerwiseReconstructor(layers, lr=1e-3, process_group=None)
    assert rec.world == world
    rec.bind([x[lo:hi].contiguous() for x in xs], [y[lo:hi].contiguous() for y in ys], Bg)
    rec.capture(parallel=True)
    for _ in range(steps):
        rec.replay()
    torch.cuda.synchronize()
    rec.check()
    sharded = rec.param.clone()

    layers, xs, ys = make()
    solo = LayerwiseReconstructor(layers, lr=1e-3, process_group=solo_groups[rank])
    assert solo.world == 1
    solo.bind(xs, ys, Bg)
    for _ in range(steps):
        solo.step(parallel=False)
    torch.cuda.synchronize()
    rel = ((sharded - solo.param).norm() / solo.param.norm()).item()
    assert rel < 2e-4, f"rank {rank}: sharded vs full-batch parameters differ by {rel}"
    gathered = [torch.empty_like(sharded) for _ in range(world)]
    dist.all_gather(gathered, sharded)
    assert all(torch.equal(gathered[0], g) for g in gathered), "ranks hold different parameters"
    assert rec.steps_taken() == steps
    dist.barrier()
    solo.close(); rec.close()
    if rank == 0:
        print(f"RECON_MULTI_OK world={world} rel={rel:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
