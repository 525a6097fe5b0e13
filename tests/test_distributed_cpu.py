"""world_size-2 gloo test (CPU) of the data-parallel reconstruction protocol: every rank scales its shard's
gradient by l2_weight / B_global, one all-reduce(SUM) of the flat buffer, identical update on every rank —
the result must equal the single-process full-batch gradient.  Gradients come from the CPU oracle here (no GPU);
the CUDA kernels' gradients are checked against the same oracle in tests/test_recon_gpu.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import approx_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _grads(x, y_old, v_w, h_w, b, global_batch):
    v_w = v_w.clone().requires_grad_(True); h_w = h_w.clone().requires_grad_(True); b = b.clone().requires_grad_(True)
    y = O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1))
    (O.l2_recon_norm(y, y_old).sum() / global_batch).backward()
    return torch.cat([v_w.grad.reshape(-1), h_w.grad.reshape(-1), b.grad.reshape(-1)])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from approx.parallel import shard_range, allreduce_flat_
    torch.manual_seed(0)                                    # same full batch on every rank
    Bg, C, K, N = 6, 5, 3, 7
    x = torch.randn(Bg, C, 8, 9); y_old = torch.randn(Bg, N, 8, 9)
    v_w = torch.randn(K, C, 3, 1); h_w = torch.randn(N, K, 1, 3); b = torch.randn(N)
    lo, hi = shard_range(Bg, rank, world)
    flat = _grads(x[lo:hi], y_old[lo:hi], v_w, h_w, b, Bg)
    allreduce_flat_(flat)
    full = _grads(x, y_old, v_w, h_w, b, Bg)
    q.put((rank, lo, hi, torch.allclose(flat, full, rtol=1e-5, atol=1e-6)))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 6)]
    assert all(r[3] for r in res)


def test_shard_range_covers_ragged_batches():
    from approx.parallel import shard_range
    for n in (1, 7, 8, 256):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _proto_worker(rank, world, port, q):
    """gloo emulation of csrc/peer.cu's protocol: reduce the OWNED slice over all ranks' gradients, AdamW on that slice only
    (sharded optimizer state), all-gather the updated parameters — must equal all-reduce + replicated AdamW."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from approx.recon import owned_slice, plan_flat_layout
    offs, n = plan_flat_layout([(3, 5, 7), (4, 7, 6)])
    torch.manual_seed(0)
    p0 = torch.randn(n)
    torch.manual_seed(100 + rank)
    g = torch.randn(n)
    mask = torch.ones(n)
    for (ov, oh, ob, end) in offs:
        mask[ob:end] = 0                                   # biases: no weight decay
    mask[offs[-1][3]:] = 0

    def adamw(p, gg, mk):
        lr, wd, b1, b2, eps = 1e-2, 0.01, 0.9, 0.999, 1e-8
        m = (1 - b1) * gg; v = (1 - b2) * gg * gg
        return p * (1 - lr * wd * mk) - lr / (1 - b1) * m / (v.sqrt() / (1 - b2) ** 0.5 + eps)
    # reference: all-reduce, then the same AdamW everywhere
    gsum = g.clone(); dist.all_reduce(gsum)
    want = adamw(p0, gsum, mask)
    # protocol: gather everyone's gradients ("pull my slice from every peer"), update my slice, all-gather slices
    allg = [torch.empty(n) for _ in range(world)]
    dist.all_gather(allg, g)
    lo, hi = owned_slice(n, rank, world)
    mine = sum(a[lo:hi] for a in allg)                      # fixed rank order
    upd = adamw(p0[lo:hi], mine, mask[lo:hi])
    per = owned_slice(n, 0, world)[1]
    padded = torch.zeros(per); padded[:hi - lo] = upd
    parts = [torch.empty(per) for _ in range(world)]
    dist.all_gather(parts, padded)
    got = torch.cat([parts[r][:owned_slice(n, r, world)[1] - owned_slice(n, r, world)[0]] for r in range(world)])
    q.put((rank, got.numel() == n, torch.allclose(got, want, rtol=1e-6, atol=1e-7)))
    dist.destroy_process_group()


def test_two_rank_sharded_optimizer_protocol_equals_allreduce_adamw():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_proto_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res


def test_flat_layout_and_owned_slices_partition_the_buffer():
    from approx.recon import owned_slice, plan_flat_layout
    offs, n = plan_flat_layout([(64, 3, 64), (64, 64, 64), (16, 48, 128)])
    assert n % 4 == 0 and offs[0][0] == 0 and all(a[3] == b[0] for a, b in zip(offs, offs[1:])) and offs[-1][3] <= n
    for (K, C, N), (ov, oh, ob, end) in zip([(64, 3, 64), (64, 64, 64), (16, 48, 128)], offs):
        assert (oh - ov, ob - oh, end - ob) == (3 * K * C, 3 * N * K, N)
    for world in (1, 2, 3, 8):
        parts = [owned_slice(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in parts)
