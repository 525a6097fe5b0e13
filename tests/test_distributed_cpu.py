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
