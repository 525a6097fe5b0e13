"""Layer-wise reconstruction of a scheme-2 low-rank layer against its source conv's output.

Restates the reference's L2Reconstruct inner step (src/approx/hooks/finetune.py:293-331) for ONE replaced layer,
with only that layer trainable (models/switchable.py:63-84 `freeze_except`):

    y_new = layer(x);  norm_b = ||y_new_b - y_old_b||_2;  loss = l2_weight * mean_b norm_b
    backward;  all-reduce(SUM) of the factor gradients over the data-parallel ranks;  AdamW step

All arithmetic runs in this library's CUDA kernels: the fused tcgen05 forward, the loss/gradient kernel, tap-conv
data/weight gradients and a fused AdamW on one flat fp32 buffer (master weights; the layer's bf16/fp32 parameters
are refreshed from it).  torch.distributed (NCCL) provides the single collective per step — one all-reduce of the
flat gradient buffer, e.g. 0.79 MB for C = N = 512, K = 64.
"""
import os
from typing import Optional

import torch
import torch.distributed as dist

from approx.parallel import allreduce_flat_

from approx import backend as B
from approx.layers import LowRankExpConvV2
from approx.layers import _ops


class Scheme2Reconstructor:
    def __init__(self, layer: LowRankExpConvV2, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.01, l2_weight: float = 1.0, process_group=None):
        if not layer.dense:
            raise ValueError("reconstruction is implemented for the dense (true scheme-2) h_conv")
        kh, kw = layer.kernel_size
        if layer.stride != (1, 1) or layer.padding != (kh // 2, kw // 2) or kh % 2 == 0 or kw % 2 == 0:
            raise NotImplementedError("reconstruction kernels cover stride-1 'same' convolutions")
        if layer.fuse_relu or layer.fuse_pool:
            # the gradients below are those of the plain h_conv(v_conv(x)) the reference fits (its teacher is the conv output)
            raise ValueError("reconstruction differentiates the bare layer: clear fuse_relu / fuse_pool first")
        self.layer = layer
        self.lr, self.betas, self.eps, self.wd, self.l2_weight = lr, betas, eps, weight_decay, l2_weight
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        v, h, b = layer.v_conv.weight, layer.h_conv.weight, layer.h_conv.bias
        B.require_cuda(v)
        self.sizes = (v.numel(), h.numel(), b.numel())
        n = sum(self.sizes)
        dev = v.device
        self.master = torch.cat([v.detach().float().reshape(-1), h.detach().float().reshape(-1),
                                 b.detach().float().reshape(-1)]).contiguous()
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.t = 0
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)     # device-side step counter (graph mode)
        self._graph = None
        self._graph_io = None
        self._packs = None                                                 # persistent packed weights (graph mode)

    def _views(self, flat):
        nv, nh, nb = self.sizes
        L = self.layer
        return (flat[:nv].view(L.v_conv.weight.shape[0], L.v_conv.weight.shape[1], L.kernel_size[0]),
                flat[nv:nv + nh].view(L.h_conv.weight.shape[0], L.h_conv.weight.shape[1], L.kernel_size[1]),
                flat[nv + nh:])

    def gradients(self, x: torch.Tensor, y_old: torch.Tensor, global_batch: Optional[int] = None):
        """Fills self.grad with d loss / d (v_w, h_w, bias) for this rank's shard; returns the per-sample norms.
        x: (B,H,W,C) NHWC (C exact, or zero-padded to 16 for the fused bf16 path); y_old: (B,H,W,N)."""
        L = self.layer
        Bn = x.shape[0]
        gb = Bn * self.world if global_batch is None else global_batch
        dv, dh, db = self._views(self.grad)
        C, K, N = L.in_channels, L.num_base, L.out_channels
        _, H, W, cx = x.shape
        if L._fusable(x.dtype, H, W) and cx == B.round_up(C, 16) and L.kernel_size == (3, 3):
            # tensor-core path: the forward writes T = v_conv(x) next to y; both weight gradients and the data gradient
            # through h_conv are tcgen05 kernels fed straight from the NHWC tensors
            if self._packs is not None:
                # graph mode: re-pack from the fp32 master weights into fixed buffers inside the captured step
                mv, mh, mb = self._views(self.master)
                wv, wh, wd = self._packs
                B.lowrank2_pack_into(mv, mh, wv, wh)
                B.lowrank2_pack_into(mh.permute(1, 0, 2).flip(2).contiguous(), self._dummy_h, wd, self._dummy_wh)
                bias = mb
            else:
                wv, wh = _ops.cached(L, "lr2_pack", (L.v_conv.weight, L.h_conv.weight), torch.bfloat16,
                                     lambda: B.lowrank2_pack(L.v_conv.weight, L.h_conv.weight))
                wd = _ops.cached(L.h_conv, "lr2_dgrad_pack", (L.h_conv.weight,), torch.bfloat16,
                                 lambda: B.lowrank2_pack(L.h_conv.weight.detach().permute(1, 0, 3, 2).flip(2).contiguous(),
                                                         torch.zeros(64, K, 1, 3, device=x.device))[0])
                bias = _ops.bias_f32(L.h_conv)
            y, t = B.lowrank2_fused_with_t(x, wv, wh, bias)
            norms, g = B.l2_recon(y, y_old.contiguous(), self.l2_weight / gb, db=db)     # loss, dL/dy and dL/db in two passes
            B.wgrad3_tc(g, t, dh, N, K, 1, 0)                               # dL/dWh[n][k][t]
            # dL/dT = g convolved along W with h_conv's flipped, transposed taps (wd: a (K, N, 3, 1) "v_conv" weight)
            gt = B.tapconv3_tc(g, wd, 1)
            B.wgrad3_tc(x, gt, dv, C, K, 0, 1)                              # dL/dWv[k][c][t]
            return norms
        y = L.forward_nhwc(x)
        norms, g = B.l2_recon(y, y_old.contiguous(), self.l2_weight / gb)
        xg = x if x.shape[3] == C else x[..., :C].contiguous()
        t = _ops.dense_taps(L.v_conv, xg)                                    # T recomputed: never stored by the fused forward
        B.conv_taps_wgrad(t, g, L.kernel_size[1], 1, L.padding[1], dw_out=dh, db_out=db)
        wh_t = _ops.cached(L.h_conv, "taps_w_dgrad", (L.h_conv.weight,), x.dtype,
                           lambda: B.pack_taps_dgrad(L.h_conv.weight, x.dtype))
        gt = B.conv_taps(g, wh_t, None, L.num_base, L.kernel_size[1], 1, (1, 1), L.kernel_size[1] - 1 - L.padding[1])
        B.conv_taps_wgrad(xg, gt, L.kernel_size[0], 0, L.padding[0], dw_out=dv)
        return norms

    def apply(self):
        """All-reduce the flat gradient, AdamW on the master copy, refresh the layer's parameters."""
        allreduce_flat_(self.grad, self.group)
        self.t += 1
        # the flat buffer is [v | h | bias]: the bias tail gets no weight decay, like the reference's timm-built optimizer
        # (create_optimizer_v2's default filter_bias_and_bn, hooks/finetune.py:188)
        nd = self.sizes[0] + self.sizes[1]
        if self._packs is not None:
            B.adamw_step_dev_(self.master, self.grad, self.exp_avg, self.exp_avg_sq, self.step_dev, self.lr, self.betas, self.eps, self.wd,
                              n_decay=nd)
        else:
            B.adamw_step_(self.master, self.grad, self.exp_avg, self.exp_avg_sq, self.t, self.lr, self.betas, self.eps, self.wd,
                          n_decay=nd)
        mv, mh, mb = self._views(self.master)
        L = self.layer
        with torch.no_grad():
            L.v_conv.weight.copy_(mv.view_as(L.v_conv.weight))
            L.h_conv.weight.copy_(mh.view_as(L.h_conv.weight))
            L.h_conv.bias.copy_(mb)

    def step(self, x, y_old, global_batch=None):
        """One reconstruction step; returns this rank's mean per-sample norm (device scalar tensor)."""
        norms = self.gradients(x, y_old, global_batch)
        self.apply()
        return norms.mean()

    # ------------------------------------------------------------------------------------------------------------------
    # CUDA-graph mode: the whole step (pack, fused forward + T, loss/gradient, three tcgen05 backward kernels, NCCL
    # all-reduce, AdamW, parameter refresh) is captured once and replayed, so a step costs no host work.  At 8 GPUs a
    # step of the tensor-core path is ~0.2 ms of kernels; ~15 launches + one collective issued from Python would
    # otherwise dominate it.
    # ------------------------------------------------------------------------------------------------------------------
    def capture(self, x: torch.Tensor, y_old: torch.Tensor, global_batch: Optional[int] = None):
        """Captures one step on static buffers `x` / `y_old` (copy new minibatches into them between replays).
        NOTE: the two warm-up steps before the capture are real optimizer steps on (x, y_old): `self.t` and the device
        counter both read 2 more afterwards and the parameters have moved."""
        L = self.layer
        C, K, N = L.in_channels, L.num_base, L.out_channels
        _, H, W, cx = x.shape
        if not (L._fusable(x.dtype, H, W) and cx == B.round_up(C, 16) and L.kernel_size == (3, 3)):
            raise NotImplementedError("graph mode covers the tensor-core path (bf16, 3x3, stride 1)")
        if self.world > 1 and not os.environ.get("CNA_RECON_GRAPH_MULTI"):
            # capturing the NCCL all-reduce hung in the one 8-GPU attempt of round 1 (cause not yet investigated)
            raise NotImplementedError("graph mode is validated single-GPU only; multi-GPU reconstruction runs eagerly")
        dev = x.device
        kp, cp = B.lowrank2_k_pad(K), B.round_up(C, 16)
        self._packs = (torch.zeros(3, kp, cp, device=dev, dtype=torch.bfloat16), torch.zeros(3, N, kp, device=dev, dtype=torch.bfloat16),
                       torch.zeros(3, kp, N, device=dev, dtype=torch.bfloat16))
        self._dummy_h = torch.zeros(64, K, 3, device=dev)
        self._dummy_wh = torch.zeros(3, 64, kp, device=dev, dtype=torch.bfloat16)
        self.step_dev.fill_(self.t)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the capture: one-time attribute / cache setup
            for _ in range(2):
                self.step(x, y_old.contiguous(), global_batch)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        t_before = self.t
        with torch.cuda.graph(self._graph):
            self._graph_loss = self.step(x, y_old, global_batch)
        self.t = t_before                                    # the capture pass recorded a step, it did not take one
        self._graph_io = (x, y_old)
        return self._graph

    def replay(self):
        """Replays the captured step; returns the mean per-sample norm tensor of this rank (overwritten by each replay)."""
        self._graph.replay()
        self.t += 1
        return self._graph_loss


# ----------------------------------------------------------------------------------------------------------------------
# All replaced layers at once: one flat fp32 buffer, one CUDA graph with a branch per layer, ONE fused
# reduce-scatter + AdamW + all-gather + re-pack kernel over NVLink peer memory per step (csrc/peer.cu).
# ----------------------------------------------------------------------------------------------------------------------
def plan_flat_layout(dims):
    """dims: [(K, C, N)] per layer (3-tap factors) -> ([(off_v, off_h, off_b, end)], n_padded).  Layer l owns
    [off_v, end) = v_conv.weight (K,C,3,1) | h_conv.weight (N,K,1,3) | h_conv.bias (N); n is padded to a multiple of 4."""
    offs, o = [], 0
    for (K, C, N) in dims:
        ov, oh = o, o + 3 * K * C
        ob = oh + 3 * N * K
        o = ob + N
        offs.append((ov, oh, ob, o))
    return offs, (o + 3) // 4 * 4


def owned_slice(n_padded: int, rank: int, world: int):
    """[lo, hi) in floats of the flat buffer that `rank` reduces, updates and broadcasts (the kernel's partition)."""
    n4 = n_padded // 4
    per = (n4 + world - 1) // world
    lo = min(n4, per * rank)
    return 4 * lo, 4 * min(n4, lo + per)


class LayerwiseReconstructor:
    """Layer-wise output-L2 reconstruction of every scheme-2 layer of a model in one step (the reference's
    train_each_layers.sh / L2Reconstruct flow in its symmetric mode, hooks/finetune.py:300-331: each replaced layer sees
    the teacher's input and is fitted to the teacher's output, so the layers' fits are independent).

        for each layer l (independent -> parallel branches of one CUDA graph):
            y, T = fused forward;  norm_b = ||y_b - y_old_b||;  g = dL/dy;  dWh, db, dWv on tcgen05
        ONE kernel: reduce-scatter of the flat gradient over NVLink peer memory, AdamW on the owned slice (optimizer state
        sharded over ranks), all-gather of the updated fp32 parameters, re-pack into the bf16 operand layouts

    loss = l2_weight * mean_b( sum_l norm_{l,b} / L )  (hooks/finetune.py:313-322 with cls_weight = 0); batch-sharded over
    the ranks of `process_group`.  Parameters live in a symmetric buffer; `sync_to_layers()` writes them back into the
    layers' nn.Parameters (the kernels read the packed copies, so this is only needed for checkpoints / inspection)."""

    def __init__(self, layers, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, l2_weight=1.0,
                 mean_over_layers=True, process_group=None):
        from approx.peer import SymmetricBuffer
        self.layers = list(layers)
        assert 1 <= len(self.layers) <= 16, "1..16 layers per reconstructor"
        for L in self.layers:
            if not (isinstance(L, LowRankExpConvV2) and L.dense and L.kernel_size == (3, 3) and L.stride == (1, 1)
                    and L.padding == (1, 1)):
                raise NotImplementedError("the multi-layer reconstructor covers dense 3x3 stride-1 'same' scheme-2 layers")
            if L.fuse_relu or L.fuse_pool:
                raise ValueError("reconstruction differentiates the bare layer: clear fuse_relu / fuse_pool first")
            if B.lowrank2_k_pad(L.num_base) is None or L.out_channels % 64 != 0:
                raise NotImplementedError("tensor-core reconstruction needs rank <= 64 and out_channels % 64 == 0")
        self.lr, self.betas, self.eps, self.wd, self.l2_weight = lr, betas, eps, weight_decay, l2_weight
        self.mean_over_layers = mean_over_layers
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        dev = self.layers[0].v_conv.weight.device
        B.require_cuda(self.layers[0].v_conv.weight)
        self.device = dev
        self.dims = [(L.num_base, L.in_channels, L.out_channels) for L in self.layers]
        self.offsets, self.n = plan_flat_layout(self.dims)
        hdr = int(B.lib().cna_recon_sync_header_bytes())
        self.off_grad = (hdr + 255) // 256 * 256
        self.off_param = self.off_grad + (self.n * 4 + 255) // 256 * 256
        self.sym = SymmetricBuffer(self.off_param + self.n * 4, dev, process_group)
        self.grad = self.sym.view(self.off_grad, self.n)
        self.param = self.sym.view(self.off_param, self.n)
        self.exp_avg = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.packs = []
        with torch.no_grad():
            for L, (K, C, N), (ov, oh, ob, end) in zip(self.layers, self.dims, self.offsets):
                self.param[ov:oh].copy_(L.v_conv.weight.detach().float().reshape(-1))
                self.param[oh:ob].copy_(L.h_conv.weight.detach().float().reshape(-1))
                self.param[ob:end].copy_(L.h_conv.bias.detach().float().reshape(-1))
            if self.world > 1:                       # identical starting point on every rank (rank 0's initialisation)
                dist.broadcast(self.param, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
            for (K, C, N), (ov, oh, ob, end) in zip(self.dims, self.offsets):
                kp, cp = B.lowrank2_k_pad(K), B.round_up(C, 16)
                wv = torch.zeros(3, kp, cp, device=dev, dtype=torch.bfloat16)
                wh = torch.zeros(3, N, kp, device=dev, dtype=torch.bfloat16)
                wd = torch.zeros(3, kp, N, device=dev, dtype=torch.bfloat16)
                bias = self.param[ob:end].clone()
                mv, mh = self.param[ov:oh].view(K, C, 3), self.param[oh:ob].view(N, K, 3)
                B.lowrank2_pack_into(mv, mh, wv, wh)
                B.lowrank2_pack_into(mh.permute(1, 0, 2).flip(2).contiguous(), torch.zeros(64, K, 3, device=dev), wd,
                                     torch.zeros(3, 64, kp, device=dev, dtype=torch.bfloat16))
                self.packs.append((wv, wh, wd, bias))
        import ctypes
        nl = len(self.layers)
        self._c_offsets = (ctypes.c_longlong * (4 * nl))(*[v for o in self.offsets for v in o])
        self._c_dims = (ctypes.c_int * (5 * nl))(*[v for (K, C, N) in self.dims
                                                    for v in (K, C, N, B.lowrank2_k_pad(K), B.round_up(C, 16))])
        self._c_ptrs = (ctypes.c_void_p * (4 * nl))(*[t.data_ptr() for pk in self.packs for t in pk])
        self._c_peers = self.sym.peer_array()
        self._bound = None
        self._graph = None
        self._streams = None
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------------------------------------------------------
    def bind(self, xs, y_olds, global_batch: Optional[int] = None):
        """Static per-layer input buffers: xs[l] (B,H,W,c_pad) bf16 NHWC (teacher-side input of layer l, channels zero-padded to
        16), y_olds[l] (B,H,W,N) bf16 (teacher output).  Copy new minibatches into them between steps."""
        assert len(xs) == len(y_olds) == len(self.layers)
        Bn = xs[0].shape[0]
        self.global_batch = Bn * self.world if global_batch is None else int(global_batch)
        dev = self.device
        self._bound = []
        for (K, C, N), x, yo in zip(self.dims, xs, y_olds):
            _, H, W, cx = x.shape
            kp = B.lowrank2_k_pad(K)
            assert x.dtype == yo.dtype == torch.bfloat16 and cx == B.round_up(C, 16) and yo.shape == (Bn, H, W, N)
            assert x.is_contiguous() and yo.is_contiguous() and x.shape[0] == Bn
            assert B.lowrank2_fused_supported(H, W, C, K, N, 3), "layer shape outside the fused kernel's envelope"
            self._bound.append(dict(x=x, y_old=yo,
                                    y=torch.empty((Bn, H, W, N), device=dev, dtype=torch.bfloat16),
                                    t=torch.empty((Bn, H, W, kp), device=dev, dtype=torch.bfloat16),
                                    g=torch.empty((Bn, H, W, N), device=dev, dtype=torch.bfloat16),
                                    gt=torch.empty((Bn, H, W, kp), device=dev, dtype=torch.bfloat16)))
        self.sumsq = torch.zeros(len(self.layers), Bn, device=dev, dtype=torch.float32)
        self.norms = torch.zeros(len(self.layers), Bn, device=dev, dtype=torch.float32)
        self._graph = None

    def _chain(self, l):
        (K, C, N), (ov, oh, ob, end), (wv, wh, wd, bias), io = self.dims[l], self.offsets[l], self.packs[l], self._bound[l]
        scale = self.l2_weight / self.global_batch / (len(self.layers) if self.mean_over_layers else 1)
        B.lowrank2_fused_with_t(io["x"], wv, wh, bias, out=(io["y"], io["t"]))
        B.l2_recon_into(io["y"], io["y_old"], scale, self.sumsq[l], self.norms[l], io["g"], self.grad[ob:end])
        B.wgrad3_tc(io["g"], io["t"], self.grad[oh:ob], N, K, 1, 0, accumulate=True)        # dL/dWh[n][k][t]
        B.tapconv3_tc(io["g"], wd, 1, out=io["gt"])                                           # dL/dT
        B.wgrad3_tc(io["x"], io["gt"], self.grad[ov:oh], C, K, 0, 1, accumulate=True)        # dL/dWv[k][c][t]

    def _sync(self, lr):
        import ctypes
        with torch.cuda.device(self.device):
            B._check(B.lib().cna_recon_sync_step(
                ctypes.cast(self._c_peers, ctypes.c_void_p), self.rank, self.world, self.n, self.off_grad, self.off_param,
                B._p(self.exp_avg), B._p(self.exp_avg_sq), B._p(self.step_dev), ctypes.cast(self._c_offsets, ctypes.c_void_p),
                ctypes.cast(self._c_dims, ctypes.c_void_p), ctypes.cast(self._c_ptrs, ctypes.c_void_p), len(self.layers),
                float(lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.wd), 1.0,
                torch.cuda.current_stream(self.device).cuda_stream))

    def step(self, parallel: bool = True, lr: Optional[float] = None):
        """One reconstruction step of every layer on the bound buffers (eager; `capture()` + `replay()` removes the host work).
        parallel: the layers' chains go to separate streams (they are independent) and join before the exchange kernel."""
        assert self._bound is not None, "bind() first"
        cur = torch.cuda.current_stream(self.device)
        self.grad.zero_()
        self.sumsq.zero_()
        if parallel and len(self.layers) > 1:
            if self._streams is None:
                self._streams = [torch.cuda.Stream(self.device) for _ in self.layers]
            fork = torch.cuda.Event()
            fork.record(cur)
            for l, st in enumerate(self._streams):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    self._chain(l)
                    done = torch.cuda.Event()
                    done.record(st)
                cur.wait_event(done)
        else:
            for l in range(len(self.layers)):
                self._chain(l)
        self._sync(self.lr if lr is None else lr)

    def capture(self, parallel: bool = True):
        """Captures step() into a CUDA graph.  The warm-up launches before the capture run with lr = 0 and the optimizer
        state is restored afterwards, so capturing does not advance the fit."""
        assert self._bound is not None, "bind() first"
        dev = self.device
        snap = (self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone())
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.step(parallel, lr=0.0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snap)
        if self.world > 1:
            dist.barrier(group=self.group)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.step(parallel)
        self._restore(snap)                      # (capture records, it does not execute — kept for symmetry / safety)
        torch.cuda.synchronize(dev)
        if self.world > 1:
            dist.barrier(group=self.group)
        return self._graph

    def _restore(self, snap):
        self.exp_avg.copy_(snap[0]); self.exp_avg_sq.copy_(snap[1]); self.step_dev.copy_(snap[2])
        # the barrier epochs follow the step counter: the flags left behind by the warm-up launches are ahead of it, and a
        # flag that is already >= the epoch a kernel waits for would let it through early -> clear them (all ranks do)
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        self.sym.view(0, 64, torch.int32).zero_()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def replay(self):
        self._graph.replay()

    def steps_taken(self) -> int:
        return int(self.step_dev.item())

    def loss(self) -> torch.Tensor:
        """l2_weight * mean over this rank's samples of sum_l norm / L, from the last step's per-sample norms."""
        tot = self.norms.sum(0) / (len(self.layers) if self.mean_over_layers else 1)
        return self.l2_weight * tot.mean()

    def check(self):
        """Raises if a peer barrier timed out in any step so far (a rank that never launched its step)."""
        import ctypes
        err = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            B._check(B.lib().cna_recon_sync_error(self.sym.local, ctypes.cast(ctypes.byref(err), ctypes.c_void_p)))
        if err.value:
            raise RuntimeError("cna_b200: a peer barrier of the reconstruction exchange timed out")

    @torch.no_grad()
    def sync_to_layers(self):
        for L, (ov, oh, ob, end) in zip(self.layers, self.offsets):
            L.v_conv.weight.copy_(self.param[ov:oh].view_as(L.v_conv.weight))
            L.h_conv.weight.copy_(self.param[oh:ob].view_as(L.h_conv.weight))
            L.h_conv.bias.copy_(self.param[ob:end])

    def close(self):
        self._graph = None
        self.grad = self.param = None
        self.sym.close()
