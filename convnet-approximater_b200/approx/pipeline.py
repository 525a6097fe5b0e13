"""Host-buffer inference pipeline: the call a user with images in (pinned) host memory makes.

The reference runs `model(x.cuda())` batch after batch on one stream (hooks/fps.py:62-71), so every batch pays
H2D copy + compute + D2H copy back to back.  Here the three legs run on three CUDA streams with a ring of
device input buffers: the H2D copy of batch i+1 and the D2H copy of batch i-1 overlap the kernels of batch i.
Results are bit-identical to the one-stream order (same kernels, same inputs); only the schedule changes.
"""
from __future__ import annotations

import torch

from approx import backend as B


class HostPipeline:
    def __init__(self, module, device, depth: int = 2):
        self.module, self.device, self.depth = module, torch.device(device), depth
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.xbuf = [None] * depth
        self.ev_in_ready = [torch.cuda.Event() for _ in range(depth)]
        self.ev_in_free = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out_ready = [torch.cuda.Event() for _ in range(depth)]
        self.n = 0

    @torch.no_grad()
    def submit(self, x_host: torch.Tensor, y_host: torch.Tensor):
        """Queue one batch: x_host (pinned, any dtype/layout the module accepts) -> y_host (pinned).  Returns at once;
        call `wait()` before reading y_host."""
        i = self.n % self.depth
        self.n += 1
        if self.xbuf[i] is None or self.xbuf[i].shape != x_host.shape or self.xbuf[i].dtype != x_host.dtype:
            self.xbuf[i] = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_in_free[i])        # the batch that used this slot has been consumed
            self.xbuf[i].copy_(x_host, non_blocking=True)
            self.ev_in_ready[i].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in_ready[i])
            y = self.module(self.xbuf[i])
            if (y.dim() == 4 and y_host.is_contiguous() and not y.is_contiguous() and y.permute(0, 2, 3, 1).is_contiguous()
                    and y.dtype == y_host.dtype):
                # the module hands back a channels_last view; an NCHW-contiguous host buffer needs the transpose: the library's
                # tile-transpose kernel here, instead of the strided element copy torch would insert before the D2H
                y = B.to_nchw_contiguous(y.permute(0, 2, 3, 1))
            self.ev_in_free[i].record(self.s_run)
            self.ev_out_ready[i].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_out_ready[i])
            y_host.copy_(y, non_blocking=True)
            y.record_stream(self.s_out)

    def wait(self):
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()
