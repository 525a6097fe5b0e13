"""SegNeXt multi-scale convolutional attention (MSCA) on the B200 backend.

Drop-in for the reference's src/approx/layers/msca.py:8-42: `x * channel_mix(sd_convs(conv0(x)))`.
`sd_convs` may be the original ParallelConv (3 strip cascades + identity) or anything
`MscaRep` puts in its place: a depthwise KxK conv, a CascadeConv, a ParallelConv of SVD
branches, each optionally wrapped as `nn.Sequential(<conv>, FixPaddingBias)`.
"""
import torch
from torch import nn

from approx import backend as B
from . import _ops
from .depth_seperable_conv import CascadeConv, FixPaddingBias, ParallelConv
from .substituton import LAYER


def run_sd_convs(sd, a):
    """Apply an `sd_convs` module (any of the forms above) to an NHWC tensor."""
    fix = None
    if isinstance(sd, nn.Sequential):
        assert len(sd) == 2 and isinstance(sd[1], FixPaddingBias), "expected Sequential(conv, FixPaddingBias)"
        sd, fix = sd[0], sd[1]
    res = fix.res_f32() if fix is not None else None
    if isinstance(sd, ParallelConv):
        y = sd.forward_nhwc(a)
        return fix.forward_nhwc_(y) if fix is not None else y
    if isinstance(sd, CascadeConv):
        return sd.forward_nhwc(a, res=res)
    if isinstance(sd, nn.Conv2d):
        return _ops.dw(sd, a, res=res)
    raise TypeError(f"unsupported sd_convs module {type(sd).__name__}")


@LAYER.register_module()
class MSCA(nn.Module):
    def __init__(self, num_channel, k1_size, k_sizes):
        super().__init__()
        self.num_channel = num_channel
        self.k1_size = k1_size
        self.k_sizes = k_sizes
        self.conv0 = nn.Conv2d(num_channel, num_channel, k1_size, padding=k1_size // 2, groups=num_channel)
        self.sd_convs = ParallelConv(num_channel, k_sizes, [k // 2 for k in k_sizes], len(k_sizes), True, True)
        self.channel_mix = nn.Conv2d(num_channel, num_channel, 1)

    def forward_nhwc(self, x):
        from . import _msca_fused
        a = _msca_fused.conv0_sd_forward(self, x)
        if a is None:
            a = run_sd_convs(self.sd_convs, _ops.dw(self.conv0, x))
        return _ops.dense_taps(self.channel_mix, a, gate=x)

    @_ops.empty_batch_safe
    def forward(self, x):
        _ops.check_input(x)
        # model.train() + autograd (the reference's L2Reconstruct fine-tuning, hooks/finetune.py:275,329): op-by-op path with
        # backward kernels; model.eval() or no_grad (every inference hook of the reference): the fused kernels
        if (self.training and torch.is_grad_enabled()
                and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            from . import _msca_autograd
            return _msca_autograd.msca_train_forward(self, x)
        return B.from_nhwc(self.forward_nhwc(B.to_nhwc(x)))

    def switchable_layer(self):
        return self.sd_convs


@LAYER.register_module()
class MSCAProfile(MSCA):
    """Same math with the reference's three profiler ranges (msca.py:32-42): CONV0 / SD_CONVS / CHANNEL_MIX
    as NVTX ranges + torch.profiler record_function ranges.  Runs the un-fused kernels so each range is
    attributable."""

    def forward_nhwc(self, x):
        import torch.autograd.profiler as profiler
        with profiler.record_function("CONV0"):
            a = _ops.dw(self.conv0, x)
        with profiler.record_function("SD_CONVS"):
            a = run_sd_convs(self.sd_convs, a)
        with profiler.record_function("CHANNEL_MIX"):
            a = _ops.dense_taps(self.channel_mix, a, gate=x)
        return a
