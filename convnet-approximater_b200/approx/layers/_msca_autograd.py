"""Training path of the MSCA block and its re-parameterised forms (`model.train()` + autograd), op by op.

The reference trains these layers through autograd over `nn.Conv2d` / in-place adds / `*` (reached from
`total_loss.backward()` in src/approx/hooks/finetune.py:329-330 when an MSCA block's `sd_convs` are the trainable replaced
layer, models/switchable.py:63-84).  Here every op is a `torch.autograd.Function` around this library's kernels on NHWC
tensors — forward: the same depthwise / 1x1 kernels the inference path uses (unfused, so every intermediate the backward needs
exists); backward: the data gradient of a stride-1 'same' depthwise conv is the forward kernel on the flipped filter, the
weight gradients / border-residue gradients / gate products are the kernels of csrc/dw_backward.cu and csrc/recon.cu.
torch only strings the ops together (graph bookkeeping, gradient accumulation).
"""
import torch
from torch import nn

from approx import backend as B
from . import _ops


def _f32_grad_like(param, g32):
    return None if g32 is None else g32.view_as(param).to(param.dtype)


class _DwConv(torch.autograd.Function):
    """Depthwise conv (groups = Cin, any channel multiplier, stride 1, 'same' padding) on an NHWC tensor."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv):
        ctx.conv = conv
        ctx.save_for_backward(x)
        return _ops.dw(conv, x)

    @staticmethod
    def backward(ctx, g):
        conv = ctx.conv
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        kh, kw = conv.kernel_size
        gx = gw = gb = None
        mult = conv.out_channels // conv.in_channels
        if ctx.needs_input_grad[0]:
            if mult == 1:      # the forward (register-tiled) kernels on the flipped filter
                wflip = _ops.cached(conv, "dw_w_flipped", (conv.weight,), g.dtype,
                                    lambda: B.pack_dw(conv.weight.detach().flip(2, 3).contiguous(), g.dtype))
                gx = B.dwconv(g, wflip, None, 1, kh, kw, (1, 1), (kh // 2, kw // 2))
            else:
                wpk = _ops.cached(conv, "dw_w", (conv.weight,), g.dtype, lambda: B.pack_dw(conv.weight, g.dtype))
                gx = B.dwconv_dgrad(g, wpk, mult, kh, kw)
        if ctx.needs_input_grad[1]:
            gw = _f32_grad_like(conv.weight, B.dwconv_wgrad(x, g, kh, kw))
        if conv.bias is not None and ctx.needs_input_grad[2]:
            gb = _f32_grad_like(conv.bias, B.bias_grad(g, torch.empty(g.shape[3], device=g.device, dtype=torch.float32)))
        return gx, gw, gb, None


class _FixBias(torch.autograd.Function):
    """FixPaddingBias on an NHWC tensor (out of place here: autograd needs the un-patched producer output intact)."""

    @staticmethod
    def forward(ctx, x, res, fix):
        ctx.fix = fix
        y = x.clone()
        return B.fix_padding_bias_(y, fix.res_f32())

    @staticmethod
    def backward(ctx, g):
        fix = ctx.fix
        g = g.contiguous()
        gres = None
        if ctx.needs_input_grad[1]:
            gres = _f32_grad_like(fix.res, B.fix_padding_bias_grad(g, fix.p))
        return g, gres, None


class _Conv1x1(torch.autograd.Function):
    """1x1 conv + bias on an NHWC tensor (channel_mix)."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv):
        ctx.conv = conv
        ctx.save_for_backward(x)
        return _ops.dense_taps(conv, x)

    @staticmethod
    def backward(ctx, g):
        conv = ctx.conv
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wt = _ops.cached(conv, "taps_w_dgrad", (conv.weight,), g.dtype, lambda: B.pack_taps_dgrad(conv.weight, g.dtype))
            gx = B.conv_taps(g, wt, None, conv.in_channels, 1, 1, (1, 1), 0)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = B.conv_taps_wgrad(x, g, 1, 1, 0, want_bias=conv.bias is not None)
            gw = _f32_grad_like(conv.weight, dw) if ctx.needs_input_grad[1] else None
            gb = _f32_grad_like(conv.bias, db) if (conv.bias is not None and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


class _Mul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return B.mul(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous()
        return (B.mul(g, b) if ctx.needs_input_grad[0] else None), (B.mul(g, a) if ctx.needs_input_grad[1] else None)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return B.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


def _dw(conv, x):
    ok = (conv.groups == conv.in_channels and conv.out_channels % conv.in_channels == 0 and tuple(conv.stride) == (1, 1)
          and conv.dilation == (1, 1)
          and tuple(conv.padding) == (conv.kernel_size[0] // 2, conv.kernel_size[1] // 2)
          and conv.kernel_size[0] % 2 == 1 and conv.kernel_size[1] % 2 == 1)
    if not ok:
        raise NotImplementedError("training kernels cover stride-1 'same' depthwise convolutions")
    return _DwConv.apply(x, conv.weight, conv.bias, conv)


def sd_convs_train(sd, a):
    """Autograd-capable `sd_convs(a)` for every form MSCA / MscaRep builds (NHWC in, NHWC out)."""
    from .depth_seperable_conv import CascadeConv, FixPaddingBias, ParallelConv
    fix = None
    if isinstance(sd, nn.Sequential):
        assert len(sd) == 2 and isinstance(sd[1], FixPaddingBias), "expected Sequential(conv, FixPaddingBias)"
        sd, fix = sd[0], sd[1]
    if isinstance(sd, ParallelConv):
        s = None
        for br in sd.cascades():                         # the reference's sum(): ((b0 + b1) + b2 ...) + x
            y = _dw(br.conv2, _dw(br.conv1, a))
            s = y if s is None else _Add.apply(s, y)
        if sd.has_identity:
            s = _Add.apply(s, a)
    elif isinstance(sd, CascadeConv):
        s = _dw(sd.conv2, _dw(sd.conv1, a))
    elif isinstance(sd, nn.Conv2d):
        s = _dw(sd, a)
    else:
        raise TypeError(f"unsupported sd_convs module {type(sd).__name__}")
    if fix is not None:
        s = _FixBias.apply(s, fix.res, fix)
    return s


class _ToNHWC(torch.autograd.Function):
    """NCHW-contiguous -> NHWC-contiguous through the library's transpose kernel; the gradient comes back as a channels_last view."""

    @staticmethod
    def forward(ctx, x):
        return B.to_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return g.permute(0, 3, 1, 2)


def _nhwc(x):
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else _ToNHWC.apply(x)      # channels_last input: a free view autograd already understands


def msca_train_forward(msca, x):
    """x (B,C,H,W) -> x * channel_mix(sd_convs(conv0(x))) with a full autograd graph over this library's kernels."""
    xn = _nhwc(x)
    a0 = _dw(msca.conv0, xn)
    s = sd_convs_train(msca.sd_convs, a0)
    m = _Conv1x1.apply(s, msca.channel_mix.weight, msca.channel_mix.bias, msca.channel_mix)
    return _Mul.apply(m, xn).permute(0, 3, 1, 2)


def lowrank1_train_forward(layer, x):
    """LowRankExpConvV1 (scheme 1) with a full autograd graph: s_conv (grouped d x d conv with channel multiplier, or its
    separable d x 1 . 1 x d form after decomp()) then the 1x1 d_conv (reference layers/low_rank_conv.py:51-52)."""
    from .low_rank_conv import SeparableConv
    xn = _nhwc(x)
    if isinstance(layer.s_conv, SeparableConv):
        t = _dw(layer.s_conv.h_conv, _dw(layer.s_conv.v_conv, xn))
    else:
        t = _dw(layer.s_conv, xn)
    y = _Conv1x1.apply(t, layer.d_conv.weight, layer.d_conv.bias, layer.d_conv)
    return y.permute(0, 3, 1, 2)
