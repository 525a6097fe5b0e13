"""Dispatch to the fused MSCA kernels (csrc/msca.cu) when a configuration is supported.

Each function returns None when the fused kernel does not apply; callers then run the
generic depthwise kernels (still CUDA, just more launches).
"""
import os

import torch
from torch import nn

from approx import backend as B
from . import _ops


def _branches(sd):
    """-> (list of CascadeConv, identity flag, FixPaddingBias | None) or None if `sd` has another shape."""
    from .depth_seperable_conv import CascadeConv, FixPaddingBias, ParallelConv
    fix = None
    if isinstance(sd, nn.Sequential):
        if len(sd) != 2 or not isinstance(sd[1], FixPaddingBias):
            return None
        sd, fix = sd[0], sd[1]
    if isinstance(sd, ParallelConv):
        return sd.cascades(), sd.has_identity, fix
    if isinstance(sd, CascadeConv):
        return [sd], False, fix
    return None


def _packed_chain(msca, cas, dtype):
    """fp32 tap-major weight blocks for the fused kernel, cached on the MSCA module."""
    params = [msca.conv0.weight, msca.conv0.bias]
    for c in cas:
        params += [c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias]

    def build():
        dev = msca.conv0.weight.device
        C = msca.num_channel
        zero = torch.zeros(C, device=dev, dtype=torch.float32)
        f = lambda t: zero if t is None else t.detach().to(torch.float32).contiguous()
        w0 = B.pack_dw(msca.conv0.weight, torch.float32)
        w1 = torch.cat([B.pack_dw(c.conv1.weight, torch.float32) for c in cas]).contiguous()
        w2 = torch.cat([B.pack_dw(c.conv2.weight, torch.float32) for c in cas]).contiguous()
        b1 = torch.stack([f(c.conv1.bias) for c in cas]).contiguous()
        b2 = torch.stack([f(c.conv2.bias) for c in cas]).contiguous()
        return w0, f(msca.conv0.bias), w1, b1, w2, b2
    return _ops.cached(msca, "sd_chain", params, dtype, build)


def conv0_sd_forward(msca, x):
    """conv0 + sd_convs of an MSCA block as one kernel, or None."""
    mode = os.environ.get("CNA_MSCA_FUSED", "auto")
    if mode == "0":
        return None
    got = _branches(msca.sd_convs)
    if got is None:
        return None
    cas, identity, fix = got
    conv0 = msca.conv0
    if (conv0.kernel_size != (5, 5) or conv0.stride != (1, 1) or conv0.padding != (2, 2)
            or conv0.groups != conv0.in_channels or conv0.out_channels != conv0.in_channels):
        return None
    ks = []
    for c in cas:
        k = c.conv1.kernel_size[1]
        if (c.conv1.kernel_size != (1, k) or c.conv2.kernel_size != (k, 1) or c.conv1.padding != (0, k // 2)
                or c.conv2.padding != (k // 2, 0) or c.conv1.stride != (1, 1) or c.conv2.stride != (1, 1)):
            return None
        ks.append(k)
    if not ks:
        return None
    if mode == "chain" and B.msca_chain_supported(msca.num_channel, 5, ks, x.dtype):
        # CNA_MSCA_FUSED=chain: the register-tiled 5x5 + strip kernels enqueued back to back by one C call (three HBM round trips per
        # cascade).  Measured on B200 at batch 32 over MSCAN-T's 13 blocks (profiles/r02_msca_mma.txt): the one-pass marching
        # kernel below wins for every form — original 0.98 vs 2.76 ms, MscaRep d1+fix 1.02 vs 1.64 ms — so it is the default.
        dw_w = lambda conv: _ops.cached(conv, "dw_w", (conv.weight,), x.dtype, lambda: B.pack_dw(conv.weight, x.dtype))
        return B.msca_chain(x.contiguous(), dw_w(conv0), _ops.bias_f32(conv0),
                            [dw_w(c.conv1) for c in cas], [_ops.bias_f32(c.conv1) for c in cas],
                            [dw_w(c.conv2) for c in cas], [_ops.bias_f32(c.conv2) for c in cas],
                            ks, identity, res=None if fix is None else fix.res_f32())
    if not B.msca_sd_supported(msca.num_channel, 5, ks, x.dtype):
        return None
    w0, b0, w1, b1, w2, b2 = _packed_chain(msca, cas, x.dtype)
    return B.msca_sd(x.contiguous(), w0, b0, w1, b1, w2, b2, ks, identity, res=None if fix is None else fix.res_f32())
