"""Strip-convolution building blocks of MSCA and of its re-parameterised forms.

Drop-in for the reference's src/approx/layers/depth_seperable_conv.py (module name spelling kept):
`CascadeConv`, `ParallelConv`, `FixPaddingBias` with identical constructors, attributes
(`conv1/conv2`, `branches`, `dim`, `res`, `p`) and state_dict keys.  Forward passes run this
library's depthwise kernels on NHWC data; ParallelConv's Python `sum()` and FixPaddingBias'
in-place row adds are fused into the kernel that produces the tensor they touch.
"""
import torch
from torch import nn

from approx import backend as B
from . import _ops


class CascadeConv(nn.Module):
    """Depthwise (1,k) conv [bias iff `first_bias`] then depthwise (k,1) conv [bias iff `bias`]
    (reference :5-16)."""

    def __init__(self, dim, kernel_size, padding, bias, first_bias):
        super().__init__()
        self.conv1 = nn.Conv2d(dim, dim, (1, kernel_size), padding=(0, padding), groups=dim, bias=first_bias)
        self.conv2 = nn.Conv2d(dim, dim, (kernel_size, 1), padding=(padding, 0), groups=dim, bias=bias)

    def forward_nhwc(self, x, res=None, add=None):
        return _ops.dw(self.conv2, _ops.dw(self.conv1, x), res=res, add=add)

    @_ops.empty_batch_safe
    def forward(self, x):
        _ops.check_input(x)
        return B.from_nhwc(self.forward_nhwc(B.to_nhwc(x)))


class ParallelConv(nn.Module):
    """Sum of `nbranch` CascadeConvs (+ the input itself when `identity`), reference :19-47.
    `all_bias=False` leaves a single bias, on the last branch's `conv2`."""

    def __init__(self, dim, kernel_sizes, paddings, nbranch, all_bias, identity):
        super().__init__()
        self.dim = dim
        if isinstance(kernel_sizes, int):
            kernel_sizes = [kernel_sizes] * nbranch
        if isinstance(paddings, int):
            paddings = [paddings] * nbranch
        assert len(kernel_sizes) == nbranch
        assert len(paddings) == nbranch
        self.branches = nn.ModuleList()
        for i in range(nbranch):
            last = i == nbranch - 1
            self.branches.append(CascadeConv(dim=dim, kernel_size=kernel_sizes[i], padding=paddings[i],
                                             bias=True if all_bias else last, first_bias=bool(all_bias)))
        if identity:
            self.branches.append(nn.Identity())

    def cascades(self):
        return [b for b in self.branches if isinstance(b, CascadeConv)]

    @property
    def has_identity(self) -> bool:
        return any(isinstance(b, nn.Identity) for b in self.branches)

    def forward_nhwc(self, x, res=None):
        """Running sum in the reference's order ((b0 + b1) + b2 ... + x); `res` (FixPaddingBias) is applied
        by the caller on the finished sum."""
        acc = None
        for br in self.cascades():
            acc = br.forward_nhwc(x, add=acc)
        if self.has_identity:
            acc = B.add(acc, x)
        return acc

    @_ops.empty_batch_safe
    def forward(self, x):
        _ops.check_input(x)
        return B.from_nhwc(self.forward_nhwc(B.to_nhwc(x)))


class FixPaddingBias(nn.Module):
    """Adds per-channel residues to the top / bottom `padding` rows, IN PLACE (reference :50-61); it
    repairs the bias a merged kernel applies where the original cascade saw zero padding."""

    def __init__(self, num_channels, padding):
        super().__init__()
        self.p = padding
        self.res = nn.Parameter(torch.randn(2, num_channels, padding), requires_grad=True)

    def res_f32(self):
        return _ops.cached(self, "res_f32", (self.res,), torch.float32,
                           lambda: self.res.detach().to(torch.float32).contiguous())

    def forward_nhwc_(self, x):
        return B.fix_padding_bias_(x, self.res_f32())

    def forward(self, x):
        _ops.check_input(x)
        if x.numel() == 0:              # nothing to patch (the reference's slice assignments are no-ops on an empty batch)
            return x
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous():           # channels_last: patch the caller's storage directly
            self.forward_nhwc_(v)
            return x
        y = self.forward_nhwc_(B.to_nhwc(x))
        x.copy_(B.from_nhwc(y))         # keep the reference's in-place contract for NCHW callers
        return x
