"""The attention half of an MSCAN block around MSCA, fused into four kernels.

Reference: `MultiScaleConvAttnModule.forward` src/approx/models/mscan.py:108-109 and `SpatialAttention.forward` :72-80 (eval mode,
drop_path = identity):

    n = norm1(x);  a = proj_2(MSCA(gelu(proj_1(n)))) + n;  y = x + layer_scale_1 * a

Here (bf16 NHWC, tcgen05 1x1 convs, csrc/conv1x1_tc.cu + the MSCA kernels):

    u = gelu(x . W1'^T + b1')                      BatchNorm folded into proj_1:  W1' = W1 diag(s), b1' = W1 o + b1,
                                                   s = gamma / sqrt(var + eps), o = beta - mean s          (kernel 1, GELU epilogue)
    v = MSCA(u)                                    depthwise chain + mix/gate                              (kernels 2, 3)
    y = ls (v . W2^T + b2) + (1 + ls s) x + ls o   proj_2, the shortcut n = s x + o, layer_scale_1 and the outer residual in ONE
                                                   epilogue: alpha = ls, beta = 1 + ls s, delta = ls o     (kernel 4)

against ~12 launches and 12 HBM passes of the module graph.  Folding changes where bf16 rounding happens (n, proj_1's output and the
three adds of the tail are not materialised), so parity with the reference's bf16 modules is to bf16 resolution of the BRANCH
(tests/test_mscan_attention_gpu.py states the tolerance), not bit-level.
"""
import torch
from torch import nn

from approx import backend as B
from . import _ops


class MSCANAttention(nn.Module):
    """Wraps the reference block's own submodules (same parameter containers, nothing is copied): `norm1`, `attn` (with `proj_1`,
    `spatial_gating_unit` = this backend's MSCA or a substituted form, `proj_2`) and `layer_scale_1`."""

    def __init__(self, norm1: nn.BatchNorm2d, attn: nn.Module, layer_scale_1: torch.Tensor):
        super().__init__()
        self.norm1 = norm1
        self.attn = attn
        self.layer_scale_1 = layer_scale_1 if isinstance(layer_scale_1, nn.Parameter) else nn.Parameter(layer_scale_1)

    @classmethod
    def from_block(cls, block: nn.Module) -> "MSCANAttention":
        return cls(block.norm1, block.attn, block.layer_scale_1)

    def _folded(self, dtype):
        bn, p1, p2 = self.norm1, self.attn.proj_1, self.attn.proj_2
        params = [bn.weight, bn.bias, bn.running_mean, bn.running_var, p1.weight, p1.bias, p2.weight, p2.bias, self.layer_scale_1]

        def build():
            f = lambda t: t.detach().to(torch.float32)
            C = p1.in_channels
            s = f(bn.weight) / torch.sqrt(f(bn.running_var) + bn.eps)
            o = f(bn.bias) - f(bn.running_mean) * s
            w1 = f(p1.weight).reshape(p1.out_channels, C)
            b1 = w1 @ o + (f(p1.bias) if p1.bias is not None else 0.0)
            w1p = B.conv1x1_pack((w1 * s[None, :]).reshape(p1.out_channels, C, 1, 1).contiguous())
            w2p = B.conv1x1_pack(p2.weight)
            ls = f(self.layer_scale_1)
            b2 = f(p2.bias).contiguous() if p2.bias is not None else None
            return (w1p, b1.contiguous(), w2p, b2, ls.contiguous(), (1.0 + ls * s).contiguous(), (ls * o).contiguous())
        return _ops.cached(self, "mscan_attn", params, dtype, build)

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("MSCANAttention: the fused block folds BatchNorm's running statistics — call .eval() first")
        if x.dtype != torch.bfloat16:
            raise RuntimeError("MSCANAttention: the fused block runs on the bf16 tensor-core kernels only")
        C = self.attn.proj_1.in_channels
        if not B.conv1x1_tc_supported(C, C):
            raise RuntimeError(f"MSCANAttention: {C} channels are not supported by the tcgen05 1x1 kernel (multiples of 32)")
        w1p, b1, w2p, b2, alpha, beta, delta = self._folded(x.dtype)
        x = x.contiguous()
        u = B.conv1x1_tc_ex(x, w1p, b1, act="gelu")
        v = self.attn.spatial_gating_unit.forward_nhwc(u)
        return B.conv1x1_tc_ex(v, w2p, b2, residual=x, alpha=alpha, beta=beta, delta=delta)

    @_ops.empty_batch_safe
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _ops.check_input(x)
        return B.from_nhwc(self.forward_nhwc(B.to_nhwc(x)))
