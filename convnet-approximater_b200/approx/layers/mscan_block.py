"""The FFN half of an MSCAN block and the whole block on this backend's kernels.

Reference: `MultiScaleConvAttnModule.forward` src/approx/models/mscan.py:108-112 and `FFN.forward` :56-62 (eval mode: dropout and
drop_path are identities):

    n = norm2(x1);  y = x1 + layer_scale_2 * fc2(gelu(dconv(fc1(n))))

Here (bf16 NHWC):

    h = x1 . W1'^T + b1'                  BatchNorm folded into fc1 (W1' = W1 diag(s), b1' = W1 o + b1), tcgen05 1x1 kernel, one launch
                                          whose CTAs split into groups per slab of <= 256 hidden channels (cna_conv1x1_tc_forward_strided)
    g = gelu(dconv3x3(h) + bd)            one TMA-pipelined pass (cna_dwconv3x3_act_forward)
    y = ls2 (g . W2^T + b2) + x1          fc2 + layer scale + residual in the 1x1 kernel's affine epilogue; the hidden channels are
                                          contracted in slabs whose weight fits in shared memory, accumulating into y in place

The hidden tensor (exp_ratio x C channels) makes three HBM round trips (written by fc1, read + written by the depthwise pass, read by
fc2) against ten in the module graph (BatchNorm, conv, bias, conv, bias, GELU, conv, bias, scale, add).
"""
import torch
from torch import nn

from approx import backend as B
from . import _ops
from .mscan_attention import MSCANAttention


class MSCANFFN(nn.Module):
    """Wraps the reference block's own `norm2`, `mlp` (fc1, dconv, fc2) and `layer_scale_2` (same parameter containers)."""

    def __init__(self, norm2: nn.BatchNorm2d, mlp: nn.Module, layer_scale_2: torch.Tensor):
        super().__init__()
        self.norm2 = norm2
        self.mlp = mlp
        self.layer_scale_2 = layer_scale_2 if isinstance(layer_scale_2, nn.Parameter) else nn.Parameter(layer_scale_2)

    @classmethod
    def from_block(cls, block: nn.Module) -> "MSCANFFN":
        return cls(block.norm2, block.mlp, block.layer_scale_2)

    def _folded(self, dtype):
        bn, fc1, dw, fc2 = self.norm2, self.mlp.fc1, self.mlp.dconv, self.mlp.fc2
        params = [bn.weight, bn.bias, bn.running_mean, bn.running_var, fc1.weight, fc1.bias, dw.weight, dw.bias, fc2.weight, fc2.bias,
                  self.layer_scale_2]

        def build():
            f = lambda t: t.detach().to(torch.float32)
            C, Hd = fc1.in_channels, fc1.out_channels
            s = f(bn.weight) / torch.sqrt(f(bn.running_var) + bn.eps)
            o = f(bn.bias) - f(bn.running_mean) * s
            w1 = f(fc1.weight).reshape(Hd, C)
            b1 = w1 @ o + (f(fc1.bias) if fc1.bias is not None else 0.0)
            w1p = B.conv1x1_pack((w1 * s[None, :]).reshape(Hd, C, 1, 1).contiguous())
            wd = B.pack_dw(f(dw.weight), torch.bfloat16)
            bd = f(dw.bias).contiguous() if dw.bias is not None else None
            w2p = B.conv1x1_pack(fc2.weight)
            b2 = f(fc2.bias).contiguous() if fc2.bias is not None else None
            return (w1p, b1.contiguous(), wd, bd, w2p, b2, f(self.layer_scale_2).contiguous())
        return _ops.cached(self, "mscan_ffn", params, dtype, build)

    @staticmethod
    def _k_slab(N: int, Hd: int) -> int:
        """Largest multiple of 32 hidden channels whose (N x slab) bf16 weight the 1x1 kernel keeps resident in shared memory."""
        k = Hd
        while k > 32 and not B.conv1x1_tc_supported(k, N):
            k -= 32
        return k

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("MSCANFFN: the fused block folds BatchNorm's running statistics - call .eval() first")
        if x.dtype != torch.bfloat16:
            raise RuntimeError("MSCANFFN: the fused block runs on the bf16 tensor-core kernels only")
        fc1, dw = self.mlp.fc1, self.mlp.dconv
        C, Hd = fc1.in_channels, fc1.out_channels
        if C % 32 or Hd % 32 or tuple(dw.kernel_size) != (3, 3) or dw.groups != Hd or tuple(dw.stride) != (1, 1) or tuple(dw.padding) != (1, 1):
            raise RuntimeError("MSCANFFN: needs channel counts that are multiples of 32 and a stride-1 3x3 depthwise conv")
        w1p, b1, wd, bd, w2p, b2, ls = self._folded(x.dtype)
        x = x.contiguous()
        Bn, H, W, _ = x.shape
        P = Bn * H * W
        x2 = x.view(P, C)
        h = torch.empty((Bn, H, W, Hd), device=x.device, dtype=x.dtype)
        h2 = h.view(P, Hd)
        B.conv1x1_tc_slab(x2, 0, C, w1p, 0, Hd, b1, h2)          # one launch: the kernel grids over slabs of 256 hidden channels
        g = B.dwconv3x3_act(h, wd, bd, "gelu")
        g2 = g.view(P, Hd)
        y = torch.empty_like(x)
        y2 = y.view(P, C)
        ks = self._k_slab(C, Hd)
        for k0 in range(0, Hd, ks):
            kc = min(ks, Hd - k0)
            if k0 == 0:
                B.conv1x1_tc_slab(g2, 0, kc, w2p, 0, C, b2, y2, aux2d=x2, alpha=ls)
            else:
                B.conv1x1_tc_slab(g2, k0, kc, w2p, 0, C, None, y2, aux2d=y2, alpha=ls)
        return y

    @_ops.empty_batch_safe
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _ops.check_input(x)
        return B.from_nhwc(self.forward_nhwc(B.to_nhwc(x)))


class MSCANBlock(nn.Module):
    """A whole `MultiScaleConvAttnModule` (reference models/mscan.py:83-112) in eval mode: the fused attention half followed by the
    fused FFN half.  `from_block` wraps a block built by the reference's model code on this backend's layers; `state_dict` keys stay
    the block's own (`norm1.*`, `attn.*`, `norm2.*`, `mlp.*`, `layer_scale_1/2`)."""

    def __init__(self, norm1, attn, layer_scale_1, norm2, mlp, layer_scale_2):
        super().__init__()
        self.norm1, self.attn, self.norm2, self.mlp = norm1, attn, norm2, mlp
        self.layer_scale_1 = layer_scale_1 if isinstance(layer_scale_1, nn.Parameter) else nn.Parameter(layer_scale_1)
        self.layer_scale_2 = layer_scale_2 if isinstance(layer_scale_2, nn.Parameter) else nn.Parameter(layer_scale_2)
        object.__setattr__(self, "_attn_half", MSCANAttention(norm1, attn, self.layer_scale_1))       # views on the same parameters,
        object.__setattr__(self, "_ffn_half", MSCANFFN(norm2, mlp, self.layer_scale_2))               # not registered twice

    @classmethod
    def from_block(cls, block: nn.Module) -> "MSCANBlock":
        return cls(block.norm1, block.attn, block.layer_scale_1, block.norm2, block.mlp, block.layer_scale_2)

    def train(self, mode: bool = True):
        super().train(mode)
        self._attn_half.train(mode)
        self._ffn_half.train(mode)
        return self

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        return self._ffn_half.forward_nhwc(self._attn_half.forward_nhwc(x))

    @_ops.empty_batch_safe
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _ops.check_input(x)
        return B.from_nhwc(self.forward_nhwc(B.to_nhwc(x)))
