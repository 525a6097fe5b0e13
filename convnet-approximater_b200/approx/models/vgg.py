"""VGG-16 convolutional trunk with every 3x3 conv replaced by a scheme-2 low-rank expansion.

The reference ships AlexNet and MSCAN only; BASELINE.json's headline configuration ("VGG-16, all 3x3
conv layers replaced by scheme-2 low-rank expansions, bf16 inference, batch 256 at 224x224") is built
here from the same layer class the registry exposes (`LowRankExpConvV2` with a dense h_conv).
The trunk runs NHWC end to end: 13 fused low-rank convs with ReLU — and, on the 5 layers VGG pools after, the
2x2 max-pool — fused into the kernel's epilogue.
"""
from typing import Sequence, Union

import torch
from torch import nn

from approx import backend as B
from approx.layers import LowRankExpConvV2

# (in_channels, out_channels) per conv, 'M' = 2x2 max-pool
VGG16_CFG = [(3, 64), (64, 64), "M", (64, 128), (128, 128), "M", (128, 256), (256, 256), (256, 256), "M",
             (256, 512), (512, 512), (512, 512), "M", (512, 512), (512, 512), (512, 512), "M"]


class VGG16Scheme2Trunk(nn.Module):
    def __init__(self, rank: Union[int, Sequence[int]] = 64):
        super().__init__()
        convs = [c for c in VGG16_CFG if c != "M"]
        ranks = [rank] * len(convs) if isinstance(rank, int) else list(rank)
        assert len(ranks) == len(convs)
        self.features = nn.ModuleList()
        it = iter(ranks)
        for i, c in enumerate(VGG16_CFG):
            if c == "M":
                continue                       # the pool is fused into the preceding conv's epilogue
            layer = LowRankExpConvV2(c[0], next(it), (3, 3), (1, 1), (1, 1), out_channels=c[1])
            layer.fuse_relu = True
            layer.fuse_pool = i + 1 < len(VGG16_CFG) and VGG16_CFG[i + 1] == "M"
            self.features.append(layer)

    def set_normalization(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        """Per-channel mean / std (on the 0..1 scale, like timm's IMAGENET_DEFAULT_MEAN / STD) applied to uint8 inputs."""
        self._norm = (tuple(float(m) * 255.0 for m in mean), tuple(1.0 / (float(s) * 255.0) for s in std))
        self._norm_dev = {}            # device -> (mean255, inv_std255) fp32 tensors (plain attributes: .bfloat16() must not round them)

    def _norm_tensors(self, device):
        if not hasattr(self, "_norm"):
            self.set_normalization()
        if device not in self._norm_dev:
            self._norm_dev[device] = tuple(torch.tensor(v, dtype=torch.float32, device=device) for v in self._norm)
        return self._norm_dev[device]

    def conv_layers(self):
        return [m for m in self.features if isinstance(m, LowRankExpConvV2)]

    def forward_nhwc(self, x):
        """x: (B,H,W,16) bf16 (3 image channels zero-padded to 16) or (B,H,W,3) fp32."""
        for m in self.features:
            x = m.forward_nhwc(x)
        return x

    def forward(self, x):
        """x: (B,3,H,W) logical NCHW, bf16 (fused tcgen05 path) or fp32 (exact FFMA path) — or uint8 images as an image loader
        holds them: normalised on the device, (x - 255 mean) / (255 std), in the same kernel that lays them out as bf16 NHWC."""
        if x.dtype == torch.uint8:
            mean255, inv_std255 = self._norm_tensors(x.device)
            return B.from_nhwc(self.forward_nhwc(B.u8_to_nhwc_norm(x.contiguous(), mean255, inv_std255, 16)))
        c_pad = 16 if x.dtype == torch.bfloat16 else None
        return B.from_nhwc(self.forward_nhwc(B.to_nhwc(x, c_pad)))
